// Swin W-MSA / SW-MSA core on the 5th-generation tensor cores (mega_core/modeling/backbone/swintransformer.py:145-176:
// softmax(q*scale k^T + relative position bias + shift mask) v per (window, head), head dim 32, 7x7 windows).
//
// Same contract as swin_window_attention_kernel (swin.cu).  One CTA = TWO windows of one head, each padded from 49 to
// 64 rows, so that one tcgen05 tile is filled:
//   S   = Q[128 x 32] . K^T[32 x 128]        M=128, N=128, K=32 -> TMEM columns 0..127; rows 0..63 x columns 0..63 and
//                                            rows 64..127 x columns 64..127 are the two windows' score blocks, the
//                                            off-diagonal blocks are never read
//   P   = softmax over the 49 real keys      one thread per query row (== TMEM lane): tcgen05.ld of its window's 64
//                                            columns, + bias (staged in shared memory) + shift mask (token regions), exp2
//   O_w = P_w[128 x 64] . V_w[64 x 32]       two MMAs (M=128, N=32, K=64) -> TMEM columns 128..159 / 160..191; rows of the
//                                            other window are don't-care in each
// Operand tiles are K-major, 128-byte swizzled (rows of 32 dims use the first 64 bytes of a 128-byte row); V is
// transposed into shared memory with 8x8 shuffle transposes.  Warps 0..3 softmax (warps 0,1 = window A, 2,3 = window
// B), warp 4 issues the MMAs; everybody loads.  72 KB of shared memory and 256 TMEM columns per CTA: two CTAs per SM.
#include "ptx_sm100.cuh"
#include "dvid_internal.h"

namespace dvid {

namespace {

constexpr int HD = 32;
constexpr int WS = 7;
constexpr int WT = WS * WS;          // 49 tokens per window
constexpr int THREADS = 160;
constexpr int OFF_Q = 0;                       // [128 rows][128 B]
constexpr int OFF_K = OFF_Q + 128 * 128;       // [128 rows][128 B]
constexpr int OFF_VT = OFF_K + 128 * 128;      // [2 windows][32 rows][128 B]
constexpr int OFF_P = OFF_VT + 2 * HD * 128;   // [2 windows][128 rows][128 B]
constexpr int OFF_BIAS = OFF_P + 2 * 128 * 128;    // [49][49] fp32 of this head
constexpr int OFF_REG = OFF_BIAS + ((WT * WT * 4 + 15) & ~15);   // [128] int: shift-mask region of every row
constexpr int OFF_BAR = OFF_REG + 128 * 4;
constexpr int SMEM_BYTES = OFF_BAR + 64 + 1024;
constexpr int COL_S = 0, COL_O = 128;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__device__ __forceinline__ void transpose8x8_h(uint32_t (&x)[4], int i) {
  {
    const bool up = (i & 4) != 0;
    const uint32_t s0 = up ? x[0] : x[2], s1 = up ? x[1] : x[3];
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
    if (up) { x[0] = r0; x[1] = r1; } else { x[2] = r0; x[3] = r1; }
  }
  {
    const bool up = (i & 2) != 0;
    const uint32_t s0 = up ? x[0] : x[1], s1 = up ? x[2] : x[3];
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (up) { x[0] = r0; x[2] = r1; } else { x[1] = r0; x[3] = r1; }
  }
  {
    const bool up = (i & 1) != 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t send = up ? (x[e] & 0xffffu) : (x[e] >> 16);
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
      x[e] = up ? ((x[e] & 0xffff0000u) | recv) : ((x[e] & 0xffffu) | (recv << 16));
    }
  }
}

__global__ void __launch_bounds__(THREADS, 2)
swin_window_attention_tc_kernel(const __half* __restrict__ qkv, const float* __restrict__ bias,
                                __half* __restrict__ out, int C, long wins, int nwy, int nwx, int Hp, int Wp, int shift,
                                float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + OFF_Q;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sVt = smem + OFF_VT;
  uint8_t* sP = smem + OFF_P;
  float* sBias = reinterpret_cast<float*>(smem + OFF_BIAS);
  int* sReg = reinterpret_cast<int*>(smem + OFF_REG);
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* p_full = s_full + 1;
  uint64_t* o_full = s_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 3);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y;
  const long win0 = 2L * blockIdx.x;
  const long ld = 3L * C;

  pdl_trigger();
  if (warp == 0 && elect_one()) {
    mbar_init(s_full, 1);
    mbar_init(p_full, 128);
    mbar_init(o_full, 1);
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // ---- Q and K tiles: row = 64 * window + token; rows of missing tokens / a missing second window are zero
  for (int id = tid; id < 128 * 4; id += THREADS) {
    const int row = id >> 2, c = id & 3;
    const int w = row >> 6, tok = row & 63;
    uint4 qv = make_uint4(0, 0, 0, 0), kv = qv;
    if (tok < WT && win0 + w < wins) {
      const __half* base = qkv + ((win0 + w) * WT + tok) * ld + head * HD + c * 8;
      qv = __ldg(reinterpret_cast<const uint4*>(base));
      kv = __ldg(reinterpret_cast<const uint4*>(base + C));
    }
    const uint32_t off = row * 128 + ((c ^ (row & 7)) << 4);
    *reinterpret_cast<uint4*>(sQ + off) = qv;
    *reinterpret_cast<uint4*>(sK + off) = kv;
  }
  // ---- V^T: item = (window, 8-key block of 8, 16-byte chunk of 4): 64 items, 4 per warp pass
  {
    const int g = lane >> 3, i = lane & 7;
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int item = (warp + p * 5) * 4 + g;            // warp-uniform validity: 64 is a multiple of 4
      const int w = item >> 5, kb8 = (item >> 2) & 7, c = item & 3;
      const int tok = kb8 * 8 + i;
      uint4 raw = make_uint4(0, 0, 0, 0);
      if (item < 64 && tok < WT && win0 + w < wins)
        raw = __ldg(reinterpret_cast<const uint4*>(qkv + ((win0 + w) * WT + tok) * ld + 2 * C + head * HD + c * 8));
      uint32_t x[4] = {raw.x, raw.y, raw.z, raw.w};
      transpose8x8_h(x, i);
      if (item < 64) {
        const int d = c * 8 + i;
        *reinterpret_cast<uint4*>(sVt + w * (HD * 128) + d * 128 + ((kb8 ^ (d & 7)) << 4)) =
            make_uint4(x[0], x[1], x[2], x[3]);
      }
    }
  }
  // ---- this head's relative position bias, and the shift-mask region of every token (BasicLayer.forward :387-406)
  for (int id = tid; id < WT * WT; id += THREADS) sBias[id] = __ldg(bias + static_cast<long>(head) * WT * WT + id);
  if (tid < 128) {
    const int w = tid >> 6, tok = tid & 63;
    int reg = 0;
    if (shift > 0 && tok < WT && win0 + w < wins) {
      const long win = win0 + w;
      const int wx = static_cast<int>(win % nwx), wy = static_cast<int>((win / nwx) % nwy);
      const int ys = wy * WS + tok / WS, xs = wx * WS + tok % WS;
      const int hr = ys < Hp - WS ? 0 : (ys < Hp - shift ? 1 : 2);
      const int wr = xs < Wp - WS ? 0 : (xs < Wp - shift ? 1 : 2);
      reg = hr * 3 + wr;
    }
    sReg[tid] = reg;
  }
  fence_proxy_async_smem();
  __syncthreads();

  if (warp == 4) {
    // ===================== MMA issuer =====================
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, 128);
      const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sQ));
      const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sK));
      umma_f16(tmem_base + COL_S, adesc, bdesc, idesc, 0u);
      umma_f16(tmem_base + COL_S, adesc + 2, bdesc + 2, idesc, 1u);
      umma_commit(s_full);
    }
    __syncwarp();
    mbar_wait(p_full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, HD);
#pragma unroll
      for (int w = 0; w < 2; ++w) {
        const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sP + w * (128 * 128)));
        const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sVt + w * (HD * 128)));
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16(tmem_base + COL_O + w * HD, adesc + 2 * kk, bdesc + 2 * kk, idesc, kk ? 1u : 0u);
      }
      umma_commit(o_full);
    }
    __syncwarp();
  } else {
    // ===================== softmax: thread = query row = TMEM lane; window w = warp / 2 =====================
    const int r = warp * 32 + lane;
    const int w = r >> 6, tok = r & 63;
    const bool row_ok = tok < WT && win0 + w < wins;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    mbar_wait(s_full, 0);
    tc_fence_after();
    uint32_t sraw[64];
    tmem_ld32(tlane + COL_S + w * 64, *reinterpret_cast<uint32_t(*)[32]>(&sraw[0]));
    tmem_ld32(tlane + COL_S + w * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&sraw[32]));
    tmem_ld_wait();
    constexpr float LOG2E = 1.4426950408889634f;
    const float* brow = sBias + (row_ok ? tok : 0) * WT;
    const int myreg = sReg[r];
    float val[WT];
    float mx = -INFINITY;
#pragma unroll
    for (int key = 0; key < WT; ++key) {
      float add = brow[key];
      if (sReg[w * 64 + key] != myreg) add += -100.0f;
      val[key] = fmaf(__uint_as_float(sraw[key]), scale_log2e, add * LOG2E);
      mx = fmaxf(mx, val[key]);
    }
    float sum = 0.f;
    uint8_t* prow = sP + w * (128 * 128) + r * 128;
#pragma unroll
    for (int c8 = 0; c8 < 8; ++c8) {
      float p[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int key = c8 * 8 + e;
        p[e] = key < WT ? ex2_approx(val[key < WT ? key : 0] - mx) : 0.f;
        sum += p[e];
      }
      uint4 pk;
      pk.x = pack_half2(p[0], p[1]);
      pk.y = pack_half2(p[2], p[3]);
      pk.z = pack_half2(p[4], p[5]);
      pk.w = pack_half2(p[6], p[7]);
      *reinterpret_cast<uint4*>(prow + ((c8 ^ (r & 7)) << 4)) = pk;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(p_full);
    mbar_wait(o_full, 0);
    tc_fence_after();
    uint32_t t[32];
    tmem_ld32(tlane + COL_O + w * HD, t);
    tmem_ld_wait();
    if (row_ok) {
      const float inv = 1.f / sum;
      __half* orow = out + ((win0 + w) * WT + tok) * C + head * HD;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 pk;
        pk.x = pack_half2(__uint_as_float(t[c * 8 + 0]) * inv, __uint_as_float(t[c * 8 + 1]) * inv);
        pk.y = pack_half2(__uint_as_float(t[c * 8 + 2]) * inv, __uint_as_float(t[c * 8 + 3]) * inv);
        pk.z = pack_half2(__uint_as_float(t[c * 8 + 4]) * inv, __uint_as_float(t[c * 8 + 5]) * inv);
        pk.w = pack_half2(__uint_as_float(t[c * 8 + 6]) * inv, __uint_as_float(t[c * 8 + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c * 8) = pk;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace

int swin_window_attention_tc_launch(const void* qkv, const float* bias, void* out, int B, int H, int W, int C,
                                    int heads, int shift, cudaStream_t stream) {
  if (B <= 0 || heads <= 0 || C != heads * HD) return DVID_ERR_SHAPE;
  const int nwy = (H + WS - 1) / WS, nwx = (W + WS - 1) / WS;
  const long wins = static_cast<long>(B) * nwy * nwx;
  if ((wins + 1) / 2 > 2147483647L || heads > 65535) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(swin_window_attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             SMEM_BYTES) != cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid(static_cast<unsigned>((wins + 1) / 2), heads);
  launch_pdl(swin_window_attention_tc_kernel, grid, dim3(THREADS), SMEM_BYTES, stream,
             static_cast<const __half*>(qkv), bias, static_cast<__half*>(out), C, wins, nwy, nwx, nwy * WS, nwx * WS,
             shift, scale_log2e);
  return check_launch();
}

}  // namespace dvid

// 256 -> 256 Linear fused with the row epilogue that follows it in the decoder: one launch instead of a tcgen05 GEMM that
// writes fp32 partials + a row kernel that reads them back.
//
//   y   = A[M x 256] . W^T[256 x 256] + bias            tcgen05.mma, M=128 rows per CTA, N=256, K=256, fp32 in TMEM
//   y  += resid                                         (fp32 [M][256], optional)
//   y   = LayerNorm(y) * gamma + beta                   (optional)
//   out_f32 = y ; out_f16 = act(y)                      act: 0 none, 1 ReLU, 2 SiLU (applied to the fp16 output only)
//
// Call sites (reference lines): self_attn.out_proj + residual + norm1 of every RCNNHead evaluation (box_head.py:516-518,
// :626-628), out_proj of the global attention followed by the SiLU that opens c_mlp (:371, :644), and c_mlp's Linear
// (:644).  In round 1 these were `gemm 2400x256->256` launches at 25 TFLOP/s (12 us, a 19-tile grid) each followed by a
// 5 us row kernel.
//
// The whole problem of a CTA fits shared memory (A tile 64 KB + the 128 KB weight matrix), so there is no ring: warp 0
// issues the eight TMA loads (the weight loads before the dependency wait - weights never depend on the previous kernel),
// warp 1 issues the 16 MMAs, warps 4..11 run the epilogue with two threads per row (column halves) straight out of TMEM,
// statistics in one pass (sum and sum of squares, exchanged between the halves through shared memory), like
// head_tail_kernel.
#include "ptx_sm100.cuh"
#include "dvid_internal.h"

namespace dvid {

namespace {

constexpr int TM = 128;
constexpr int D = 256;
constexpr int A_KB = TM * 128;            // 16 KB per 64-wide k-block
constexpr int W_KB = D * 128;             // 32 KB
constexpr int OFF_A = 0;
constexpr int OFF_W = OFF_A + 4 * A_KB;
constexpr int OFF_PAR = OFF_W + 4 * W_KB;           // bias | gamma | beta (256 floats each)
constexpr int OFF_STAT = OFF_PAR + 3 * D * 4;       // [2][128] float2
constexpr int OFF_BAR = OFF_STAT + 2 * TM * 8;
constexpr int SMEM_BYTES = OFF_BAR + 64 + 1024;
constexpr int EPI_THREADS = 256;

struct RowGemmMaps {
  CUtensorMap a;
  CUtensorMap w;
};

struct RowGemmArgs {
  int M;
  int act;                 // on the fp16 output: 0 none, 1 ReLU, 2 SiLU
  const float* bias;       // may be null
  const float* resid;      // may be null
  const float* gamma;      // null: no LayerNorm
  const float* beta;
  float* out_f32;          // may be null
  __half* out_f16;         // may be null
  const uint8_t* w_ptr;    // for the pre-wait L2 prefetch
};

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

__global__ void __launch_bounds__(384, 1)
gemm256_row_kernel(const __grid_constant__ RowGemmMaps tm, const RowGemmArgs p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem + OFF_A;
  uint8_t* sW = smem + OFF_W;
  float* sBias = reinterpret_cast<float*>(smem + OFF_PAR);
  float* sG = sBias + D;
  float* sB = sG + D;
  float2* sStat = reinterpret_cast<float2*>(smem + OFF_STAT);
  uint64_t* a_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* w_full = a_full + 1;
  uint64_t* acc_full = a_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_full + 3);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = blockIdx.x * TM;
  pdl_trigger();
  if (warp == 1 && elect_one()) {
    mbar_init(a_full, 1);
    mbar_init(w_full, 1);
    mbar_init(acc_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm.a);
    tma_prefetch_desc(&tm.w);
    mbar_expect_tx(w_full, 4 * W_KB);          // weights: independent of the previous kernel
    for (int kb = 0; kb < 4; ++kb) tma_load_2d(sW + kb * W_KB, &tm.w, w_full, kb * 64, 0);
  }
  if (warp >= 4) {                             // per-column parameters are weights too
    const int t = threadIdx.x - 128;
    sBias[t] = p.bias ? __ldg(p.bias + t) : 0.f;
    sG[t] = p.gamma ? __ldg(p.gamma + t) : 1.f;
    sB[t] = p.gamma ? __ldg(p.beta + t) : 0.f;
  }
  pdl_wait();

  if (warp == 0) {
    if (elect_one()) {
      mbar_expect_tx(a_full, 4 * A_KB);
      for (int kb = 0; kb < 4; ++kb) tma_load_2d(sA + kb * A_KB, &tm.a, a_full, kb * 64, m0);
    }
  } else if (warp == 1) {
    mbar_wait(w_full, 0);
    mbar_wait(a_full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(TM, D);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sA + kb * A_KB));
        const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sW + kb * W_KB));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(acc_full);
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue: two threads per row (column halves) =====================
    const int ew = warp & 3;                  // TMEM sub-partition
    const int ch = (warp - 4) >> 2;           // 0: columns 0..127, 1: 128..255
    const int r = ew * 32 + lane;
    const long grow = static_cast<long>(m0) + r;
    const bool valid = grow < p.M;
    const uint32_t tcol = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + ch * 128;
    const float* rrow = (p.resid && valid) ? p.resid + grow * D + ch * 128 : nullptr;
    named_bar_sync(1, EPI_THREADS);           // parameters staged
    mbar_wait(acc_full, 0);
    tc_fence_after();
    float mean = 0.f, rstd = 1.f;
    if (p.gamma != nullptr) {
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c2 = 0; c2 < 2; ++c2) {
        uint32_t v0[32], v1[32];
        tmem_ld32(tcol + c2 * 64, v0);
        tmem_ld32(tcol + c2 * 64 + 32, v1);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
          if (rrow) {
            ra = *reinterpret_cast<const float4*>(rrow + c2 * 64 + j);
            rb = *reinterpret_cast<const float4*>(rrow + c2 * 64 + 32 + j);
          }
          const float ra4[4] = {ra.x, ra.y, ra.z, ra.w}, rb4[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float a = __uint_as_float(v0[j + e]) + sBias[ch * 128 + c2 * 64 + j + e] + ra4[e];
            const float b = __uint_as_float(v1[j + e]) + sBias[ch * 128 + c2 * 64 + 32 + j + e] + rb4[e];
            sum += a + b;
            sq = fmaf(a, a, sq);
            sq = fmaf(b, b, sq);
          }
        }
      }
      sStat[ch * TM + r] = make_float2(sum, sq);
      named_bar_sync(2, EPI_THREADS);
      const float2 other = sStat[(ch ^ 1) * TM + r];
      mean = (sum + other.x) * (1.f / D);
      rstd = rsqrtf(fmaxf((sq + other.y) * (1.f / D) - mean * mean, 0.f) + 1e-5f);
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) {             // 32-column pieces of this thread's 128 columns
      uint32_t v[32];
      tmem_ld32(tcol + c * 32, v);
      tmem_ld_wait();
      if (valid) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int col = ch * 128 + c * 32 + q * 8;
          float y[8];
          float4 ra = make_float4(0.f, 0.f, 0.f, 0.f), rb = ra;
          if (rrow) {
            ra = *reinterpret_cast<const float4*>(rrow + c * 32 + q * 8);
            rb = *reinterpret_cast<const float4*>(rrow + c * 32 + q * 8 + 4);
          }
          const float rr[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float t = __uint_as_float(v[q * 8 + e]) + sBias[col + e] + rr[e];
            y[e] = (t - mean) * rstd * sG[col + e] + sB[col + e];
          }
          if (p.out_f32) {
            float* o = p.out_f32 + grow * D + col;
            *reinterpret_cast<float4*>(o) = make_float4(y[0], y[1], y[2], y[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(y[4], y[5], y[6], y[7]);
          }
          if (p.out_f16) {
            if (p.act == 1) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
            } else if (p.act == 2) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = silu_f(y[e]);
            }
            uint4 pk;
            pk.x = pack_half2(y[0], y[1]);
            pk.y = pack_half2(y[2], y[3]);
            pk.z = pack_half2(y[4], y[5]);
            pk.w = pack_half2(y[6], y[7]);
            *reinterpret_cast<uint4*>(p.out_f16 + grow * D + col) = pk;
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

}  // namespace

// a [M][256] fp16, w [256][256] fp16 ([out][in]); bias / gamma / beta fp32 [256]; resid / out_f32 fp32 [M][256];
// out_f16 fp16 [M][256].  gamma == null: no LayerNorm.
int gemm256_row_launch(const void* a, const void* w, const float* bias, const float* resid, const float* gamma,
                       const float* beta, int act, float* out_f32, void* out_f16, int M, cudaStream_t stream) {
  if (M <= 0 || act < 0 || act > 2) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(gemm256_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  RowGemmMaps tm;
  {
    const uint64_t dims[2] = {D, static_cast<uint64_t>(M)};
    const uint64_t strides[1] = {D * 2};
    const uint32_t box[2] = {64, TM};
    int r = make_tmap_f16(&tm.a, a, 2, dims, strides, box, nullptr);
    if (r) return r;
  }
  {
    const uint64_t dims[2] = {D, D};
    const uint64_t strides[1] = {D * 2};
    const uint32_t box[2] = {64, D};
    int r = make_tmap_f16(&tm.w, w, 2, dims, strides, box, nullptr);
    if (r) return r;
  }
  RowGemmArgs p;
  p.M = M; p.act = act; p.bias = bias; p.resid = resid; p.gamma = gamma; p.beta = beta;
  p.out_f32 = out_f32; p.out_f16 = static_cast<__half*>(out_f16);
  p.w_ptr = static_cast<const uint8_t*>(w);
  launch_pdl(gemm256_row_kernel, dim3((M + TM - 1) / TM), dim3(384), SMEM_BYTES, stream, tm, p);
  return check_launch();
}

}  // namespace dvid

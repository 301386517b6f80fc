// Multi-head attention core (head dim 32) on the 5th-generation tensor cores: softmax(Q K^T / sqrt(32)) V.
//
// Same contract as attention_hd32_kernel (attention.cu) - the two attentions of the decoder, torch.nn.MultiheadAttention
// in the reference: per-frame self-attention over the N=300 boxes (box_head.py:515-516) and the global cross-attention
// of 2400 queries over the 900-row video memory (box_head.py:366-371) - with both contractions as tcgen05.mma:
//
//   S   = Q[128 x 32] . K_c^T[32 x 112]      M=128, N=112, K=32  -> TMEM columns 0..111 / 112..223 (two buffers, fp32)
//   P   = exp2((S - m) * log2e/sqrt(32))     four softmax warps, one query row per thread (row == TMEM lane), the row
//                                            read with tcgen05.ld, P written to shared memory as the next A operand
//   O_c = P[128 x 112] . V_c[112 x 32]       M=128, N=32, K=112  -> TMEM columns 224..255
//   O   = O * alpha + O_c                    online-softmax rescale in registers (32 fp32 per thread)
//
// One CTA = 128 queries of one (batch, head); keys in chunks of 112 (300 keys = 3 chunks, 900 = 9).  Operand tiles are K-major with the 128-byte
// swizzle: a 32-wide head occupies the first 64 bytes of each 128-byte row (chunks 0..3 before the XOR), the MMA only
// walks K = 0..31.  V is transposed on the way into shared memory (B operand rows = head dims, K = keys; 8x8 shuffle
// transposes, 16-byte stores).  The chain of a chunk is kept short: S is double-buffered in TMEM and issued two chunks
// ahead, the K tile is triple-buffered, the softmax threads compute P(j) in registers while P(j-1).V is still in the
// tensor pipe and add O_c(j-1) afterwards (deferred rescale), and a loader warp fills the tiles of the coming chunks
// ahead of the MMA warp (ready / free mbarriers per tile, requests batched so that 8 per lane are in flight).
// Warps: 0..3 softmax, 4 MMA issue, 5 loader.  Two CTAs per SM (106 KB of shared memory, 256 TMEM columns each).
#include "ptx_sm100.cuh"
#include "dvid_internal.h"

namespace dvid {

namespace {

constexpr int HD = 32;
constexpr int BQ = 128;                  // queries per CTA
constexpr int KC = 112;                  // keys per chunk: two S buffers + O_c = 2*112 + 32 = 256 TMEM columns exactly
constexpr int SQ_BYTES = BQ * 128;       // 16 KB
constexpr int KBS = (KC + 63) / 64;      // 64-key blocks per chunk (the last one half used)
constexpr int SK_BYTES = KC * 128;       // 12 KB per buffer, three buffers
constexpr int SVT_BYTES = KBS * HD * 128;   // 8 KB per buffer: [kb][32 rows][128 B], two buffers
constexpr int SP_BYTES = KBS * BQ * 128;    // 32 KB: [kb][128 rows][128 B]
constexpr int OFF_Q = 0;
constexpr int OFF_K = OFF_Q + SQ_BYTES;
constexpr int OFF_VT = OFF_K + 3 * SK_BYTES;
constexpr int OFF_P = OFF_VT + 2 * SVT_BYTES;
constexpr int OFF_BAR = OFF_P + SP_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr int COL_S = 0, COL_O = 2 * KC;
constexpr int THREADS = 224;                  // warps 0..3 softmax, 4 MMA issue, 5 and 6 tile loaders (odd / even chunks)

// 2^x for x <= 0 (softmax arguments after the running-max shift): one MUFU, denormal results flush to zero
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// K chunk: rows = keys, 4 x 16-byte chunks of the head's 32 dims, swizzled; rows past Lk are zero.  NT threads take
// part (thread t); the loads are issued in batches of 8 before the first store so that a single warp keeps 8 requests
// per lane in flight (a load-store-load chain costs one L2 round trip per 16 bytes).
template <int NT>
__device__ __forceinline__ void load_k_chunk(uint8_t* dst, const __half* k, int key0, int Lk, long rs, int t) {
  constexpr int PER = (KC * 4 + NT - 1) / NT;
  constexpr int B = PER < 8 ? PER : 8;
#pragma unroll
  for (int i0 = 0; i0 < PER; i0 += B) {
    uint4 v[B];
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int id = t + (i0 + i) * NT;
      const int row = id >> 2, c = id & 3;
      v[i] = make_uint4(0, 0, 0, 0);
      if (id < KC * 4 && key0 + row < Lk)
        v[i] = __ldg(reinterpret_cast<const uint4*>(k + static_cast<long>(key0 + row) * rs + c * 8));
    }
#pragma unroll
    for (int i = 0; i < B; ++i) {
      const int id = t + (i0 + i) * NT;
      const int row = id >> 2, c = id & 3;
      if (id < KC * 4) *reinterpret_cast<uint4*>(dst + row * 128 + ((c ^ (row & 7)) << 4)) = v[i];
    }
  }
}

// 8x8 transpose of 16-bit elements held one row per lane (4 registers) across 8 consecutive lanes.
__device__ __forceinline__ void transpose8x8_h(uint32_t (&x)[4], int i) {
  {   // exchange 4-element halves between lanes i and i^4
    const bool up = (i & 4) != 0;
    const uint32_t s0 = up ? x[0] : x[2], s1 = up ? x[1] : x[3];
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 4), r1 = __shfl_xor_sync(0xffffffffu, s1, 4);
    if (up) { x[0] = r0; x[1] = r1; } else { x[2] = r0; x[3] = r1; }
  }
  {   // exchange 2-element quarters between lanes i and i^2
    const bool up = (i & 2) != 0;
    const uint32_t s0 = up ? x[0] : x[1], s1 = up ? x[2] : x[3];
    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, s0, 2), r1 = __shfl_xor_sync(0xffffffffu, s1, 2);
    if (up) { x[0] = r0; x[2] = r1; } else { x[1] = r0; x[3] = r1; }
  }
  {   // exchange single elements between lanes i and i^1
    const bool up = (i & 1) != 0;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const uint32_t send = up ? (x[e] & 0xffffu) : (x[e] >> 16);
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, 1);
      x[e] = up ? ((x[e] & 0xffff0000u) | recv) : ((x[e] & 0xffffu) | (recv << 16));
    }
  }
}

// V chunk transposed: element (d, key) -> k-block key/64, row d, 2-byte slot key%64 (swizzled 16-byte chunks).
// NW whole warps take part (warp w).  A group of 8 lanes loads the same 8 dims (16-byte chunk c) of 8 consecutive keys;
// the 8x8 shuffle transpose turns that into "8 consecutive keys of one dim" per lane = one 16-byte store into row d.
template <int NW>
__device__ __forceinline__ void load_vt_chunk(uint8_t* dst, const __half* v, int key0, int Lk, long rs, int w,
                                              int lane) {
  const int g = lane >> 3, i = lane & 7;            // 4 groups of 8 lanes
  constexpr int ITEMS = (KC / 8) * 4;               // (8-key block, chunk): 64 items, 4 per warp pass
  constexpr int PER = (ITEMS / 4 + NW - 1) / NW;    // passes of this warp
  constexpr int B = PER < 8 ? PER : 8;
#pragma unroll
  for (int p0 = 0; p0 < PER; p0 += B) {
    uint4 raw[B];
#pragma unroll
    for (int p = 0; p < B; ++p) {
      const int item = (w + (p0 + p) * NW) * 4 + g;
      const int key = (item >> 2) * 8 + i, c = item & 3;
      raw[p] = make_uint4(0, 0, 0, 0);
      if (item < ITEMS && key0 + key < Lk)
        raw[p] = __ldg(reinterpret_cast<const uint4*>(v + static_cast<long>(key0 + key) * rs + c * 8));
    }
#pragma unroll
    for (int p = 0; p < B; ++p) {
      const int item = (w + (p0 + p) * NW) * 4 + g;       // warp-uniform validity: ITEMS is a multiple of 4
      uint32_t x[4] = {raw[p].x, raw[p].y, raw[p].z, raw[p].w};
      transpose8x8_h(x, i);
      if (item < ITEMS) {
        const int kb8 = item >> 2, c = item & 3;
        const int d = c * 8 + i;                          // lane i now holds dim 8c+i of keys kb8*8 .. kb8*8+7
        const int kk = (kb8 * 8) & 63;
        uint8_t* blk = dst + ((kb8 * 8) >> 6) * (HD * 128);
        *reinterpret_cast<uint4*>(blk + d * 128 + (((kk >> 3) ^ (d & 7)) << 4)) = make_uint4(x[0], x[1], x[2], x[3]);
      }
    }
  }
}

// K and V^T tiles of one chunk by ONE warp with every request in flight before the first store: one L2 round trip per
// chunk (28 x 16 bytes per lane), so that the loader warp keeps ahead of the softmax chain.
__device__ __forceinline__ void load_kv_chunk_warp(uint8_t* dk, uint8_t* dvt, const __half* k, const __half* v, int key0k,
                                                   int key0v, int Lk, long k_rs, long v_rs, int lane) {
  constexpr int PK = (KC * 4) / 32;                 // 14 K chunks per lane
  constexpr int PV = (KC / 8);                      // 14 passes of 4 (8-key block, chunk) items
  const int g = lane >> 3, i = lane & 7;
  uint4 kv[PK], vv[PV];
#pragma unroll
  for (int p = 0; p < PK; ++p) {
    const int id = lane + p * 32;
    const int row = id >> 2, c = id & 3;
    kv[p] = make_uint4(0, 0, 0, 0);
    if (key0k + row < Lk) kv[p] = __ldg(reinterpret_cast<const uint4*>(k + static_cast<long>(key0k + row) * k_rs + c * 8));
  }
#pragma unroll
  for (int p = 0; p < PV; ++p) {
    const int item = p * 4 + g;
    const int key = (item >> 2) * 8 + i, c = item & 3;
    vv[p] = make_uint4(0, 0, 0, 0);
    if (key0v + key < Lk) vv[p] = __ldg(reinterpret_cast<const uint4*>(v + static_cast<long>(key0v + key) * v_rs + c * 8));
  }
#pragma unroll
  for (int p = 0; p < PK; ++p) {
    const int id = lane + p * 32;
    const int row = id >> 2, c = id & 3;
    *reinterpret_cast<uint4*>(dk + row * 128 + ((c ^ (row & 7)) << 4)) = kv[p];
  }
#pragma unroll
  for (int p = 0; p < PV; ++p) {
    const int item = p * 4 + g;
    uint32_t x[4] = {vv[p].x, vv[p].y, vv[p].z, vv[p].w};
    transpose8x8_h(x, i);
    const int kb8 = item >> 2, c = item & 3;
    const int d = c * 8 + i;
    const int kk = (kb8 * 8) & 63;
    uint8_t* blk = dvt + ((kb8 * 8) >> 6) * (HD * 128);
    *reinterpret_cast<uint4*>(blk + d * 128 + (((kk >> 3) ^ (d & 7)) << 4)) = make_uint4(x[0], x[1], x[2], x[3]);
  }
}

__global__ void __launch_bounds__(THREADS, 2)
attention_tc_kernel(const __half* __restrict__ Q, const __half* __restrict__ K, const __half* __restrict__ V,
                    __half* __restrict__ O, int Lq, int Lk, long q_rs, long k_rs, long v_rs, long o_rs, long q_bs,
                    long k_bs, long v_bs, long o_bs, float scale_log2e) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sQ = smem + OFF_Q;
  uint8_t* sK = smem + OFF_K;
  uint8_t* sVt = smem + OFF_VT;
  uint8_t* sP = smem + OFF_P;
  uint64_t* s_full = reinterpret_cast<uint64_t*>(smem + OFF_BAR);       // [2]: one per S buffer
  uint64_t* p_full = s_full + 2;
  uint64_t* o_full = s_full + 3;
  uint64_t* k_ready = s_full + 4;      // [3] loader -> MMA: K tile landed
  uint64_t* v_ready = s_full + 7;      // [2] loader -> MMA: V^T tile landed
  uint64_t* k_free = s_full + 9;       // [3] MMA -> loader: S(c) done with its K tile
  uint64_t* v_free = s_full + 12;      // [2] MMA -> loader: P(j).V done with its V^T tile
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 14);

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = blockIdx.x * BQ;
  const __half* q = Q + batch * q_bs + head * HD;
  const __half* k = K + batch * k_bs + head * HD;
  const __half* v = V + batch * v_bs + head * HD;
  __half* o = O + batch * o_bs + head * HD;
  const int nchunks = (Lk + KC - 1) / KC;

  pdl_trigger();
  if (warp == 0 && elect_one()) {
    mbar_init(&s_full[0], 1);
    mbar_init(&s_full[1], 1);
    mbar_init(p_full, BQ);
    mbar_init(o_full, 1);
    for (int i = 0; i < 3; ++i) { mbar_init(&k_ready[i], 1); mbar_init(&k_free[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&v_ready[i], 1); mbar_init(&v_free[i], 1); }
    fence_mbar_init();
  }
  if (warp == 4) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // ---- Q tile, K(0), K(1), V^T(0): everybody; rows past Lq / Lk are zero
  for (int id = tid; id < BQ * 4; id += THREADS) {
    const int row = id >> 2, c = id & 3;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (q0 + row < Lq) val = __ldg(reinterpret_cast<const uint4*>(q + static_cast<long>(q0 + row) * q_rs + c * 8));
    *reinterpret_cast<uint4*>(sQ + row * 128 + ((c ^ (row & 7)) << 4)) = val;
  }
  load_k_chunk<THREADS>(sK, k, 0, Lk, k_rs, tid);
  if (nchunks > 1) load_k_chunk<THREADS>(sK + SK_BYTES, k, KC, Lk, k_rs, tid);
  load_vt_chunk<THREADS / 32>(sVt, v, 0, Lk, v_rs, warp, lane);
  fence_proxy_async_smem();
  __syncthreads();

  if (warp >= 5) {
    // ===================== tile loaders: run ahead of the MMAs, bounded by the buffers =====================
    // warp 5 takes the odd chunks, warp 6 the even ones: a chunk costs its loader one L2 round trip + 14 shuffle
    // transposes, which one warp alone cannot hide behind the softmax of a 112-key chunk
    // K(c), c >= 2, into tile c % 3 (free when S(c-3) completed); V^T(c), c >= 1, into tile c & 1 (free when P(c-2).V did)
    // chunk c's V^T tile and chunk c+1's K tile are needed at the same time (P(c).V and S(c+1) are issued back to back),
    // so they are loaded together
    for (int c = (warp == 5 ? 1 : 2); c < nchunks; c += 2) {
      const int ck = c + 1;
      if (c >= 2) mbar_wait(&v_free[c & 1], ((c - 2) >> 1) & 1);
      if (ck < nchunks) {
        if (ck >= 3) mbar_wait(&k_free[ck % 3], ((ck - 3) / 3) & 1);
        load_kv_chunk_warp(sK + (ck % 3) * SK_BYTES, sVt + (c & 1) * SVT_BYTES, k, v, ck * KC, c * KC, Lk, k_rs, v_rs,
                           lane);
      } else {
        load_vt_chunk<1>(sVt + (c & 1) * SVT_BYTES, v, c * KC, Lk, v_rs, 0, lane);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&v_ready[c & 1]);
        if (ck < nchunks) mbar_arrive(&k_ready[ck % 3]);
      }
    }
  } else if (warp == 4) {
    // ===================== MMA issuer =====================
    const uint32_t idesc_s = umma_idesc_f16(BQ, KC);
    const uint32_t idesc_o = umma_idesc_f16(BQ, HD);
    const uint64_t qdesc = umma_desc_sw128_kmajor(smem_u32(sQ));
    auto issue_s = [&](int c) {      // S(c) = Q . K_c^T into S buffer c & 1, K tile c % 3 (one elected lane)
      const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sK + (c % 3) * SK_BYTES));
      const uint32_t d = tmem_base + COL_S + (c & 1) * KC;
      umma_f16(d, qdesc, bdesc, idesc_s, 0u);
      umma_f16(d, qdesc + 2, bdesc + 2, idesc_s, 1u);
      umma_commit(&s_full[c & 1]);
      umma_commit(&k_free[c % 3]);
    };
    tc_fence_after();
    if (elect_one()) {
      issue_s(0);
      if (nchunks > 1) issue_s(1);
    }
    __syncwarp();
    for (int j = 0; j < nchunks; ++j) {
      // V^T(j): chunk 0 came with the prologue; tile j & 1 is completed for the ((j >> 1) - (j even))-th time
      if (j >= 1) mbar_wait(&v_ready[j & 1], ((j >> 1) - ((j & 1) ? 0 : 1)) & 1);
      mbar_wait(p_full, j & 1);            // P(j) is in shared memory; S(j) and O_c(j-1) have been read
      tc_fence_after();
      if (elect_one()) {
        const uint8_t* vt = sVt + (j & 1) * SVT_BYTES;
#pragma unroll
        for (int ks = 0; ks < KC / 16; ++ks) {
          const int kb = ks >> 2, kk = ks & 3;
          const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sP + kb * (BQ * 128)));
          const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(vt + kb * (HD * 128)));
          umma_f16(tmem_base + COL_O, adesc + 2 * kk, bdesc + 2 * kk, idesc_o, ks ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(&v_free[j & 1]);
      }
      __syncwarp();
      if (j + 2 < nchunks) {
        // S(j+2): its S buffer (j & 1) has been drained (p_full(j)); K tile (j+2) % 3 from the loader
        const int c = j + 2, b = c % 3;
        mbar_wait(&k_ready[b], ((c / 3) - (b < 2 ? 1 : 0)) & 1);
        tc_fence_after();
        if (elect_one()) issue_s(c);
        __syncwarp();
      }
    }
  } else {
    // ===================== softmax warps 0..3: thread = query row = TMEM lane =====================
    const int r = warp * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float m_run = -INFINITY, l_run = 0.f, alpha_prev = 0.f;
    float oacc[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) oacc[d] = 0.f;
    for (int j = 0; j < nchunks; ++j) {
      const int valid = min(KC, Lk - j * KC);
      mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      tc_fence_after();
      // two passes over the chunk's scores in TMEM (32-column pieces): the maximum first, then the probabilities -
      // reading twice is cheaper than holding 112 scores in registers next to 56 packed probabilities
      const uint32_t scol = tlane + COL_S + (j & 1) * KC;
      float cmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < (KC + 31) / 32; ++c) {
        uint32_t t[32];
        if (c * 32 + 32 <= KC) {
          tmem_ld32(scol + c * 32, t);
        } else {
          tmem_ld16(scol + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&t[0]));
#pragma unroll
          for (int i = 16; i < 32; ++i) t[i] = 0xff800000u;
        }
        tmem_ld_wait();
        if (valid < KC) {             // last chunk only: keys past Lk never win the max
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) t[i] = 0xff800000u;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) cmax = fmaxf(cmax, __uint_as_float(t[i]));
      }
      const float m_new = fmaxf(m_run, cmax);
      const float alpha = ex2_approx((m_run - m_new) * scale_log2e);      // ex2(-inf) = 0 on the first chunk
      const float mb = m_new * scale_log2e;
      float rsum = 0.f;
      uint4 pk[KC / 8];
#pragma unroll
      for (int c = 0; c < (KC + 31) / 32; ++c) {
        uint32_t t[32];
        if (c * 32 + 32 <= KC) {
          tmem_ld32(scol + c * 32, t);
        } else {
          tmem_ld16(scol + c * 32, *reinterpret_cast<uint32_t(*)[16]>(&t[0]));
        }
        tmem_ld_wait();
        if (valid < KC) {             // ... and get probability 0
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (c * 32 + i >= valid) t[i] = 0xff800000u;
        }
#pragma unroll
        for (int q8 = 0; q8 < 4; ++q8) {
          if (c * 32 + q8 * 8 < KC) {
            float p[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              p[e] = ex2_approx(fmaf(__uint_as_float(t[q8 * 8 + e]), scale_log2e, -mb));
              rsum += p[e];
            }
            pk[c * 4 + q8].x = pack_half2(p[0], p[1]);
            pk[c * 4 + q8].y = pack_half2(p[2], p[3]);
            pk[c * 4 + q8].z = pack_half2(p[4], p[5]);
            pk[c * 4 + q8].w = pack_half2(p[6], p[7]);
          }
        }
      }
      l_run = fmaf(l_run, alpha, rsum);
      m_run = m_new;
      if (j >= 1) {
        // O_c(j-1): P(j-1).V has had the whole softmax of chunk j to finish
        mbar_wait(o_full, (j - 1) & 1);
        tc_fence_after();
        uint32_t t[32];
        tmem_ld32(tlane + COL_O, t);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < HD; ++d) oacc[d] = fmaf(oacc[d], alpha_prev, __uint_as_float(t[d]));
      }
      alpha_prev = alpha;
      // P(j) -> A operand (the tile is free: P(j-1).V completed)
#pragma unroll
      for (int c8 = 0; c8 < KC / 8; ++c8)
        *reinterpret_cast<uint4*>(sP + (c8 >> 3) * (BQ * 128) + r * 128 + (((c8 & 7) ^ (r & 7)) << 4)) = pk[c8];
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    {
      mbar_wait(o_full, (nchunks - 1) & 1);
      tc_fence_after();
      uint32_t t[32];
      tmem_ld32(tlane + COL_O, t);
      tmem_ld_wait();
#pragma unroll
      for (int d = 0; d < HD; ++d) oacc[d] = fmaf(oacc[d], alpha_prev, __uint_as_float(t[d]));
    }
    if (q0 + r < Lq) {
      const float inv = 1.f / l_run;
      __half* orow = o + static_cast<long>(q0 + r) * o_rs;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint4 pk;
        pk.x = pack_half2(oacc[c * 8 + 0] * inv, oacc[c * 8 + 1] * inv);
        pk.y = pack_half2(oacc[c * 8 + 2] * inv, oacc[c * 8 + 3] * inv);
        pk.z = pack_half2(oacc[c * 8 + 4] * inv, oacc[c * 8 + 5] * inv);
        pk.w = pack_half2(oacc[c * 8 + 6] * inv, oacc[c * 8 + 7] * inv);
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

int attention_tc_launch(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                        long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                        cudaStream_t stream) {
  if (batch <= 0 || heads <= 0 || lq <= 0 || lk <= 0) return DVID_ERR_SHAPE;
  // 16-byte vector accesses on every operand row
  if ((q_rs | k_rs | v_rs | o_rs | q_bs | k_bs | v_bs | o_bs) % 8 != 0) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(attention_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES) != cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((lq + BQ - 1) / BQ, heads, batch);
  launch_pdl(attention_tc_kernel, grid, dim3(THREADS), SMEM_BYTES, stream, static_cast<const __half*>(q),
             static_cast<const __half*>(k), static_cast<const __half*>(v), static_cast<__half*>(o), lq, lk, q_rs, k_rs,
             v_rs, o_rs, q_bs, k_bs, v_bs, o_bs, scale_log2e);
  return check_launch();
}

}  // namespace dvid

// Multi-head attention core for the DiffusionVID decoder (head dim 32): softmax(Q K^T / sqrt(32)) V, flash-style.
//
// Serves both attentions of the hot path (torch.nn.MultiheadAttention in the reference):
//   * per-frame self-attention over the N=300 boxes   (mega_core/modeling/roi_heads/box_head/box_head.py:515-516)
//   * global cross-attention, 2400 queries x 900 keys  (box_head.py:366-371)
// Projections (in_proj / out_proj) are tcgen05 GEMMs (conv_gemm.cu); this kernel consumes their fp16 outputs with
// arbitrary row/batch strides so it can read the packed [M,768] qkv buffer in place.
// One CTA = 64 queries of one (batch, head); 4 warps x 16 query rows; K/V streamed in 64-key chunks through a
// double-buffered cp.async pipeline; S and P stay in registers; fp32 online softmax (exp2 with log2e folded in).
#include "dvid_internal.h"
#include "warp_mma.cuh"

namespace dvid {

namespace {

constexpr int HD = 32;       // head dim
constexpr int BQ = 64;       // queries per CTA
constexpr int BK = 64;       // keys per chunk
constexpr int ROW_B = HD * 2;  // 64 bytes per row

// 64-byte rows: two rows share a 128-byte line; XOR the 16-byte chunk index with (row>>1)&3 so that the 8 rows of an
// ldmatrix phase hit 8 distinct bank groups.
__device__ __forceinline__ uint32_t tile_off(int row, int chunk) {
  return static_cast<uint32_t>(row * ROW_B + ((chunk ^ ((row >> 1) & 3)) << 4));
}

__device__ __forceinline__ void load_tile(uint8_t* dst, const __half* src, int row0, int nrows_valid, long row_stride,
                                          int tid) {
  // 64 rows x 4 chunks = 256 chunks, 128 threads -> 2 each
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int id = tid + i * 128;
    const int row = id >> 2, chunk = id & 3;
    const bool ok = (row0 + row) < nrows_valid;
    const __half* g = src + static_cast<long>(ok ? (row0 + row) : 0) * row_stride + chunk * 8;
    cp_async16(dst + tile_off(row, chunk), g, ok);
  }
}

__global__ void __launch_bounds__(128)
attention_hd32_kernel(const __half* __restrict__ Q, const __half* __restrict__ K, const __half* __restrict__ V,
                      __half* __restrict__ O, int Lq, int Lk, long q_rs, long k_rs, long v_rs, long o_rs, long q_bs,
                      long k_bs, long v_bs, long o_bs, float scale_log2e) {
  pdl_prologue();
  __shared__ __align__(128) uint8_t sQ[BQ * ROW_B];
  __shared__ __align__(128) uint8_t sK[2][BK * ROW_B];
  __shared__ __align__(128) uint8_t sV[2][BK * ROW_B];

  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const int head = blockIdx.y, batch = blockIdx.z;
  const int q0 = blockIdx.x * BQ;

  const __half* q = Q + batch * q_bs + head * HD;
  const __half* k = K + batch * k_bs + head * HD;
  const __half* v = V + batch * v_bs + head * HD;
  __half* o = O + batch * o_bs + head * HD;

  load_tile(sQ, q, q0, Lq, q_rs, tid);
  load_tile(sK[0], k, 0, Lk, k_rs, tid);
  load_tile(sV[0], v, 0, Lk, v_rs, tid);
  cp_async_commit();

  const int nchunks = (Lk + BK - 1) / BK;

  uint32_t qa[2][4];
  float oacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) oacc[i][j] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY};
  float l_run[2] = {0.f, 0.f};

  for (int c = 0; c < nchunks; ++c) {
    const int buf = c & 1;
    if (c + 1 < nchunks) {
      load_tile(sK[buf ^ 1], k, (c + 1) * BK, Lk, k_rs, tid);
      load_tile(sV[buf ^ 1], v, (c + 1) * BK, Lk, v_rs, tid);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();

    if (c == 0) {
      // A fragments of this warp's 16 query rows, both 16-wide k-steps of the 32-dim head
      const int j = lane >> 3, r = lane & 7;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const int row = warp * 16 + (j & 1) * 8 + r;
        const int chunk = ks * 2 + (j >> 1);
        ldmatrix_x4(qa[ks], smem_addr(sQ + tile_off(row, chunk)));
      }
    }

    // S = Q K^T for 16 x 64
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) s[i][jj] = 0.f;
    {
      const int j = lane >> 3, r = lane & 7;
#pragma unroll
      for (int np = 0; np < 4; ++np) {      // pairs of 8-key tiles
#pragma unroll
        for (int ks = 0; ks < 2; ++ks) {
          uint32_t b[4];
          const int row = np * 16 + (j >> 1) * 8 + r;  // key
          const int chunk = ks * 2 + (j & 1);          // dims
          ldmatrix_x4(b, smem_addr(sK[buf] + tile_off(row, chunk)));
          mma_16816(s[np * 2], qa[ks], b[0], b[1]);
          mma_16816(s[np * 2 + 1], qa[ks], b[2], b[3]);
        }
      }
    }

    // scale, mask, online softmax
    const int key_base = c * BK;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int key = key_base + i * 8 + 2 * t + (jj & 1);
        float val = s[i][jj] * scale_log2e;
        val = key < Lk ? val : -INFINITY;
        s[i][jj] = val;
        mx[jj >> 1] = fmaxf(mx[jj >> 1], val);
      }
    }
    float corr[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = quad_max(mx[h]);
      const float m_new = fmaxf(m_run[h], mx[h]);
      corr[h] = exp2f(m_run[h] - m_new);
      m_run[h] = m_new;
    }
    float rs[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const float p = exp2f(s[i][jj] - m_run[jj >> 1]);
        s[i][jj] = p;
        rs[jj >> 1] += p;
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      rs[h] = quad_sum(rs[h]);
      l_run[h] = l_run[h] * corr[h] + rs[h];
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      oacc[i][0] *= corr[0];
      oacc[i][1] *= corr[0];
      oacc[i][2] *= corr[1];
      oacc[i][3] *= corr[1];
    }

    // O += P V
    {
      const int j = lane >> 3, r = lane & 7;
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {   // 16 keys per step
        uint32_t pa[4];
        pa[0] = pack2h(s[2 * kk][0], s[2 * kk][1]);
        pa[1] = pack2h(s[2 * kk][2], s[2 * kk][3]);
        pa[2] = pack2h(s[2 * kk + 1][0], s[2 * kk + 1][1]);
        pa[3] = pack2h(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
        for (int np = 0; np < 2; ++np) {  // pairs of 8-dim tiles
          uint32_t b[4];
          const int row = kk * 16 + (j & 1) * 8 + r;  // key
          const int chunk = np * 2 + (j >> 1);        // dims
          ldmatrix_x4_trans(b, smem_addr(sV[buf] + tile_off(row, chunk)));
          mma_16816(oacc[np * 2], pa, b[0], b[1]);
          mma_16816(oacc[np * 2 + 1], pa, b[2], b[3]);
        }
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }

  // epilogue
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = q0 + warp * 16 + g + h * 8;
    if (row < Lq) {
      const float inv = 1.f / l_run[h];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t pk = pack2h(oacc[i][2 * h] * inv, oacc[i][2 * h + 1] * inv);
        *reinterpret_cast<uint32_t*>(o + static_cast<long>(row) * o_rs + i * 8 + 2 * t) = pk;
      }
    }
  }
}

}  // namespace

int attention_launch(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                     long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                     cudaStream_t stream) {
  if (batch <= 0 || heads <= 0 || lq <= 0 || lk <= 0) return DVID_ERR_SHAPE;
  if ((q_rs | k_rs | v_rs | q_bs | k_bs | v_bs) % 8 != 0 || (o_rs | o_bs) % 2 != 0) return DVID_ERR_SHAPE;
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid((lq + BQ - 1) / BQ, heads, batch);
  launch_pdl(attention_hd32_kernel, dim3(grid), dim3(128), 0, stream, 
      static_cast<const __half*>(q), static_cast<const __half*>(k), static_cast<const __half*>(v),
      static_cast<__half*>(o), lq, lk, q_rs, k_rs, v_rs, o_rs, q_bs, k_bs, v_bs, o_bs, scale_log2e);
  return check_launch();
}

}  // namespace dvid

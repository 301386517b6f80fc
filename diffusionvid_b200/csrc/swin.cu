// Swin Transformer backbone kernels (config vid_Swin_B_DiffusionVID.yaml): everything between the tcgen05 GEMMs.
//
// Reference: mega_core/modeling/backbone/swintransformer.py.  The reference materialises, per block, F.pad, two
// torch.roll, window_partition / window_reverse permute+contiguous copies and separate LayerNorm / add kernels
// (:220-276).  Here the residual stream X stays fp32 in token order [B,H,W,C] and ONE row kernel does
//   v = X[token] (+ GEMM output, read in token order or through the window-reverse + un-shift mapping)
//   X[token] = v                     (optional)
//   out = LayerNorm(v) as fp16       written in token order or directly in shifted-window order with zero pad rows
// so padding, cyclic shift, partition, reverse and crop are index arithmetic on loads/stores of a kernel that has to
// run anyway.  Window attention (49 tokens, head dim 32, relative-position bias, shift mask computed from the token
// coordinates instead of a materialised (nW,49,49) tensor, :387-406) is one CTA per (window, head) on mma.sync tiles.
#include "dvid_internal.h"
#include "warp_mma.cuh"

namespace dvid {

namespace {

constexpr int WS = 7;          // window size
constexpr int WT = WS * WS;    // tokens per window

struct SwinGeom {
  int B, H, W, C;
  int Hp, Wp;        // padded to multiples of 7
  int nwy, nwx;      // windows per frame
  int shift;         // 0 or 3
};

// windowed row -> token (b,y,x); returns false for a zero-pad token
// (row counts stay below 2^31 - checked by the launchers - so all index arithmetic is 32-bit: a 64-bit division by a
// run-time value is a ~100-instruction sequence and the C = 128 rows were instruction-bound on six of them per row)
__device__ __forceinline__ bool win_row_to_token(const SwinGeom& g, int r, int& b, int& y, int& x) {
  const int t = r % WT;
  int win = r / WT;
  const int wx = win % g.nwx;
  win /= g.nwx;
  const int wy = win % g.nwy;
  b = win / g.nwy;
  int ys = wy * WS + t / WS, xs = wx * WS + t % WS;      // coordinates in the shifted, padded grid
  y = ys + g.shift; if (y >= g.Hp) y -= g.Hp;             // roll(-shift): shifted[ys] = x[(ys + shift) mod Hp]
  x = xs + g.shift; if (x >= g.Wp) x -= g.Wp;
  return y < g.H && x < g.W;
}
__device__ __forceinline__ long token_to_win_row(const SwinGeom& g, int b, int y, int x) {
  int ys = y - g.shift; if (ys < 0) ys += g.Hp;
  int xs = x - g.shift; if (xs < 0) xs += g.Wp;
  const int win = (b * g.nwy + ys / WS) * g.nwx + xs / WS;
  return static_cast<long>(win) * WT + (ys % WS) * WS + xs % WS;
}

struct RowArgs {
  SwinGeom g;
  float* X;                 // fp32 residual stream [B,H,W,C] (may be null when add supplies the value)
  int write_x;
  const __half* add;        // fp16 GEMM output or null
  int add_mode;             // 0 none, 1 token order, 2 windowed order (window-reverse + un-shift + crop)
  const float* gamma; const float* beta;   // LayerNorm (null: no norm)
  __half* out16;            // fp16 output or null
  float* out32;             // fp32 output (token order) or null
  int out_mode;             // 1 token order, 2 windowed order (pad + shift + partition)
  long rows;                // iteration space: windowed rows if out_mode == 2 else tokens
};

// One warp per row.  V4 = C / 128 float4 groups per lane, group j of lane l covers channels (j*32 + l)*4 .. +3.
template <int V4>
__global__ void __launch_bounds__(256) swin_rows_kernel(const RowArgs a) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int r = static_cast<int>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= a.rows) return;
  const SwinGeom& g = a.g;
  const int C = g.C;
  int b, y, x;
  long tok;
  if (a.out_mode == 2) {
    if (!win_row_to_token(g, r, b, y, x)) {           // zero pad token (the reference pads after norm1, :236-240)
      uint2* o = reinterpret_cast<uint2*>(a.out16 + static_cast<long>(r) * C);
#pragma unroll
      for (int j = 0; j < V4; ++j) o[j * 32 + lane] = make_uint2(0u, 0u);
      return;
    }
    tok = (static_cast<long>(b) * g.H + y) * g.W + x;
  } else {
    tok = r;
    const int ry = r / g.W;
    x = r - ry * g.W;
    b = ry / g.H;
    y = ry - b * g.H;
  }
  float v[V4 * 4];
  if (a.X != nullptr) {
    const float4* xp = reinterpret_cast<const float4*>(a.X + tok * C);
#pragma unroll
    for (int j = 0; j < V4; ++j) {
      const float4 f = xp[j * 32 + lane];
      v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
    }
  } else {
#pragma unroll
    for (int e = 0; e < V4 * 4; ++e) v[e] = 0.f;
  }
  if (a.add_mode != 0) {
    const long ar = (a.add_mode == 1) ? tok : token_to_win_row(g, b, y, x);
    const uint2* ap = reinterpret_cast<const uint2*>(a.add + ar * C);
#pragma unroll
    for (int j = 0; j < V4; ++j) {
      const uint2 u = ap[j * 32 + lane];
      const float2 f0 = __half22float2(*reinterpret_cast<const __half2*>(&u.x));
      const float2 f1 = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
      v[4 * j] += f0.x; v[4 * j + 1] += f0.y; v[4 * j + 2] += f1.x; v[4 * j + 3] += f1.y;
    }
  }
  if (a.write_x) {
    float4* xp = reinterpret_cast<float4*>(a.X + tok * C);
#pragma unroll
    for (int j = 0; j < V4; ++j) xp[j * 32 + lane] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (a.out16 == nullptr && a.out32 == nullptr) return;
  if (a.gamma != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < V4 * 4; ++e) s += v[e];
    const float mean = warp_sum(s) / static_cast<float>(C);
    float q = 0.f;
#pragma unroll
    for (int e = 0; e < V4 * 4; ++e) {
      const float d = v[e] - mean;
      q += d * d;
    }
    const float rstd = rsqrtf(warp_sum(q) / static_cast<float>(C) + 1e-5f);
    const float4* gp = reinterpret_cast<const float4*>(a.gamma);
    const float4* bp = reinterpret_cast<const float4*>(a.beta);
#pragma unroll
    for (int j = 0; j < V4; ++j) {
      const float4 gg = __ldg(gp + j * 32 + lane), bb = __ldg(bp + j * 32 + lane);
      v[4 * j] = (v[4 * j] - mean) * rstd * gg.x + bb.x;
      v[4 * j + 1] = (v[4 * j + 1] - mean) * rstd * gg.y + bb.y;
      v[4 * j + 2] = (v[4 * j + 2] - mean) * rstd * gg.z + bb.z;
      v[4 * j + 3] = (v[4 * j + 3] - mean) * rstd * gg.w + bb.w;
    }
  }
  if (a.out32 != nullptr) {
    float4* op = reinterpret_cast<float4*>(a.out32 + tok * C);
#pragma unroll
    for (int j = 0; j < V4; ++j) op[j * 32 + lane] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  }
  if (a.out16 != nullptr) {
    const long orow = (a.out_mode == 2) ? r : tok;
    uint2* op = reinterpret_cast<uint2*>(a.out16 + orow * C);
#pragma unroll
    for (int j = 0; j < V4; ++j)
      op[j * 32 + lane] = make_uint2(pack2h(v[4 * j], v[4 * j + 1]), pack2h(v[4 * j + 2], v[4 * j + 3]));
  }
}

// PatchMerging (:279-317): out[b,y2,x2,:] = LayerNorm_{4C}(cat(X[2y2,2x2], X[2y2+1,2x2], X[2y2,2x2+1], X[2y2+1,2x2+1]))
// with zero padding for odd H / W.  One warp per output token; V4 = 4C / 128.
template <int V4>
__global__ void __launch_bounds__(256)
swin_merge_kernel(const float* __restrict__ X, int B, int H, int W, int C, const float* __restrict__ gamma,
                  const float* __restrict__ beta, __half* __restrict__ out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int H2 = (H + 1) / 2, W2 = (W + 1) / 2;
  const int r = static_cast<int>(blockIdx.x) * 8 + (threadIdx.x >> 5);
  if (r >= B * H2 * W2) return;
  const int ry = r / W2;
  const int x2 = r - ry * W2;
  const int b = ry / H2;
  const int y2 = ry - b * H2;
  const int C4 = 4 * C;
  float v[V4 * 4];
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const int ch = (j * 32 + lane) * 4;       // channel in the 4C concat
    const int q = ch / C, c = ch - q * C;     // source quadrant: 0 (0,0), 1 (1,0), 2 (0,1), 3 (1,1) as (dy,dx)
    const int y = 2 * y2 + (q & 1), x = 2 * x2 + (q >> 1);
    float4 f = make_float4(0.f, 0.f, 0.f, 0.f);
    if (y < H && x < W) f = *reinterpret_cast<const float4*>(X + ((static_cast<long>(b) * H + y) * W + x) * C + c);
    v[4 * j] = f.x; v[4 * j + 1] = f.y; v[4 * j + 2] = f.z; v[4 * j + 3] = f.w;
  }
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < V4 * 4; ++e) s += v[e];
  const float mean = warp_sum(s) / static_cast<float>(C4);
  float qq = 0.f;
#pragma unroll
  for (int e = 0; e < V4 * 4; ++e) {
    const float d = v[e] - mean;
    qq += d * d;
  }
  const float rstd = rsqrtf(warp_sum(qq) / static_cast<float>(C4) + 1e-5f);
  uint2* op = reinterpret_cast<uint2*>(out + static_cast<long>(r) * C4);
#pragma unroll
  for (int j = 0; j < V4; ++j) {
    const float4 gg = __ldg(reinterpret_cast<const float4*>(gamma) + j * 32 + lane);
    const float4 bb = __ldg(reinterpret_cast<const float4*>(beta) + j * 32 + lane);
    op[j * 32 + lane] = make_uint2(pack2h((v[4 * j] - mean) * rstd * gg.x + bb.x, (v[4 * j + 1] - mean) * rstd * gg.y + bb.y),
                                   pack2h((v[4 * j + 2] - mean) * rstd * gg.z + bb.z, (v[4 * j + 3] - mean) * rstd * gg.w + bb.w));
  }
}

// normalizer (diffusion_det.py:301-303) + 4x4/4 patch extraction (PatchEmbed, :422-461): img [B,3,H,W] fp32 ->
// out [B*(H/4)*(W/4)][64] fp16, k = c*16 + py*4 + px (the flattening of proj.weight[embed][3][4][4]), k >= 48 zero.
// T = uint8_t: 8-bit frames with the reference's ToTensor (u8 / 255, transforms.py:295-297) evaluated here.
__device__ __forceinline__ float4 load_px4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load_px4(const uint8_t* p) {
  const uchar4 u = *reinterpret_cast<const uchar4*>(p);
  return make_float4(__fdiv_rn(static_cast<float>(u.x), 255.f), __fdiv_rn(static_cast<float>(u.y), 255.f),
                     __fdiv_rn(static_cast<float>(u.z), 255.f), __fdiv_rn(static_cast<float>(u.w), 255.f));
}

template <typename T>
__global__ void swin_patch_gather_kernel(const T* __restrict__ img, __half* __restrict__ out, int B, int H, int W,
                                         float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_prologue();
  const int H4 = H / 4, W4 = W / 4;
  const long total = static_cast<long>(B) * H4 * W4 * 4;     // 4 threads per token: thread q writes k in [16q, 16q+16)
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int q = static_cast<int>(i & 3);
  const long tok = i >> 2;
  const int tk = static_cast<int>(tok);
  const int ty4 = tk / W4;
  const int x4 = tk - ty4 * W4;
  const int b = ty4 / H4;
  const int y4 = ty4 - b * H4;
  uint4 o[2] = {make_uint4(0, 0, 0, 0), make_uint4(0, 0, 0, 0)};
  if (q < 3) {
    const float mean = q == 0 ? m0 : (q == 1 ? m1 : m2);
    const float sd = q == 0 ? s0 : (q == 1 ? s1 : s2);
    const T* p = img + ((static_cast<long>(b) * 3 + q) * H + y4 * 4) * W + x4 * 4;
    uint32_t w[8];
#pragma unroll
    for (int py = 0; py < 4; ++py) {
      const float4 f = load_px4(p + static_cast<long>(py) * W);
      w[2 * py] = pack2h(__fdiv_rn(__fsub_rn(f.x, mean), sd), __fdiv_rn(__fsub_rn(f.y, mean), sd));
      w[2 * py + 1] = pack2h(__fdiv_rn(__fsub_rn(f.z, mean), sd), __fdiv_rn(__fsub_rn(f.w, mean), sd));
    }
    o[0] = make_uint4(w[0], w[1], w[2], w[3]);
    o[1] = make_uint4(w[4], w[5], w[6], w[7]);
  }
  uint4* op = reinterpret_cast<uint4*>(out + tok * 64 + q * 16);
  op[0] = o[0];
  op[1] = o[1];
}

// ------------------------------------------------------------------------------------------------ window attention
constexpr int HD = 32;
constexpr int ROW_B = HD * 2;
__device__ __forceinline__ uint32_t wt_off(int row, int chunk) {
  return static_cast<uint32_t>(row * ROW_B + ((chunk ^ ((row >> 1) & 3)) << 4));
}

// qkv [rows][3C] fp16 (q | k | v, head h = columns 32h..32h+31 of each third), rows = windows * 49 in windowed order.
// bias [heads][49][49] fp32 = relative_position_bias_table gathered by relative_position_index (:160-163).
// out [rows][C] fp16.  grid (windows, heads), 128 threads: warp w owns query rows 16w..16w+15 (49 valid of 64).
__global__ void __launch_bounds__(128)
swin_window_attention_kernel(const __half* __restrict__ qkv, const float* __restrict__ bias, __half* __restrict__ out,
                             int C, int nwy, int nwx, int Hp, int Wp, int shift, float scale_log2e) {
  pdl_prologue();
  __shared__ __align__(128) uint8_t sQ[64 * ROW_B];
  __shared__ __align__(128) uint8_t sK[64 * ROW_B];
  __shared__ __align__(128) uint8_t sV[64 * ROW_B];
  __shared__ int sReg[64];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  const long win = blockIdx.x;
  const int head = blockIdx.y;
  const long row0 = win * WT;
  const long ld = 3L * C;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int id = tid + i * 128;          // 64 rows x 4 chunks
    const int row = id >> 2, chunk = id & 3;
    const bool ok = row < WT;
    const __half* base = qkv + (row0 + (ok ? row : 0)) * ld + head * HD + chunk * 8;
    cp_async16(sQ + wt_off(row, chunk), base, ok);
    cp_async16(sK + wt_off(row, chunk), base + C, ok);
    cp_async16(sV + wt_off(row, chunk), base + 2 * C, ok);
  }
  cp_async_commit();
  if (tid < 64) {
    int reg = 0;
    if (shift > 0 && tid < WT) {
      const int wx = static_cast<int>(win % nwx), wy = static_cast<int>((win / nwx) % nwy);
      const int ys = wy * WS + tid / WS, xs = wx * WS + tid % WS;
      const int hr = ys < Hp - WS ? 0 : (ys < Hp - shift ? 1 : 2);
      const int wr = xs < Wp - WS ? 0 : (xs < Wp - shift ? 1 : 2);
      reg = hr * 3 + wr;
    }
    sReg[tid] = reg;
  }
  cp_async_wait<0>();
  __syncthreads();

  uint32_t qa[2][4];
  {
    const int j = lane >> 3, r = lane & 7;
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
      ldmatrix_x4(qa[ks], smem_addr(sQ + wt_off(warp * 16 + (j & 1) * 8 + r, ks * 2 + (j >> 1))));
  }
  float s[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) s[i][jj] = 0.f;
  {
    const int j = lane >> 3, r = lane & 7;
#pragma unroll
    for (int np = 0; np < 4; ++np) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t b[4];
        ldmatrix_x4(b, smem_addr(sK + wt_off(np * 16 + (j >> 1) * 8 + r, ks * 2 + (j & 1))));
        mma_16816(s[np * 2], qa[ks], b[0], b[1]);
        mma_16816(s[np * 2 + 1], qa[ks], b[2], b[3]);
      }
    }
  }
  // (q*scale) k^T + bias + mask, softmax over the 49 real keys; everything in log2 units for exp2f
  constexpr float LOG2E = 1.4426950408889634f;
  const float* bh = bias + static_cast<long>(head) * WT * WT;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int row = warp * 16 + g + 8 * (jj >> 1);
      const int key = i * 8 + 2 * t + (jj & 1);
      float val = -INFINITY;
      if (key < WT) {
        float add = 0.f;
        if (row < WT) {
          add = __ldg(bh + row * WT + key);
          if (sReg[row] != sReg[key]) add += -100.0f;
        }
        val = s[i][jj] * scale_log2e + add * LOG2E;
      }
      s[i][jj] = val;
      mx[jj >> 1] = fmaxf(mx[jj >> 1], val);
    }
  }
  float sum[2] = {0.f, 0.f};
#pragma unroll
  for (int h = 0; h < 2; ++h) mx[h] = quad_max(mx[h]);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const float p = exp2f(s[i][jj] - mx[jj >> 1]);
      s[i][jj] = p;
      sum[jj >> 1] += p;
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) sum[h] = quad_sum(sum[h]);
  float oacc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) oacc[i][jj] = 0.f;
  {
    const int j = lane >> 3, r = lane & 7;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      uint32_t pa[4];
      pa[0] = pack2h(s[2 * kk][0], s[2 * kk][1]);
      pa[1] = pack2h(s[2 * kk][2], s[2 * kk][3]);
      pa[2] = pack2h(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pa[3] = pack2h(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        uint32_t b[4];
        ldmatrix_x4_trans(b, smem_addr(sV + wt_off(kk * 16 + (j & 1) * 8 + r, np * 2 + (j >> 1))));
        mma_16816(oacc[np * 2], pa, b[0], b[1]);
        mma_16816(oacc[np * 2 + 1], pa, b[2], b[3]);
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int row = warp * 16 + g + h * 8;
    if (row < WT) {
      const float inv = 1.f / sum[h];
      __half* o = out + (row0 + row) * C + head * HD;
#pragma unroll
      for (int i = 0; i < 4; ++i)
        *reinterpret_cast<uint32_t*>(o + i * 8 + 2 * t) = pack2h(oacc[i][2 * h] * inv, oacc[i][2 * h + 1] * inv);
    }
  }
}

}  // namespace

int swin_rows_launch(float* X, int write_x, const void* add, int add_mode, const float* gamma, const float* beta,
                     void* out16, float* out32, int out_mode, int B, int H, int W, int C, int shift,
                     cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C % 128 != 0 || C > 2048) return DVID_ERR_SHAPE;
  if (out_mode != 1 && out_mode != 2) return DVID_ERR_ARG;
  if (out_mode == 2 && (out16 == nullptr || out32 != nullptr)) return DVID_ERR_ARG;
  if ((add_mode != 0) != (add != nullptr)) return DVID_ERR_ARG;
  if (X == nullptr && add == nullptr) return DVID_ERR_ARG;
  RowArgs a;
  a.g.B = B; a.g.H = H; a.g.W = W; a.g.C = C;
  a.g.nwy = (H + WS - 1) / WS; a.g.nwx = (W + WS - 1) / WS;
  a.g.Hp = a.g.nwy * WS; a.g.Wp = a.g.nwx * WS;
  a.g.shift = shift;
  a.X = X; a.write_x = write_x; a.add = static_cast<const __half*>(add); a.add_mode = add_mode;
  a.gamma = gamma; a.beta = beta; a.out16 = static_cast<__half*>(out16); a.out32 = out32; a.out_mode = out_mode;
  a.rows = out_mode == 2 ? static_cast<long>(B) * a.g.nwy * a.g.nwx * WT : static_cast<long>(B) * H * W;
  if (a.rows >= (1L << 31) - 8) return DVID_ERR_SHAPE;      // 32-bit row arithmetic in the kernel
  const unsigned grid = static_cast<unsigned>((a.rows + 7) / 8);
  switch (C / 128) {
    case 1: launch_pdl(swin_rows_kernel<1>, dim3(grid), dim3(256), 0, stream, a); break;
    case 2: launch_pdl(swin_rows_kernel<2>, dim3(grid), dim3(256), 0, stream, a); break;
    case 4: launch_pdl(swin_rows_kernel<4>, dim3(grid), dim3(256), 0, stream, a); break;
    case 8: launch_pdl(swin_rows_kernel<8>, dim3(grid), dim3(256), 0, stream, a); break;
    case 16: launch_pdl(swin_rows_kernel<16>, dim3(grid), dim3(256), 0, stream, a); break;
    default: return DVID_ERR_SHAPE;
  }
  return check_launch();
}

int swin_merge_launch(const float* X, int B, int H, int W, int C, const float* gamma, const float* beta, void* out,
                      cudaStream_t stream) {
  if (B <= 0 || H <= 0 || W <= 0 || C % 128 != 0) return DVID_ERR_SHAPE;
  const long rows = static_cast<long>(B) * ((H + 1) / 2) * ((W + 1) / 2);
  if (rows >= (1L << 31) - 8) return DVID_ERR_SHAPE;
  const unsigned grid = static_cast<unsigned>((rows + 7) / 8);
  __half* o = static_cast<__half*>(out);
  switch (4 * C / 128) {
    case 4: launch_pdl(swin_merge_kernel<4>, dim3(grid), dim3(256), 0, stream, X, B, H, W, C, gamma, beta, o); break;
    case 8: launch_pdl(swin_merge_kernel<8>, dim3(grid), dim3(256), 0, stream, X, B, H, W, C, gamma, beta, o); break;
    case 16: launch_pdl(swin_merge_kernel<16>, dim3(grid), dim3(256), 0, stream, X, B, H, W, C, gamma, beta, o); break;
    default: return DVID_ERR_SHAPE;
  }
  return check_launch();
}

int swin_patch_gather_launch(const void* img, int is_u8, void* out, int B, int H, int W, const float* mean,
                             const float* std, cudaStream_t stream) {
  if (B <= 0 || H % 4 != 0 || W % 4 != 0) return DVID_ERR_SHAPE;
  const long total = static_cast<long>(B) * (H / 4) * (W / 4) * 4;
  if (total >= (1L << 31)) return DVID_ERR_SHAPE;
  const dim3 grid(static_cast<unsigned>((total + 255) / 256));
  if (is_u8)
    launch_pdl(swin_patch_gather_kernel<uint8_t>, grid, dim3(256), 0, stream, static_cast<const uint8_t*>(img),
               static_cast<__half*>(out), B, H, W, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
  else
    launch_pdl(swin_patch_gather_kernel<float>, grid, dim3(256), 0, stream, static_cast<const float*>(img),
               static_cast<__half*>(out), B, H, W, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
  return check_launch();
}

int swin_window_attention_launch(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads,
                                 int shift, cudaStream_t stream) {
  if (B <= 0 || heads <= 0 || C != heads * HD) return DVID_ERR_SHAPE;
  const int nwy = (H + WS - 1) / WS, nwx = (W + WS - 1) / WS;
  const long wins = static_cast<long>(B) * nwy * nwx;
  if (wins > 2147483647L || heads > 65535) return DVID_ERR_SHAPE;
  const float scale_log2e = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  dim3 grid(static_cast<unsigned>(wins), heads);
  launch_pdl(swin_window_attention_kernel, dim3(grid), dim3(128), 0, stream, static_cast<const __half*>(qkv), bias,
                                                         static_cast<__half*>(out), C, nwy, nwx, nwy * WS, nwx * WS,
                                                         shift, scale_log2e);
  return check_launch();
}

}  // namespace dvid

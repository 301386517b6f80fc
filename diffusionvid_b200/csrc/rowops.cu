// Memory-bound row / pixel kernels of the DiffusionVID hot path: input normalisation + layout, stem max-pool, the fused
// "GEMM epilogue" row kernel (split-K reduce + bias + LayerNorm + ReLU + residual + LayerNorm + time/cond modulation),
// the tiny time-embedding linears, box decode (apply_deltas) and the DDIM update with box renewal.
// All are vectorised (16-byte accesses), one warp per 256-wide row where a row reduction is needed.
#include "dvid_internal.h"
#include "warp_mma.cuh"

namespace dvid {

namespace {

// ------------------------------------------------------------------------------------------------ preprocess
// normalizer of mega_core/modeling/detector/diffusion_det.py:301-303,422: (x - mean/255) / (std/255), then NCHW fp32 ->
// NHWC fp16 with 8 channels (3 real + 5 zero) and a zero halo of `halo` pixels (the stem convolution's padding).
//
// T = uint8_t is the clip-loader variant: the frame arrives as the decoded 8-bit image and the
// reference's ToTensor (mega_core/data/transforms/transforms.py:295-297, x = u8 / 255 in fp32) is evaluated here, so
// the result is bit-identical to the fp32 entry while a quarter of the bytes cross PCIe and HBM.
__device__ __forceinline__ float pixel_value(float v) { return v; }
__device__ __forceinline__ float pixel_value(uint8_t v) { return __fdiv_rn(static_cast<float>(v), 255.f); }

template <typename T>
__global__ void preprocess_kernel(const T* __restrict__ img, __half* __restrict__ out, int n, int H, int W, int halo,
                                  int Hp, int Wp, float m0, float m1, float m2, float s0, float s1, float s2) {
  pdl_prologue();
  // grid (x blocks, padded row, image): no index arithmetic beyond adds (the first version decoded a flat 64-bit index
  // with three 64-bit divisions per pixel and was instruction-bound at 2.3x its HBM time)
  const int xp = blockIdx.x * blockDim.x + threadIdx.x;
  if (xp >= Wp) return;
  const int yp = blockIdx.y, im = blockIdx.z;
  const int x = xp - halo, y = yp - halo;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (x >= 0 && x < W && y >= 0 && y < H) {
    const long plane = static_cast<long>(H) * W;
    const T* p = img + static_cast<long>(im) * 3 * plane + static_cast<long>(y) * W + x;
    const float r = __fdiv_rn(__fsub_rn(pixel_value(__ldg(p)), m0), s0);
    const float g = __fdiv_rn(__fsub_rn(pixel_value(__ldg(p + plane)), m1), s1);
    const float b = __fdiv_rn(__fsub_rn(pixel_value(__ldg(p + 2 * plane)), m2), s2);
    v.x = pack2h(r, g);
    v.y = pack2h(b, 0.f);
  }
  *reinterpret_cast<uint4*>(out + ((static_cast<long>(im) * Hp + yp) * Wp + xp) * 8) = v;
}

// ------------------------------------------------------------------------------------------------ maxpool 3x3 s2 p1
__global__ void maxpool3x3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, int n, int H, int W, int C,
                                    int Ho, int Wo) {
  pdl_prologue();
  // grid (blocks over (x, 8-channel group), output row, image)
  const int cg = C / 8;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= Wo * cg) return;
  const int xo = idx / cg, c = idx - xo * cg;
  const int yo = blockIdx.y, im = blockIdx.z;
  __half2 m[4];
  const __half2 ninf = __float2half2_rn(-65504.f);
#pragma unroll
  for (int e = 0; e < 4; ++e) m[e] = ninf;
#pragma unroll
  for (int dy = 0; dy < 3; ++dy) {
    const int y = yo * 2 - 1 + dy;
    if (y < 0 || y >= H) continue;
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int x = xo * 2 - 1 + dx;
      if (x < 0 || x >= W) continue;
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long>(im) * H + y) * W + x) * C + c * 8));
      const __half2* hp = reinterpret_cast<const __half2*>(&v);
#pragma unroll
      for (int e = 0; e < 4; ++e) m[e] = __hmax2(m[e], hp[e]);
    }
  }
  *reinterpret_cast<uint4*>(out + ((static_cast<long>(im) * Ho + yo) * Wo + xo) * C + c * 8) =
      *reinterpret_cast<uint4*>(m);
}

// ------------------------------------------------------------------------------------------------ row_post (D = 256)
struct RowPostArgs {
  const float* partials;   // [splits][M][256] fp32 (GEMM split-K partial sums), or nullptr
  int splits;
  long split_stride;       // elements between splits (M*256)
  const __half* in_f16;    // alternative fp16 input [M][256] (used when partials == nullptr)
  const float* bias;       // [256] or nullptr
  const float* ln1_g; const float* ln1_b;   // optional LayerNorm #1
  int relu1;
  const float* resid;      // fp32 [M][256] or nullptr (added after LN1/ReLU1)
  const float* ln2_g; const float* ln2_b;   // optional LayerNorm #2
  int act2;                // 0 none, 1 ReLU, 2 SiLU (applied to the fp16 output only when act2_f16_only)
  int act2_f16_only;
  float* out_f32;          // optional
  __half* out_f16;         // optional
  // optional modulation fc = y * (scale + 1) + shift written to out_mod_f16 (box_head.py:533-536, :643-647)
  const float* mod_scale;  // [groups][scale_stride]
  const float* mod_shift;  // per group ([groups][shift_stride]) or per row ([M][256]) when shift_per_row
  int rows_per_group; int scale_stride; int shift_stride; int shift_per_row;
  __half* out_mod_f16;
  int M;
};

__device__ __forceinline__ void ln8(float (&v)[8], const float* __restrict__ g, const float* __restrict__ b, int lane) {
  float s = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) s += v[e];
  const float mean = warp_sum(s) * (1.f / 256.f);
  float q = 0.f;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const float d = v[e] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / 256.f) + 1e-5f);
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + lane * 8));
  const float4 g1 = __ldg(reinterpret_cast<const float4*>(g + lane * 8 + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + lane * 8));
  const float4 b1 = __ldg(reinterpret_cast<const float4*>(b + lane * 8 + 4));
  const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = (v[e] - mean) * rstd * gg[e] + bb[e];
}

__device__ __forceinline__ void load8(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}

__device__ __forceinline__ float silu(float x) { return x / (1.f + expf(-x)); }

__global__ void __launch_bounds__(256) row_post_kernel(const RowPostArgs a) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= a.M) return;
  const long off = static_cast<long>(row) * 256 + lane * 8;
  float v[8];
  if (a.partials != nullptr) {
    load8(a.partials + off, v);
    for (int s = 1; s < a.splits; ++s) {
      float w[8];
      load8(a.partials + s * a.split_stride + off, w);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] += w[e];
    }
  } else {
    const uint4 h = *reinterpret_cast<const uint4*>(a.in_f16 + off);
    const __half2* hp = reinterpret_cast<const __half2*>(&h);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(hp[e]);
      v[2 * e] = f.x;
      v[2 * e + 1] = f.y;
    }
  }
  if (a.bias != nullptr) {
    float w[8];
    load8(a.bias + lane * 8, w);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += w[e];
  }
  if (a.ln1_g != nullptr) ln8(v, a.ln1_g, a.ln1_b, lane);
  if (a.relu1) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = fmaxf(v[e], 0.f);
  }
  if (a.resid != nullptr) {
    float w[8];
    load8(a.resid + off, w);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += w[e];
  }
  if (a.ln2_g != nullptr) ln8(v, a.ln2_g, a.ln2_b, lane);
  float h16[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float y = v[e];
    if (a.act2 == 1) y = fmaxf(y, 0.f);
    else if (a.act2 == 2) y = silu(y);
    h16[e] = y;
    if (!a.act2_f16_only) v[e] = y;
  }
  if (a.out_f32 != nullptr) {
    *reinterpret_cast<float4*>(a.out_f32 + off) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(a.out_f32 + off + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  if (a.out_f16 != nullptr) *reinterpret_cast<uint4*>(a.out_f16 + off) = make_uint4(
      pack2h(h16[0], h16[1]), pack2h(h16[2], h16[3]), pack2h(h16[4], h16[5]), pack2h(h16[6], h16[7]));
  if (a.out_mod_f16 != nullptr) {
    const int grp = row / a.rows_per_group;
    float sc[8], sh[8];
    load8(a.mod_scale + static_cast<long>(grp) * a.scale_stride + lane * 8, sc);
    if (a.shift_per_row) load8(a.mod_shift + off, sh);
    else load8(a.mod_shift + static_cast<long>(grp) * a.shift_stride + lane * 8, sh);
    float m[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) m[e] = __fadd_rn(__fmul_rn(v[e], __fadd_rn(sc[e], 1.0f)), sh[e]);
    *reinterpret_cast<uint4*>(a.out_mod_f16 + off) =
        make_uint4(pack2h(m[0], m[1]), pack2h(m[2], m[3]), pack2h(m[4], m[5]), pack2h(m[6], m[7]));
  }
}

// ------------------------------------------------------------------------------------------------ small linear
// out[m][n] = act_out( bias[n] + sum_k act_in(a[m][k]) * w[n][k] ), m <= 8 rows (the per-frame time embeddings of
// box_head.py:218-223 and block_time_mlp :464,:602).  fp32 activations, fp16 weights, fp32 accumulate.
// One warp per output feature; act_in: 0 none, 1 SiLU; act_out: 0 none, 1 GELU(erf).
__global__ void __launch_bounds__(256)
small_linear_kernel(const float* __restrict__ a, const __half* __restrict__ w, const float* __restrict__ bias,
                    float* __restrict__ out, int m, int n, int k, int act_in, int act_out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int col = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (col >= n) return;
  float acc[8];
#pragma unroll
  for (int r = 0; r < 8; ++r) acc[r] = 0.f;
  for (int k0 = lane * 8; k0 < k; k0 += 256) {
    const uint4 wv = __ldg(reinterpret_cast<const uint4*>(w + static_cast<long>(col) * k + k0));
    const __half2* hp = reinterpret_cast<const __half2*>(&wv);
    float wf[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 f = __half22float2(hp[e]);
      wf[2 * e] = f.x;
      wf[2 * e + 1] = f.y;
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      if (r < m) {
        float x[8];
        load8(a + static_cast<long>(r) * k + k0, x);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          const float xv = act_in == 1 ? silu(x[e]) : x[e];
          acc[r] = fmaf(xv, wf[e], acc[r]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    if (r < m) {
      float s = warp_sum(acc[r]);
      if (lane == 0) {
        if (bias != nullptr) s += bias[col];
        if (act_out == 1) s = 0.5f * s * (1.f + erff(s * 0.70710678118654752440f));
        out[static_cast<long>(r) * n + col] = s;
      }
    }
  }
}

// sinusoidal time embedding (box_head.py:729-741), dim 256: out[m][0:128] = sin(t*f_k), [128:256] = cos(t*f_k)
// `freq` = exp(arange(128) * -(ln 10000 / 127)) is passed in (computed once by the host exactly as the reference does).
__global__ void time_sinusoid_kernel(const float* __restrict__ t, const float* __restrict__ freq,
                                     float* __restrict__ out, int m) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= m * 128) return;
  const int r = i / 128, kx = i % 128;
  const float arg = __fmul_rn(t[r], freq[kx]);
  out[r * 256 + kx] = sinf(arg);
  out[r * 256 + 128 + kx] = cosf(arg);
}

// ------------------------------------------------------------------------------------------------ boxes
constexpr float kScaleClamp = 8.740336742730447f;   // log(100000/16), box_head.py scale_clamp

// class_logits / bboxes_delta epilogue + apply_deltas (box_head.py:544-590).  logit_part [M][ldl] (first C valid),
// delta_part [M][ldd] (first 4 valid) are the fp32 GEMM outputs without bias.
__global__ void head_final_kernel(const float* __restrict__ logit_part, int ldl, const float* __restrict__ cls_bias,
                                  int C, const float* __restrict__ delta_part, int ldd,
                                  const float* __restrict__ delta_bias, const float* __restrict__ boxes_in,
                                  float* __restrict__ logits_out, float* __restrict__ boxes_out, int M) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  for (int c = 0; c < C; ++c) logits_out[static_cast<long>(i) * C + c] = logit_part[static_cast<long>(i) * ldl + c] + cls_bias[c];
  const float4 b = *reinterpret_cast<const float4*>(boxes_in + static_cast<long>(i) * 4);
  const float* d = delta_part + static_cast<long>(i) * ldd;
  const float w = __fsub_rn(b.z, b.x), h = __fsub_rn(b.w, b.y);
  const float cx = __fadd_rn(b.x, __fmul_rn(0.5f, w)), cy = __fadd_rn(b.y, __fmul_rn(0.5f, h));
  const float dx = __fdiv_rn(d[0] + delta_bias[0], 2.0f), dy = __fdiv_rn(d[1] + delta_bias[1], 2.0f);
  const float dw = fminf(d[2] + delta_bias[2], kScaleClamp), dh = fminf(d[3] + delta_bias[3], kScaleClamp);
  const float pcx = __fadd_rn(__fmul_rn(dx, w), cx), pcy = __fadd_rn(__fmul_rn(dy, h), cy);
  const float pw = __fmul_rn(expf(dw), w), ph = __fmul_rn(expf(dh), h);
  float4 o;
  o.x = __fsub_rn(pcx, __fmul_rn(0.5f, pw));
  o.y = __fsub_rn(pcy, __fmul_rn(0.5f, ph));
  o.z = __fadd_rn(pcx, __fmul_rn(0.5f, pw));
  o.w = __fadd_rn(pcy, __fmul_rn(0.5f, ph));
  *reinterpret_cast<float4*>(boxes_out + static_cast<long>(i) * 4) = o;
}

// noise-space boxes -> absolute xyxy (diffusion_det.py:657-660): clamp(+-scale), (x/scale+1)/2, cxcywh->xyxy, *whwh
__device__ __forceinline__ float4 noise_to_box(float4 x, float scale, float W, float H) {
  float v[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    v[e] = fminf(fmaxf(v[e], -scale), scale);
    v[e] = __fdiv_rn(__fadd_rn(__fdiv_rn(v[e], scale), 1.0f), 2.0f);
  }
  float4 o;
  o.x = __fmul_rn(__fsub_rn(v[0], __fmul_rn(0.5f, v[2])), W);
  o.y = __fmul_rn(__fsub_rn(v[1], __fmul_rn(0.5f, v[3])), H);
  o.z = __fmul_rn(__fadd_rn(v[0], __fmul_rn(0.5f, v[2])), W);
  o.w = __fmul_rn(__fadd_rn(v[1], __fmul_rn(0.5f, v[3])), H);
  return o;
}

__global__ void noise_to_boxes_kernel(const float* __restrict__ x, float* __restrict__ boxes, int M, float scale,
                                      float W, float H) {
  pdl_prologue();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M) return;
  *reinterpret_cast<float4*>(boxes + static_cast<long>(i) * 4) =
      noise_to_box(*reinterpret_cast<const float4*>(x + static_cast<long>(i) * 4), scale, W, H);
}

// One DDIM step with box renewal for one frame per CTA (diffusion_det.py:559-596, :668-676):
//   x_start = clamp((cxcywh(coord / whwh) * 2 - 1) * scale); pred_noise = (sqrt(1/a_t) * x_t - x_start) / sqrt(1/a_t - 1)
//   keep = sigmoid(max_c logit) > 0.5; kept boxes (compacted, order preserved):
//   x_next = x_start * sqrt(a_next) + c * pred_noise + sigma * eps[j]; the rest is refilled from `fill` in order.
// Also emits the next step's absolute boxes.  N <= 1024.
__global__ void __launch_bounds__(1024)
ddim_step_kernel(const float* __restrict__ logits, int C, const float* __restrict__ coord,
                 const float* __restrict__ x_t, const float* __restrict__ eps, const float* __restrict__ fill,
                 float* __restrict__ x_next, float* __restrict__ boxes_next, int* __restrict__ num_kept, int N,
                 float scale, float W, float H, float sqrt_recip_a, float sqrt_recipm1_a, float sqrt_a_next, float c_coef,
                 float sigma) {
  pdl_prologue();
  __shared__ int warp_cnt[32];
  __shared__ int warp_off[33];
  const int f = blockIdx.x;
  const int n = threadIdx.x;
  const int lane = n & 31, warp = n >> 5;
  bool keep = false;
  float xs[4] = {0.f, 0.f, 0.f, 0.f}, pn[4] = {0.f, 0.f, 0.f, 0.f};
  if (n < N) {
    const long row = static_cast<long>(f) * N + n;
    float mx = -INFINITY;
    for (int c = 0; c < C; ++c) mx = fmaxf(mx, logits[row * C + c]);
    const float sg = __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-mx)));
    keep = sg > 0.5f;
    const float4 b = *reinterpret_cast<const float4*>(coord + row * 4);
    const float bx = __fdiv_rn(b.x, W), by = __fdiv_rn(b.y, H), bz = __fdiv_rn(b.z, W), bw = __fdiv_rn(b.w, H);
    float cxy[4] = {__fdiv_rn(__fadd_rn(bx, bz), 2.0f), __fdiv_rn(__fadd_rn(by, bw), 2.0f), __fsub_rn(bz, bx),
                    __fsub_rn(bw, by)};
    const float4 xt = *reinterpret_cast<const float4*>(x_t + row * 4);
    const float xtv[4] = {xt.x, xt.y, xt.z, xt.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = __fmul_rn(__fsub_rn(__fmul_rn(cxy[e], 2.0f), 1.0f), scale);
      v = fminf(fmaxf(v, -scale), scale);
      xs[e] = v;
      pn[e] = __fdiv_rn(__fsub_rn(__fmul_rn(sqrt_recip_a, xtv[e]), v), sqrt_recipm1_a);
    }
  }
  const unsigned ball = __ballot_sync(0xffffffffu, keep);
  if (lane == 0) warp_cnt[warp] = __popc(ball);
  __syncthreads();
  if (n == 0) {
    int s = 0;
    for (int w = 0; w < 32; ++w) {
      warp_off[w] = s;
      s += (w * 32 < N) ? warp_cnt[w] : 0;
    }
    warp_off[32] = s;
    if (num_kept) num_kept[f] = s;
  }
  __syncthreads();
  const int kept = warp_off[32];
  if (n < N) {
    const long base = static_cast<long>(f) * N;
    if (keep) {
      const int pos = warp_off[warp] + __popc(ball & ((1u << lane) - 1u));
      const float4 e4 = *reinterpret_cast<const float4*>(eps + (base + pos) * 4);
      const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
      float o[4];
#pragma unroll
      for (int e = 0; e < 4; ++e)
        o[e] = __fadd_rn(__fadd_rn(__fmul_rn(xs[e], sqrt_a_next), __fmul_rn(c_coef, pn[e])), __fmul_rn(sigma, ev[e]));
      const float4 o4 = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(x_next + (base + pos) * 4) = o4;
      *reinterpret_cast<float4*>(boxes_next + (base + pos) * 4) = noise_to_box(o4, scale, W, H);
    }
    // refill: thread n handles refill slot n (if any)
    if (n < N - kept) {
      const float4 fz = *reinterpret_cast<const float4*>(fill + (base + n) * 4);
      *reinterpret_cast<float4*>(x_next + (base + kept + n) * 4) = fz;
      *reinterpret_cast<float4*>(boxes_next + (base + kept + n) * 4) = noise_to_box(fz, scale, W, H);
    }
  }
}

inline int grid_for(long total, int block) {
  long g = (total + block - 1) / block;
  const long cap = static_cast<long>(num_sms()) * 16;
  return static_cast<int>(g < cap ? (g < 1 ? 1 : g) : cap);
}

}  // namespace

int preprocess_launch(const void* img, int is_u8, void* out, int n, int H, int W, int halo, int Hp, int Wp,
                      const float* mean, const float* std, cudaStream_t stream) {
  if (n <= 0 || H <= 0 || W <= 0 || Hp < H + 2 * halo || Wp < W + 2 * halo) return DVID_ERR_SHAPE;
  if (Hp > 65535 || n > 65535) return DVID_ERR_SHAPE;
  const dim3 grid((Wp + 255) / 256, Hp, n);
  if (is_u8)
    launch_pdl(preprocess_kernel<uint8_t>, grid, dim3(256), 0, stream, static_cast<const uint8_t*>(img),
               static_cast<__half*>(out), n, H, W, halo, Hp, Wp, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
  else
    launch_pdl(preprocess_kernel<float>, grid, dim3(256), 0, stream, static_cast<const float*>(img),
               static_cast<__half*>(out), n, H, W, halo, Hp, Wp, mean[0], mean[1], mean[2], std[0], std[1], std[2]);
  return check_launch();
}

int maxpool_launch(const void* in, void* out, int n, int H, int W, int C, cudaStream_t stream) {
  if (n <= 0 || H <= 0 || W <= 0 || C % 8 != 0) return DVID_ERR_SHAPE;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  if (Ho > 65535 || n > 65535) return DVID_ERR_SHAPE;
  launch_pdl(maxpool3x3s2_kernel, dim3((Wo * (C / 8) + 255) / 256, Ho, n), dim3(256), 0, stream, static_cast<const __half*>(in),
                                                                static_cast<__half*>(out), n, H, W, C, Ho, Wo);
  return check_launch();
}

int row_post_launch(const float* partials, int splits, long split_stride, const void* in_f16, const float* bias,
                    const float* ln1_g, const float* ln1_b, int relu1, const float* resid, const float* ln2_g,
                    const float* ln2_b, int act2, int act2_f16_only, float* out_f32, void* out_f16,
                    const float* mod_scale, const float* mod_shift, int rows_per_group, int scale_stride,
                    int shift_stride, int shift_per_row, void* out_mod_f16, int M, cudaStream_t stream) {
  if (M <= 0 || (partials == nullptr && in_f16 == nullptr)) return DVID_ERR_ARG;
  if (out_mod_f16 != nullptr && (mod_scale == nullptr || mod_shift == nullptr || rows_per_group <= 0))
    return DVID_ERR_ARG;
  RowPostArgs a;
  a.partials = partials; a.splits = splits; a.split_stride = split_stride;
  a.in_f16 = static_cast<const __half*>(in_f16);
  a.bias = bias; a.ln1_g = ln1_g; a.ln1_b = ln1_b; a.relu1 = relu1; a.resid = resid; a.ln2_g = ln2_g; a.ln2_b = ln2_b;
  a.act2 = act2; a.act2_f16_only = act2_f16_only; a.out_f32 = out_f32; a.out_f16 = static_cast<__half*>(out_f16);
  a.mod_scale = mod_scale; a.mod_shift = mod_shift; a.rows_per_group = rows_per_group; a.scale_stride = scale_stride;
  a.shift_stride = shift_stride; a.shift_per_row = shift_per_row; a.out_mod_f16 = static_cast<__half*>(out_mod_f16);
  a.M = M;
  launch_pdl(row_post_kernel, dim3((M + 7) / 8), dim3(256), 0, stream, a);
  return check_launch();
}

int small_linear_launch(const float* a, const void* w, const float* bias, float* out, int m, int n, int k, int act_in,
                        int act_out, cudaStream_t stream) {
  if (m <= 0 || m > 8 || n <= 0 || k <= 0 || k % 8 != 0) return DVID_ERR_SHAPE;
  launch_pdl(small_linear_kernel, dim3((n + 7) / 8), dim3(256), 0, stream, a, static_cast<const __half*>(w), bias, out, m, n, k, act_in,
                                                        act_out);
  return check_launch();
}

int time_sinusoid_launch(const float* t, const float* freq, float* out, int m, cudaStream_t stream) {
  if (m <= 0) return DVID_ERR_SHAPE;
  launch_pdl(time_sinusoid_kernel, dim3((m * 128 + 127) / 128), dim3(128), 0, stream, t, freq, out, m);
  return check_launch();
}

int head_final_launch(const float* logit_part, int ldl, const float* cls_bias, int C, const float* delta_part, int ldd,
                      const float* delta_bias, const float* boxes_in, float* logits_out, float* boxes_out, int M,
                      cudaStream_t stream) {
  if (M <= 0 || C <= 0 || C > ldl || ldd < 4) return DVID_ERR_SHAPE;
  launch_pdl(head_final_kernel, dim3((M + 127) / 128), dim3(128), 0, stream, logit_part, ldl, cls_bias, C, delta_part, ldd, delta_bias,
                                                        boxes_in, logits_out, boxes_out, M);
  return check_launch();
}

int noise_to_boxes_launch(const float* x, float* boxes, int M, float scale, float W, float H, cudaStream_t stream) {
  if (M <= 0) return DVID_ERR_SHAPE;
  launch_pdl(noise_to_boxes_kernel, dim3((M + 127) / 128), dim3(128), 0, stream, x, boxes, M, scale, W, H);
  return check_launch();
}

int ddim_step_launch(const float* logits, int C, const float* coord, const float* x_t, const float* eps,
                     const float* fill, float* x_next, float* boxes_next, int* num_kept, int frames, int N,
                     float scale, float W, float H, float sqrt_recip_a, float sqrt_recipm1_a, float sqrt_a_next,
                     float c_coef, float sigma, cudaStream_t stream) {
  if (frames <= 0 || N <= 0 || N > 1024) return DVID_ERR_SHAPE;
  launch_pdl(ddim_step_kernel, dim3(frames), dim3(1024), 0, stream, logits, C, coord, x_t, eps, fill, x_next, boxes_next, num_kept, N,
                                               scale, W, H, sqrt_recip_a, sqrt_recipm1_a, sqrt_a_next, c_coef, sigma);
  return check_launch();
}

}  // namespace dvid

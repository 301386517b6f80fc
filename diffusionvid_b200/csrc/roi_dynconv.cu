// Multi-level ROIAlign + DynamicConv instance interaction, one CTA per box.
//
// Replaces, for every box of every frame of a DDIM step:
//   * detectron2 ROIPooler -> torchvision roi_align(aligned=True, 7x7, sampling_ratio 2) over the p3/p4/p5 maps
//     (call sites mega_core/modeling/roi_heads/box_head/box_head.py:507,617; level rule SURVEY.md A2/A3)
//   * DynamicConv.forward's two per-box bmm's with their LayerNorm+ReLU (box_head.py:698-704)
// Feature maps are NHWC fp16 (256 channels = one 512-byte vector per tap, read by one warp as 32 x 16 B), so the
// data-dependent bilinear gather is fully coalesced.  The 7x7x256 ROI tile, the box's generated 256x64 / 64x256
// weights (fp16, produced by the tcgen05 `dynamic_layer` GEMM) and the 49x64 intermediate all live in shared memory;
// only the final 49x256 fp16 activations go back to HBM (they feed the K=12544 `out_layer` tcgen05 GEMM).
//
// Two kernels share the gather code:
//   roi_align_kernel      : ROI tile -> global (and the per-box mean that seeds pro_features, box_head.py:509-510)
//   roi_dynconv_kernel<G> : G=false gathers the ROI tile itself (fused ROIAlign), G=true reads it from global.
#include <stdlib.h>
#include "ptx_sm100.cuh"
#include "dvid_internal.h"
#include "warp_mma.cuh"

namespace dvid {

namespace {

constexpr int D = 256;     // hidden dim / feature channels
constexpr int DD = 64;     // dynamic dim
constexpr int P = 7;       // pooler resolution
constexpr int NBIN = P * P;

struct RoiLevels {
  const __half* feat[3];   // NHWC fp16 [frames][H_l][W_l][256]
  int h[3], w[3];
  float scale[3];
};

// ---- shared-memory layouts (all fp16, 16-byte chunks XOR-swizzled by row&7 so ldmatrix phases are conflict-free)
__device__ __forceinline__ uint32_t off512(int row, int chunk) {   // rows of 256 halfs (32 chunks)
  return static_cast<uint32_t>(row * 512 + ((chunk ^ (row & 7)) << 4));
}
__device__ __forceinline__ uint32_t off128(int row, int chunk) {   // rows of 64 halfs (8 chunks)
  return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

// One axis of the 2-sample bilinear footprint of a bin: up to 4 (index, weight) taps, duplicates merged.
struct AxisTaps {
  int idx[4];
  float w[4];
};

// torchvision bilinear_interpolate semantics along one axis: sample invalid if v < -1 or v > size;
// clamp to >= 0; low = (int)v; if low >= size-1 -> low = high = size-1, v = low.
__device__ __forceinline__ void axis_taps(float start, float bin, int pbin, int size, AxisTaps& a) {
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    // y = roi_start + ph*bin + (iy + .5) * bin / 2
    float v = __fadd_rn(__fadd_rn(start, __fmul_rn(static_cast<float>(pbin), bin)),
                        __fdiv_rn(__fmul_rn(static_cast<float>(i) + 0.5f, bin), 2.0f));
    const bool bad = (v < -1.0f) || (v > static_cast<float>(size));
    v = fmaxf(v, 0.f);
    int lo = static_cast<int>(v);
    int hi;
    if (lo >= size - 1) {
      lo = hi = size - 1;
      v = static_cast<float>(lo);
    } else {
      hi = lo + 1;
    }
    const float l = v - static_cast<float>(lo);
    const float h = 1.0f - l;
    a.idx[2 * i] = lo;
    a.w[2 * i] = bad ? 0.f : h;
    a.idx[2 * i + 1] = hi;
    a.w[2 * i + 1] = bad ? 0.f : l;
  }
  // merge duplicate indices (adjacent samples of a small bin share taps)
#pragma unroll
  for (int i = 1; i < 4; ++i) {
#pragma unroll
    for (int j = 0; j < i; ++j) {
      if (a.idx[i] == a.idx[j] && a.w[i] != 0.f) {
        a.w[j] += a.w[i];
        a.w[i] = 0.f;
      }
    }
  }
}

struct RoiGeom {
  const __half* feat;   // frame base
  int H, W;
  float x1, y1, bw, bh;
};

// Level assignment (detectron2 assign_boxes_to_levels) + aligned ROI geometry for box `b`.
__device__ __forceinline__ RoiGeom roi_geometry(const RoiLevels& lv, const float* __restrict__ boxes, int b,
                                               int boxes_per_frame) {
  const float4 bx = __ldg(reinterpret_cast<const float4*>(boxes) + b);
  const float area = __fmul_rn(__fsub_rn(bx.z, bx.x), __fsub_rn(bx.w, bx.y));
  const float size = sqrtf(area);
  float lf = floorf(__fadd_rn(4.0f, log2f(__fadd_rn(__fdiv_rn(size, 224.0f), 1e-8f))));
  // NaN (negative area) compares false everywhere -> clamp to the lowest level like torch.clamp(min) then long cast
  int l = (lf >= 5.0f) ? 2 : ((lf >= 4.0f) ? 1 : 0);
  const int frame = b / boxes_per_frame;
  RoiGeom gm;
  gm.H = lv.h[l];
  gm.W = lv.w[l];
  gm.feat = lv.feat[l] + static_cast<long>(frame) * gm.H * gm.W * D;
  const float sc = lv.scale[l];
  gm.x1 = __fsub_rn(__fmul_rn(bx.x, sc), 0.5f);
  gm.y1 = __fsub_rn(__fmul_rn(bx.y, sc), 0.5f);
  const float x2 = __fsub_rn(__fmul_rn(bx.z, sc), 0.5f);
  const float y2 = __fsub_rn(__fmul_rn(bx.w, sc), 0.5f);
  gm.bw = __fdiv_rn(__fsub_rn(x2, gm.x1), static_cast<float>(P));
  gm.bh = __fdiv_rn(__fsub_rn(y2, gm.y1), static_cast<float>(P));
  return gm;
}

// Gather one 7x7 bin (all 256 channels) with the calling warp: lane owns channels [8*lane, 8*lane+8).
template <bool kDeep>
__device__ __forceinline__ void roi_bin(const RoiGeom& gm, int bin, int lane, float (&acc)[8]) {
  const int ph = bin / P, pw = bin - ph * P;
  AxisTaps ty, tx;
  axis_taps(gm.y1, gm.bh, ph, gm.H, ty);
  axis_taps(gm.x1, gm.bw, pw, gm.W, tx);
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  if (kDeep) {
    // all (up to 16) taps of the bin are requested before the first one is consumed: inside the fused kernel (2 CTAs
    // per SM) the gather is bound by the number of 512-byte requests in flight, not by arithmetic
    uint4 v[4][4];
#pragma unroll
    for (int iy = 0; iy < 4; ++iy) {
      if (ty.w[iy] == 0.f) continue;   // warp-uniform
      const __half* rowp = gm.feat + static_cast<long>(ty.idx[iy]) * gm.W * D + lane * 8;
#pragma unroll
      for (int ix = 0; ix < 4; ++ix) {
        if (tx.w[ix] != 0.f) v[iy][ix] = __ldg(reinterpret_cast<const uint4*>(rowp + static_cast<long>(tx.idx[ix]) * D));
      }
    }
#pragma unroll
    for (int iy = 0; iy < 4; ++iy) {
      if (ty.w[iy] == 0.f) continue;
#pragma unroll
      for (int ix = 0; ix < 4; ++ix) {
        if (tx.w[ix] != 0.f) {
          const float wgt = ty.w[iy] * tx.w[ix];
          const __half2* hp = reinterpret_cast<const __half2*>(&v[iy][ix]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            acc[2 * e] = fmaf(wgt, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(wgt, f.y, acc[2 * e + 1]);
          }
        }
      }
    }
  } else {
    // one tap row at a time (few registers: the stand-alone ROIAlign kernel runs at full occupancy instead)
#pragma unroll
    for (int iy = 0; iy < 4; ++iy) {
      if (ty.w[iy] == 0.f) continue;   // warp-uniform
      const __half* rowp = gm.feat + static_cast<long>(ty.idx[iy]) * gm.W * D + lane * 8;
      uint4 v[4];
#pragma unroll
      for (int ix = 0; ix < 4; ++ix) {
        if (tx.w[ix] != 0.f) v[ix] = __ldg(reinterpret_cast<const uint4*>(rowp + static_cast<long>(tx.idx[ix]) * D));
      }
#pragma unroll
      for (int ix = 0; ix < 4; ++ix) {
        if (tx.w[ix] != 0.f) {
          const float wgt = ty.w[iy] * tx.w[ix];
          const __half2* hp = reinterpret_cast<const __half2*>(&v[ix]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            acc[2 * e] = fmaf(wgt, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(wgt, f.y, acc[2 * e + 1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= 0.25f;   // mean of the 2x2 samples
}

// Per-box tap tables: the 7 bins of each axis share nothing with the other axis, so the 14 axis footprints are
// computed ONCE per box (14 threads) instead of once per bin and warp (49 x 2 evaluations of ~150 instructions, about
// half of the gather's instruction stream); offsets are pre-multiplied element offsets into the frame's feature map.
struct TapTable {
  int xoff[P][4];     // tap column * D
  float xw[P][4];
  int yoff[P][4];     // tap row * W * D
  float yw[P][4];
};

__device__ __forceinline__ void build_tap_table(const RoiGeom& gm, TapTable& tb, int tid) {
  if (tid < 2 * P) {
    const bool is_y = tid >= P;
    const int pb = is_y ? tid - P : tid;
    AxisTaps a;
    if (is_y) axis_taps(gm.y1, gm.bh, pb, gm.H, a);
    else axis_taps(gm.x1, gm.bw, pb, gm.W, a);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (is_y) { tb.yoff[pb][i] = a.idx[i] * gm.W * D; tb.yw[pb][i] = a.w[i]; }
      else { tb.xoff[pb][i] = a.idx[i] * D; tb.xw[pb][i] = a.w[i]; }
    }
  }
}

// fp32 += fp16 * fp16 in ONE instruction (FHFMA, PTX mixed-precision fma): the product of two halfs is exact in fp32.
__device__ __forceinline__ float fhfma(unsigned short a, unsigned short b, float c) {
  float d;
  asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(a), "h"(b), "f"(c));
  return d;
}

// Gather one bin with the calling warp from the tap table (all <= 16 taps in flight before the first is consumed).
// kHalfW: the tap weight is rounded to fp16 and applied with FHFMA - 8 instructions per tap instead of 8 conversions +
// 8 FFMA (the gather is instruction-issue bound).  The fp16 x fp16 product is exact and the accumulation stays fp32; the
// only difference to the fp32-weight path is the rounding of the weight (<= 2^-12 relative per tap, i.e. at most half an
// fp16 ulp of the fp16-rounded result).
template <bool kHalfW>
__device__ __forceinline__ void roi_bin_table(const RoiGeom& gm, const TapTable& tb, int bin, int lane,
                                              float (&acc)[8]) {
  const int ph = bin / P, pw = bin - ph * P;
  const int4 yo = *reinterpret_cast<const int4*>(tb.yoff[ph]);
  const float4 yw = *reinterpret_cast<const float4*>(tb.yw[ph]);
  const int4 xo = *reinterpret_cast<const int4*>(tb.xoff[pw]);
  const float4 xw = *reinterpret_cast<const float4*>(tb.xw[pw]);
  const int yoff[4] = {yo.x, yo.y, yo.z, yo.w};
  const float ywt[4] = {yw.x, yw.y, yw.z, yw.w};
  const int xoff[4] = {xo.x, xo.y, xo.z, xo.w};
  const float xwt[4] = {xw.x, xw.y, xw.z, xw.w};
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] = 0.f;
  const __half* base = gm.feat + lane * 8;
  uint4 v[4][4];
#pragma unroll
  for (int iy = 0; iy < 4; ++iy) {
    if (ywt[iy] == 0.f) continue;   // warp-uniform
#pragma unroll
    for (int ix = 0; ix < 4; ++ix) {
      if (xwt[ix] != 0.f) v[iy][ix] = __ldg(reinterpret_cast<const uint4*>(base + (yoff[iy] + xoff[ix])));
    }
  }
#pragma unroll
  for (int iy = 0; iy < 4; ++iy) {
    if (ywt[iy] == 0.f) continue;
#pragma unroll
    for (int ix = 0; ix < 4; ++ix) {
      if (xwt[ix] != 0.f) {
        const float wgt = ywt[iy] * xwt[ix];
        if (kHalfW) {
          const unsigned short wh = __half_as_ushort(__float2half_rn(wgt));
          const uint32_t w4[4] = {v[iy][ix].x, v[iy][ix].y, v[iy][ix].z, v[iy][ix].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            acc[2 * e] = fhfma(static_cast<unsigned short>(w4[e] & 0xffffu), wh, acc[2 * e]);
            acc[2 * e + 1] = fhfma(static_cast<unsigned short>(w4[e] >> 16), wh, acc[2 * e + 1]);
          }
        } else {
          const __half2* hp = reinterpret_cast<const __half2*>(&v[iy][ix]);
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float2 f = __half22float2(hp[e]);
            acc[2 * e] = fmaf(wgt, f.x, acc[2 * e]);
            acc[2 * e + 1] = fmaf(wgt, f.y, acc[2 * e + 1]);
          }
        }
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 8; ++e) acc[e] *= 0.25f;   // mean of the 2x2 samples
}

__device__ __forceinline__ uint4 pack8(const float (&a)[8]) {
  uint4 r;
  r.x = pack2h(a[0], a[1]);
  r.y = pack2h(a[2], a[3]);
  r.z = pack2h(a[4], a[5]);
  r.w = pack2h(a[6], a[7]);
  return r;
}

// ------------------------------------------------------------------------------------------------ ROIAlign only
__global__ void __launch_bounds__(256)
roi_align_kernel(RoiLevels lv, const float* __restrict__ boxes, int boxes_per_frame, __half* __restrict__ roi_out,
                 float* __restrict__ mean_f32, __half* __restrict__ mean_f16) {
  pdl_prologue();
  __shared__ float sred[8][D];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const RoiGeom gm = roi_geometry(lv, boxes, b, boxes_per_frame);
  float msum[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) msum[e] = 0.f;
  for (int bin = warp; bin < NBIN; bin += 8) {
    float acc[8];
    roi_bin<false>(gm, bin, lane, acc);
    const uint4 pk = pack8(acc);
    if (roi_out) *reinterpret_cast<uint4*>(roi_out + (static_cast<long>(b) * NBIN + bin) * D + lane * 8) = pk;
    const __half2* hp = reinterpret_cast<const __half2*>(&pk);
#pragma unroll
    for (int e = 0; e < 4; ++e) {   // the mean is taken over the fp16-rounded values that downstream code sees
      const float2 f = __half22float2(hp[e]);
      msum[2 * e] += f.x;
      msum[2 * e + 1] += f.y;
    }
  }
  if (mean_f32 == nullptr && mean_f16 == nullptr) return;
#pragma unroll
  for (int e = 0; e < 8; ++e) sred[warp][lane * 8 + e] = msum[e];
  __syncthreads();
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < 8; ++w) s += sred[w][c];
  s = __fdiv_rn(s, static_cast<float>(NBIN));
  if (mean_f32) mean_f32[static_cast<long>(b) * D + c] = s;
  if (mean_f16) mean_f16[static_cast<long>(b) * D + c] = __float2half_rn(s);
}

// ------------------------------------------------------------------------------------------------ fused DynamicConv
constexpr int SM_ROI = 0;                   // 64 x 512 B (rows 49..63 zero); reused as the output staging tile
constexpr int SM_P1 = SM_ROI + 64 * 512;    // 256 x 128 B
constexpr int SM_P2 = SM_P1 + 256 * 128;    // 64 x 512 B
constexpr int SM_F1 = SM_P2 + 64 * 512;     // 64 x 128 B
constexpr int SM_STAT = SM_F1 + 64 * 128;   // 64 x 2 floats
constexpr int SM_TOTAL = SM_STAT + 64 * 2 * 4;

// LayerNorm statistics of rows split across the two warps (nh = 0/1) that share a 16-row slab.
// part[h] = this thread's partial for row g + 8h; returns the full-row sum for both rows.
__device__ __forceinline__ void row_allreduce(float (&part)[2], float* sstat, int row0, int g, int t, int nh) {
#pragma unroll
  for (int h = 0; h < 2; ++h) part[h] = quad_sum(part[h]);
  __syncthreads();   // previous users of sstat are done
  if (t == 0) {
    sstat[(row0 + g) * 2 + nh] = part[0];
    sstat[(row0 + g + 8) * 2 + nh] = part[1];
  }
  __syncthreads();
  part[0] = sstat[(row0 + g) * 2] + sstat[(row0 + g) * 2 + 1];
  part[1] = sstat[(row0 + g + 8) * 2] + sstat[(row0 + g + 8) * 2 + 1];
}

template <bool kRoiFromGlobal, bool kHalfW = false>
__global__ void __launch_bounds__(256, 2)
roi_dynconv_kernel(RoiLevels lv, const float* __restrict__ boxes, int boxes_per_frame,
                   const __half* __restrict__ roi_in,      // [M][49][256] (kRoiFromGlobal)
                   const __half* __restrict__ params,      // [M][2*256*64]: P1 [256][64] then P2 [64][256]
                   const float* __restrict__ g1, const float* __restrict__ b1,   // LayerNorm(64)
                   const float* __restrict__ g2, const float* __restrict__ b2,   // LayerNorm(256)
                   __half* __restrict__ out) {
  pdl_prologue();              // [M][49][256]
  extern __shared__ __align__(128) uint8_t smem[];
  uint8_t* sRoi = smem + SM_ROI;
  uint8_t* sP1 = smem + SM_P1;
  uint8_t* sP2 = smem + SM_P2;
  uint8_t* sF1 = smem + SM_F1;
  float* sStat = reinterpret_cast<float*>(smem + SM_STAT);

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;

  // ---- async loads of the box's generated weights (and ROI tile when it comes from global)
  {
    const __half* p1 = params + static_cast<long>(b) * (2 * D * DD);
    const __half* p2 = p1 + D * DD;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int id = tid + i * 256;            // 2048 chunks: [256 rows][8 chunks]
      cp_async16(sP1 + off128(id >> 3, id & 7), p1 + id * 8, true);
    }
    if (kRoiFromGlobal) {
      const __half* r = roi_in + static_cast<long>(b) * NBIN * D;
      for (int id = tid; id < NBIN * 32; id += 256) cp_async16(sRoi + off512(id >> 5, id & 31), r + id * 8, true);
    }
    cp_async_commit();
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int id = tid + i * 256;            // 2048 chunks: [64 rows][32 chunks]
      cp_async16(sP2 + off512(id >> 5, id & 31), p2 + id * 8, true);
    }
    cp_async_commit();
  }
  // zero the padding rows 49..63 of the ROI tile
  for (int id = tid; id < (64 - NBIN) * 32; id += 256)
    *reinterpret_cast<uint4*>(sRoi + off512(NBIN + (id >> 5), id & 31)) = make_uint4(0, 0, 0, 0);

  if (!kRoiFromGlobal) {
    __shared__ __align__(16) TapTable taps;
    const RoiGeom gm = roi_geometry(lv, boxes, b, boxes_per_frame);
    build_tap_table(gm, taps, tid);
    __syncthreads();
    for (int bin = warp; bin < NBIN; bin += 8) {
      float acc[8];
      roi_bin_table<kHalfW>(gm, taps, bin, lane, acc);
      *reinterpret_cast<uint4*>(sRoi + off512(bin, lane)) = pack8(acc);
    }
  }
  cp_async_wait<1>();   // P1 (+ROI) landed
  __syncthreads();

  const int mt = warp & 3;    // 16-row slab
  const int nh = warp >> 2;   // column half
  const int lj = lane >> 3, lr = lane & 7;

  // ---- bmm1: F1[64x64] = ROI[64x256] @ P1[256x64]; this warp: rows mt*16.., cols nh*32..+31
  float acc1[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc1[i][j] = 0.f;
#pragma unroll 4
  for (int ks = 0; ks < D / 16; ++ks) {
    uint32_t a[4];
    ldmatrix_x4(a, smem_addr(sRoi + off512(mt * 16 + (lj & 1) * 8 + lr, ks * 2 + (lj >> 1))));
#pragma unroll
    for (int np = 0; np < 2; ++np) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, smem_addr(sP1 + off128(ks * 16 + (lj & 1) * 8 + lr, nh * 4 + np * 2 + (lj >> 1))));
      mma_16816(acc1[np * 2], a, bb[0], bb[1]);
      mma_16816(acc1[np * 2 + 1], a, bb[2], bb[3]);
    }
  }
  // LayerNorm(64) + ReLU -> sF1 (fp16)
  {
    float part[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      part[0] += acc1[i][0] + acc1[i][1];
      part[1] += acc1[i][2] + acc1[i][3];
    }
    row_allreduce(part, sStat, mt * 16, g, t, nh);
    const float mean0 = part[0] * (1.f / DD), mean1 = part[1] * (1.f / DD);
    part[0] = part[1] = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float d0 = acc1[i][0] - mean0, d1 = acc1[i][1] - mean0, d2 = acc1[i][2] - mean1, d3 = acc1[i][3] - mean1;
      part[0] += d0 * d0 + d1 * d1;
      part[1] += d2 * d2 + d3 * d3;
    }
    row_allreduce(part, sStat, mt * 16, g, t, nh);
    const float rstd0 = rsqrtf(part[0] * (1.f / DD) + 1e-5f), rstd1 = rsqrtf(part[1] * (1.f / DD) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int col = nh * 32 + i * 8 + 2 * t;
      const float2 gg = __ldg(reinterpret_cast<const float2*>(g1 + col));
      const float2 be = __ldg(reinterpret_cast<const float2*>(b1 + col));
      const float y0 = fmaxf((acc1[i][0] - mean0) * rstd0 * gg.x + be.x, 0.f);
      const float y1 = fmaxf((acc1[i][1] - mean0) * rstd0 * gg.y + be.y, 0.f);
      const float y2 = fmaxf((acc1[i][2] - mean1) * rstd1 * gg.x + be.x, 0.f);
      const float y3 = fmaxf((acc1[i][3] - mean1) * rstd1 * gg.y + be.y, 0.f);
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(sF1 + off128(r0, col >> 3) + (col & 7) * 2) = pack2h(y0, y1);
      *reinterpret_cast<uint32_t*>(sF1 + off128(r1, col >> 3) + (col & 7) * 2) = pack2h(y2, y3);
    }
  }
  cp_async_wait<0>();   // P2 landed
  __syncthreads();      // sF1 complete; all warps finished reading sRoi

  // ---- bmm2: F2[64x256] = F1[64x64] @ P2[64x256]; this warp: rows mt*16.., cols nh*128..+127
  float acc2[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc2[i][j] = 0.f;
#pragma unroll
  for (int ks = 0; ks < DD / 16; ++ks) {
    uint32_t a[4];
    ldmatrix_x4(a, smem_addr(sF1 + off128(mt * 16 + (lj & 1) * 8 + lr, ks * 2 + (lj >> 1))));
#pragma unroll
    for (int np = 0; np < 8; ++np) {
      uint32_t bb[4];
      ldmatrix_x4_trans(bb, smem_addr(sP2 + off512(ks * 16 + (lj & 1) * 8 + lr, nh * 16 + np * 2 + (lj >> 1))));
      mma_16816(acc2[np * 2], a, bb[0], bb[1]);
      mma_16816(acc2[np * 2 + 1], a, bb[2], bb[3]);
    }
  }
  // LayerNorm(256) + ReLU -> staging tile (reuses sRoi) -> global
  {
    float part[2] = {0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      part[0] += acc2[i][0] + acc2[i][1];
      part[1] += acc2[i][2] + acc2[i][3];
    }
    row_allreduce(part, sStat, mt * 16, g, t, nh);
    const float mean0 = part[0] * (1.f / D), mean1 = part[1] * (1.f / D);
    part[0] = part[1] = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      float d0 = acc2[i][0] - mean0, d1 = acc2[i][1] - mean0, d2 = acc2[i][2] - mean1, d3 = acc2[i][3] - mean1;
      part[0] += d0 * d0 + d1 * d1;
      part[1] += d2 * d2 + d3 * d3;
    }
    row_allreduce(part, sStat, mt * 16, g, t, nh);
    const float rstd0 = rsqrtf(part[0] * (1.f / D) + 1e-5f), rstd1 = rsqrtf(part[1] * (1.f / D) + 1e-5f);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const int col = nh * 128 + i * 8 + 2 * t;
      const float2 gg = __ldg(reinterpret_cast<const float2*>(g2 + col));
      const float2 be = __ldg(reinterpret_cast<const float2*>(b2 + col));
      const float y0 = fmaxf((acc2[i][0] - mean0) * rstd0 * gg.x + be.x, 0.f);
      const float y1 = fmaxf((acc2[i][1] - mean0) * rstd0 * gg.y + be.y, 0.f);
      const float y2 = fmaxf((acc2[i][2] - mean1) * rstd1 * gg.x + be.x, 0.f);
      const float y3 = fmaxf((acc2[i][3] - mean1) * rstd1 * gg.y + be.y, 0.f);
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      *reinterpret_cast<uint32_t*>(sRoi + off512(r0, col >> 3) + (col & 7) * 2) = pack2h(y0, y1);
      *reinterpret_cast<uint32_t*>(sRoi + off512(r1, col >> 3) + (col & 7) * 2) = pack2h(y2, y3);
    }
  }
  __syncthreads();
  {
    __half* o = out + static_cast<long>(b) * NBIN * D;
    for (int id = tid; id < NBIN * 32; id += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(sRoi + off512(id >> 5, id & 31));
      *reinterpret_cast<uint4*>(o + id * 8) = v;
    }
  }
}


// ------------------------------------------------------------------------------------------------ tcgen05 DynamicConv
// Same operation as roi_dynconv_kernel with the two per-box contractions on the 5th-generation tensor cores:
//   bmm1  F1[49x64]  = ROI[49x256] . P1[256x64]   one tcgen05.mma chain M=128 (rows >= 64 are don't-care, see below),
//                                                 N=64, K=256, accumulator in TMEM columns 0..63
//   LN(64)+ReLU      straight out of TMEM (tcgen05.ld), written back to shared memory as the next A operand
//   bmm2  F2[49x256] = F1[49x64] . P2[64x256]     M=128, N=256, K=64, accumulator in TMEM columns 0..255
//   LN(256)+ReLU     out of TMEM -> staging tile -> coalesced global store
// Operands are K-major, 128-byte swizzled (the layout umma_desc_sw128_kmajor describes):
//   ROI tile   4 k-blocks x [64 rows][128 B], written by the gather warps (st.shared + fence.proxy.async)
//   P1^T       4 k-blocks x [64 rows (j)][128 B (64 i)]   TMA, box {64,64,1} of the 3-D view [box][j][i]
//   P2^T       [256 rows (i)][128 B (64 j)]               TMA, box {64,256,1} of the 3-D view [box][i][j]
// so the generated weights arrive TRANSPOSED from the dynamic_layer GEMM: the rows of its weight matrix are permuted
// once at pack time (params[j*256+i] = P1[i][j], params[16384+i*64+j] = P2[j][i]; box_head.py:693-696 layout otherwise).
// A tcgen05 tile has at least 64 rows in the documented lane layout only for M=128, so every box is issued as an
// M=128 tile whose upper 64 rows read whatever follows the operand in shared memory (the next k-block / the next
// buffer, always inside the CTA's allocation); rows are independent in a GEMM, TMEM lanes 64..127 are never read.
// One box per CTA, two CTAs per SM (each allocates 256 TMEM columns), 8 warps: all gather, warp 1 issues the MMAs,
// warps 0,1,4,5 (TMEM sub-partitions 0 and 1 = lanes 0..63) run the LayerNorm(64) epilogue, all eight the LayerNorm(256)
// one (bmm2 is issued as two N=128 halves, the second into lanes 64..127), all store.
constexpr int TC_ROI = 0;                       // 4 x 8 KB; reused as the [64][512 B] output staging tile
constexpr int TC_P1 = TC_ROI + 4 * 8192;        // 4 x 8 KB
constexpr int TC_F1 = TC_P1 + 4 * 8192;         // 8 KB
constexpr int TC_P2 = TC_F1 + 8192;             // 32 KB (also the don't-care rows of the F1 operand)
constexpr int TC_LN = TC_P2 + 32768;            // gamma2 | beta2 (256 floats each)
constexpr int TC_STAT = TC_LN + 2 * 256 * 4;    // [4 column quarters][64 rows] float2
constexpr int TC_BAR = TC_STAT + 4 * 64 * 8;    // mbarriers + tmem slot
constexpr int TC_TOTAL = TC_BAR + 64 + 1024;    // + alignment slack

struct DynMaps {
  CUtensorMap p1;
  CUtensorMap p2;
};

template <bool kRoiFromGlobal>
__global__ void __launch_bounds__(256, 2)
roi_dynconv_tc_kernel(const __grid_constant__ DynMaps maps, RoiLevels lv, const float* __restrict__ boxes,
                      int boxes_per_frame, const __half* __restrict__ roi_in,
                      const float* __restrict__ g1, const float* __restrict__ b1,
                      const float* __restrict__ g2, const float* __restrict__ b2, __half* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sRoi = smem + TC_ROI;
  uint8_t* sP1 = smem + TC_P1;
  uint8_t* sF1 = smem + TC_F1;
  uint8_t* sP2 = smem + TC_P2;
  float* sG2 = reinterpret_cast<float*>(smem + TC_LN);
  float* sB2 = sG2 + 256;
  float2* sStat = reinterpret_cast<float2*>(smem + TC_STAT);
  uint64_t* p1_full = reinterpret_cast<uint64_t*>(smem + TC_BAR);
  uint64_t* p2_full = p1_full + 1;
  uint64_t* mma_bar = p1_full + 2;
  uint64_t* f1_ready = p1_full + 3;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(p1_full + 4);

  const int b = blockIdx.x;
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;
  pdl_trigger();
  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&maps.p1);
    tma_prefetch_desc(&maps.p2);
    mbar_init(p1_full, 1);
    mbar_init(p2_full, 1);
    mbar_init(mma_bar, 1);
    mbar_init(f1_ready, 128);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  // ---- the box's generated weights: TMA into the swizzled operand layouts
  if (warp == 1 && elect_one()) {
    mbar_expect_tx(p1_full, 4 * 8192);
#pragma unroll
    for (int kb = 0; kb < 4; ++kb) tma_load_3d(sP1 + kb * 8192, &maps.p1, p1_full, kb * 64, 0, b);
    mbar_expect_tx(p2_full, 32768);
    tma_load_3d(sP2, &maps.p2, p2_full, 0, 0, b);
  }
  sG2[tid] = __ldg(g2 + tid);
  sB2[tid] = __ldg(b2 + tid);

  // ---- ROI tile -> A operand: row = bin, 16-byte chunk c of the 256 channels -> k-block c/8, chunk (c%8)^(row%8)
  auto roi_dst = [&](int row, int c) -> uint8_t* {
    return sRoi + (c >> 3) * 8192 + row * 128 + (((c & 7) ^ (row & 7)) << 4);
  };
  if (kRoiFromGlobal) {
    const uint4* r = reinterpret_cast<const uint4*>(roi_in + static_cast<long>(b) * NBIN * D);
    for (int id = tid; id < NBIN * 32; id += 256) *reinterpret_cast<uint4*>(roi_dst(id >> 5, id & 31)) = __ldg(r + id);
  } else {
    __shared__ __align__(16) TapTable taps;
    const RoiGeom gm = roi_geometry(lv, boxes, b, boxes_per_frame);
    build_tap_table(gm, taps, tid);
    __syncthreads();
    for (int bin = warp; bin < NBIN; bin += 8) {
      float acc[8];
      roi_bin_table<false>(gm, taps, bin, lane, acc);
      *reinterpret_cast<uint4*>(roi_dst(bin, lane)) = pack8(acc);
    }
  }
  fence_proxy_async_smem();     // the tensor core reads the tile through the async proxy
  __syncthreads();

  const bool epi = (warp & 3) < 2;            // warps 0,1,4,5: TMEM lanes 0..63
  const int q = warp & 3;                     // TMEM sub-partition
  const int half = warp >> 2;                 // column half owned by this thread
  const int r = q * 32 + lane;                // row of the tile == TMEM lane (epilogue warps only)
  const uint32_t tlane = tmem_base + (static_cast<uint32_t>(q * 32) << 16);

  if (warp == 1) {
    // ===================== MMA issue: bmm1 =====================
    mbar_wait(p1_full, 0);
    tc_fence_after();
    if (elect_one()) {
      const uint32_t idesc = umma_idesc_f16(128, DD);
#pragma unroll
      for (int kb = 0; kb < 4; ++kb) {
        const uint64_t adesc = umma_desc_sw128_kmajor(smem_u32(sRoi + kb * 8192));
        const uint64_t bdesc = umma_desc_sw128_kmajor(smem_u32(sP1 + kb * 8192));
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) ? 1u : 0u);
      }
      umma_commit(mma_bar);
    }
    __syncwarp();
  }
  if (epi) {
    // ===================== LayerNorm(64) + ReLU -> F1 operand =====================
    mbar_wait(mma_bar, 0);
    tc_fence_after();
    uint32_t v[32];
    tmem_ld32(tlane + half * 32, v);
    tmem_ld_wait();
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = __uint_as_float(v[j]);
      sum += a;
      sq = fmaf(a, a, sq);
    }
    sStat[half * 64 + r] = make_float2(sum, sq);
    named_bar_sync(1, 128);
    const float2 other = sStat[(half ^ 1) * 64 + r];
    const float mean = (sum + other.x) * (1.f / DD);
    const float rstd = rsqrtf(fmaxf((sq + other.y) * (1.f / DD) - mean * mean, 0.f) + 1e-5f);
    uint8_t* rowp = sF1 + r * 128;
#pragma unroll
    for (int c4 = 0; c4 < 4; ++c4) {
      float y[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const int col = half * 32 + c4 * 8 + e;
        y[e] = fmaxf((__uint_as_float(v[c4 * 8 + e]) - mean) * rstd * __ldg(g1 + col) + __ldg(b1 + col), 0.f);
      }
      uint4 pk;
      pk.x = pack_half2(y[0], y[1]);
      pk.y = pack_half2(y[2], y[3]);
      pk.z = pack_half2(y[4], y[5]);
      pk.w = pack_half2(y[6], y[7]);
      *reinterpret_cast<uint4*>(rowp + (((half * 4 + c4) ^ (r & 7)) << 4)) = pk;
    }
    fence_proxy_async_smem();
    tc_fence_before();
    mbar_arrive(f1_ready);
  }
  if (warp == 1) {
    // ===================== MMA issue: bmm2 (the F1 accumulator columns are drained: f1_ready) =====================
    mbar_wait(f1_ready, 0);
    mbar_wait(p2_full, 0);
    tc_fence_after();
    if (elect_one()) {
      // two N=128 halves. The second one reads the F1 tile through a view that starts 8 KB earlier, so its 64 real rows
      // are rows 64..127 of the M=128 tile and land in TMEM lanes 64..127: all four sub-partitions hold one half of
      // the output and all eight warps share the LayerNorm(256) epilogue.
      const uint32_t idesc = umma_idesc_f16(128, D / 2);
      const uint64_t adesc0 = umma_desc_sw128_kmajor(smem_u32(sF1));
      const uint64_t bdesc0 = umma_desc_sw128_kmajor(smem_u32(sP2));
      const uint64_t adesc1 = umma_desc_sw128_kmajor(smem_u32(sF1 - 8192));
      const uint64_t bdesc1 = umma_desc_sw128_kmajor(smem_u32(sP2 + 128 * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base, adesc0 + 2 * k, bdesc0 + 2 * k, idesc, k ? 1u : 0u);
#pragma unroll
      for (int k = 0; k < 4; ++k) umma_f16(tmem_base + 128, adesc1 + 2 * k, bdesc1 + 2 * k, idesc, k ? 1u : 0u);
      umma_commit(mma_bar);
    }
    __syncwarp();
  }
  {
    // ===================== LayerNorm(256) + ReLU -> staging tile (the ROI buffer, plain swizzled rows) =====================
    // warp w: sub-partition q = w % 4 -> output columns 128*(q/2) .. +127 of rows 32*(q%2) .. +31; 64-column part w / 4.
    if (!epi) mbar_wait(mma_bar, 0);
    mbar_wait(mma_bar, 1);
    tc_fence_after();
    const int side = q >> 1, part = warp >> 2;
    const int row = (q & 1) * 32 + lane;
    const int col0 = side * 128 + part * 64;
    uint32_t v[2][32];
    tmem_ld32(tlane + col0, v[0]);
    tmem_ld32(tlane + col0 + 32, v[1]);
    tmem_ld_wait();
    float sum = 0.f, sq = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float a = __uint_as_float(v[0][j]), c = __uint_as_float(v[1][j]);
      sum += a + c;
      sq = fmaf(a, a, sq);
      sq = fmaf(c, c, sq);
    }
    sStat[(side * 2 + part) * 64 + row] = make_float2(sum, sq);
    __syncthreads();
    float tsum = 0.f, tsq = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 o = sStat[i * 64 + row];
      tsum += o.x;
      tsq += o.y;
    }
    const float mean = tsum * (1.f / D);
    const float rstd = rsqrtf(fmaxf(tsq * (1.f / D) - mean * mean, 0.f) + 1e-5f);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const int col = col0 + h * 32 + c4 * 8;
        const float4 ga = *reinterpret_cast<const float4*>(sG2 + col), gb = *reinterpret_cast<const float4*>(sG2 + col + 4);
        const float4 ba = *reinterpret_cast<const float4*>(sB2 + col), bb = *reinterpret_cast<const float4*>(sB2 + col + 4);
        const float gg[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float be[8] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float y[8];
#pragma unroll
        for (int e = 0; e < 8; ++e)
          y[e] = fmaxf((__uint_as_float(v[h][c4 * 8 + e]) - mean) * rstd * gg[e] + be[e], 0.f);
        uint4 pk;
        pk.x = pack_half2(y[0], y[1]);
        pk.y = pack_half2(y[2], y[3]);
        pk.z = pack_half2(y[4], y[5]);
        pk.w = pack_half2(y[6], y[7]);
        *reinterpret_cast<uint4*>(sRoi + off512(row, col >> 3)) = pk;
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  {
    __half* o = out + static_cast<long>(b) * NBIN * D;
    for (int id = tid; id < NBIN * 32; id += 256) {
      const uint4 v = *reinterpret_cast<const uint4*>(sRoi + off512(id >> 5, id & 31));
      *reinterpret_cast<uint4*>(o + id * 8) = v;
    }
  }
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

RoiLevels make_levels(const void* const* feats, const int* hs, const int* ws, const float* scales) {
  RoiLevels lv;
  for (int i = 0; i < 3; ++i) {
    lv.feat[i] = static_cast<const __half*>(feats[i]);
    lv.h[i] = hs[i];
    lv.w[i] = ws[i];
    lv.scale[i] = scales[i];
  }
  return lv;
}

}  // namespace

int roi_align_launch(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                     int num_boxes, int boxes_per_frame, void* roi_out, float* mean_f32, void* mean_f16,
                     cudaStream_t stream) {
  if (num_boxes <= 0 || boxes_per_frame <= 0) return DVID_ERR_SHAPE;
  launch_pdl(roi_align_kernel, dim3(num_boxes), dim3(256), 0, stream, make_levels(feats, hs, ws, scales), boxes, boxes_per_frame,
                                                  static_cast<__half*>(roi_out), mean_f32,
                                                  static_cast<__half*>(mean_f16));
  return check_launch();
}

int roi_dynconv_launch(const void* const* feats, const int* hs, const int* ws, const float* scales,
                       const float* boxes, int num_boxes, int boxes_per_frame, const void* roi_in, const void* params,
                       const float* g1, const float* b1, const float* g2, const float* b2, void* out,
                       cudaStream_t stream) {
  if (num_boxes <= 0 || boxes_per_frame <= 0) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(roi_dynconv_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) !=
            cudaSuccess ||
        cudaFuncSetAttribute(roi_dynconv_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) !=
            cudaSuccess ||
        cudaFuncSetAttribute(roi_dynconv_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, SM_TOTAL) !=
            cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  static int half_w = -1;
  if (half_w < 0) { const char* e = getenv("DVID_ROI_F16W"); half_w = e ? atoi(e) : 0; }
  const RoiLevels lv = make_levels(feats, hs, ws, scales);
  if (roi_in == nullptr && half_w) {
    launch_pdl(roi_dynconv_kernel<false, true>, dim3(num_boxes), dim3(256), SM_TOTAL, stream,
        lv, boxes, boxes_per_frame, nullptr, static_cast<const __half*>(params), g1, b1, g2, b2,
        static_cast<__half*>(out));
    return check_launch();
  }
  if (roi_in != nullptr) {
    launch_pdl(roi_dynconv_kernel<true>, dim3(num_boxes), dim3(256), SM_TOTAL, stream, 
        lv, boxes, boxes_per_frame, static_cast<const __half*>(roi_in), static_cast<const __half*>(params), g1, b1,
        g2, b2, static_cast<__half*>(out));
  } else {
    launch_pdl(roi_dynconv_kernel<false>, dim3(num_boxes), dim3(256), SM_TOTAL, stream, 
        lv, boxes, boxes_per_frame, nullptr, static_cast<const __half*>(params), g1, b1, g2, b2,
        static_cast<__half*>(out));
  }
  return check_launch();
}


// params: [M][32768] fp16 in the TRANSPOSED layout (P1^T [64][256] then P2^T [256][64] per box), see the kernel comment.
int roi_dynconv_tc_launch(const void* const* feats, const int* hs, const int* ws, const float* scales,
                          const float* boxes, int num_boxes, int boxes_per_frame, const void* roi_in,
                          const void* params_t, const float* g1, const float* b1, const float* g2, const float* b2,
                          void* out, cudaStream_t stream) {
  if (num_boxes <= 0 || boxes_per_frame <= 0) return DVID_ERR_SHAPE;
  static bool attr_set = false;
  if (!attr_set) {
    if (cudaFuncSetAttribute(roi_dynconv_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_TOTAL) !=
            cudaSuccess ||
        cudaFuncSetAttribute(roi_dynconv_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_TOTAL) !=
            cudaSuccess)
      return DVID_ERR_CUDA;
    attr_set = true;
  }
  DynMaps maps;
  {
    const uint64_t dims[3] = {256, 64, static_cast<uint64_t>(num_boxes)};
    const uint64_t strides[2] = {512, 65536};
    const uint32_t box[3] = {64, 64, 1};
    int r = make_tmap_f16(&maps.p1, params_t, 3, dims, strides, box, nullptr);
    if (r) return r;
  }
  {
    const uint64_t dims[3] = {64, 256, static_cast<uint64_t>(num_boxes)};
    const uint64_t strides[2] = {128, 65536};
    const uint32_t box[3] = {64, 256, 1};
    int r = make_tmap_f16(&maps.p2, static_cast<const __half*>(params_t) + D * DD, 3, dims, strides, box, nullptr);
    if (r) return r;
  }
  const RoiLevels lv = make_levels(feats, hs, ws, scales);
  if (roi_in != nullptr)
    launch_pdl(roi_dynconv_tc_kernel<true>, dim3(num_boxes), dim3(256), TC_TOTAL, stream, maps, lv, boxes,
               boxes_per_frame, static_cast<const __half*>(roi_in), g1, b1, g2, b2, static_cast<__half*>(out));
  else
    launch_pdl(roi_dynconv_tc_kernel<false>, dim3(num_boxes), dim3(256), TC_TOTAL, stream, maps, lv, boxes,
               boxes_per_frame, static_cast<const __half*>(nullptr), g1, b1, g2, b2, static_cast<__half*>(out));
  return check_launch();
}

}  // namespace dvid

// Clip loader, first stage: the frame resize of the reference's test transform on the GPU.
//
// Reference: Resize(min_size, max_size) (mega_core/data/transforms/transforms.py:31-67) -> torchvision F.resize on a
// PIL image -> Pillow Image.resize(size, BILINEAR): an antialiased two-pass (horizontal, then vertical) triangle
// filter in 8-bit fixed point (Pillow libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc,
// ImagingResampleHorizontal_8bpc / Vertical_8bpc; restated and pinned against the installed Pillow in
// oracle/resize.py + tests/test_oracle_resize.py).  The output must be the same BYTES as Pillow's, so
//   * the per-coordinate filter windows and weights are computed on the device in IEEE double precision with explicit
//     round-to-nearest operations in Pillow's order (no fused multiply-add), normalised, and rounded to 22-bit fixed
//     point exactly like normalize_coeffs_8bpc;
//   * each pass accumulates from 2^21 in 32-bit integers, shifts right by 22 and clamps to [0, 255]; the horizontal
//     pass's rounded bytes are what the vertical pass reads.
// Input: decoded frames [n][Hin][Win][3] uint8 (PIL / HWC order).  Output: [n][3][Hp][Wp] uint8 planes, the resized
// oh x ow image in the top-left corner and zeros elsewhere - what pil_to_tensor + to_image_list(size_divisible=32)
// hand to the model (mega_core/structures/image_list.py:36-66) - ready for dvid_preprocess_u8.
// Byte work, a few MB per frame, bound by load/store INSTRUCTIONS rather than bytes if done one byte at a time (the
// first version: 15 byte loads + 3 byte stores per thread and pass, 69 us for eight 720p frames).  Now the source row is
// read as aligned 32-bit words and pixels are cut out with funnel shifts, the horizontal pass stores one RGBX word per
// pixel, and the vertical pass handles four pixels per thread with 16-byte loads and one 4-byte store per plane.
#include "dvid_internal.h"

namespace dvid {

namespace {

constexpr int PRECISION_BITS = 32 - 8 - 2;
constexpr int MAX_KSIZE = 64;      // scale factors up to ~31

__host__ __device__ inline int resize_ksize(int in_size, int out_size) {
  // ksize = (int)ceil(support) * 2 + 1 with support = max(in/out, 1): ceil of a ratio of integers, exactly
  const int s = in_size > out_size ? (in_size + out_size - 1) / out_size : 1;
  return s * 2 + 1;
}

// precompute_coeffs + normalize_coeffs_8bpc for one axis; one thread per output coordinate.
// Both axes in one launch: threads [0, out_h) do the horizontal table, [out_h, out_h + out_v) the vertical one.
__global__ void resize_coeffs_kernel(int in_h, int out_h, int ks_h, int* __restrict__ bounds_h,
                                     int* __restrict__ coeffs_h, int in_v, int out_v, int ks_v,
                                     int* __restrict__ bounds_v, int* __restrict__ coeffs_v) {
  int xx = blockIdx.x * blockDim.x + threadIdx.x;
  const bool vert = xx >= out_h;
  if (vert) xx -= out_h;
  const int in_size = vert ? in_v : in_h, out_size = vert ? out_v : out_h, ksize = vert ? ks_v : ks_h;
  int* bounds = vert ? bounds_v : bounds_h;
  int* coeffs = vert ? coeffs_v : coeffs_h;
  if (xx >= out_size) return;
  const double scale = __ddiv_rn(static_cast<double>(in_size), static_cast<double>(out_size));
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = fscale;                      // bilinear: filter support 1.0 * filterscale
  const double ss = __ddiv_rn(1.0, fscale);
  const double center = __dmul_rn(__dadd_rn(static_cast<double>(xx), 0.5), scale);
  int xmin = static_cast<int>(__dadd_rn(__dsub_rn(center, support), 0.5));
  if (xmin < 0) xmin = 0;
  int xmax = static_cast<int>(__dadd_rn(__dadd_rn(center, support), 0.5));
  if (xmax > in_size) xmax = in_size;
  xmax -= xmin;
  double w[MAX_KSIZE];
  double ww = 0.0;
  for (int x = 0; x < ksize; ++x) {
    double v = 0.0;
    if (x < xmax) {
      double a = __dmul_rn(__dadd_rn(__dsub_rn(static_cast<double>(x + xmin), center), 0.5), ss);
      if (a < 0.0) a = -a;
      v = a < 1.0 ? __dsub_rn(1.0, a) : 0.0;
      ww = __dadd_rn(ww, v);
    }
    w[x] = v;
  }
  for (int x = 0; x < ksize; ++x) {
    double v = w[x];
    if (x < xmax && ww != 0.0) v = __ddiv_rn(v, ww);
    // (int)(0.5 + k * 2^22): weights of the triangle filter are never negative
    coeffs[xx * ksize + x] = static_cast<int>(__dadd_rn(0.5, __dmul_rn(v, static_cast<double>(1 << PRECISION_BITS))));
  }
  bounds[2 * xx] = xmin;
  bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ unsigned char clip8(int acc) {
  const int v = acc >> PRECISION_BITS;
  return static_cast<unsigned char>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: src [n][H][Win][3] bytes -> tmp [n][H][ow] RGBX words.  The taps of an output pixel are the
// contiguous bytes [3 x0, 3 (x0 + cnt)) of the row: aligned 32-bit loads, pixel t cut out by a funnel shift.
__global__ void resize_horizontal_kernel(const unsigned char* __restrict__ src, uint32_t* __restrict__ tmp, int H,
                                         int Win, int ow, int ow4, int ksize, const int* __restrict__ bounds,
                                         const int* __restrict__ coeffs) {
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  if (xo >= ow4) return;
  const int y = blockIdx.y, im = blockIdx.z;
  if (xo >= ow) {                             // alignment columns of tmp: defined, never used
    tmp[(static_cast<long>(im) * H + y) * ow4 + xo] = 0u;
    return;
  }
  const int x0 = bounds[2 * xo], cnt = bounds[2 * xo + 1];
  const long row_byte = (static_cast<long>(im) * H + y) * Win * 3;       // rows start at any byte alignment
  const long first = row_byte + static_cast<long>(x0) * 3;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(src) + (first >> 2);
  const long last_word = (row_byte + static_cast<long>(Win) * 3 - 1) >> 2;   // never read past the row's last word
  const long w0 = first >> 2;
  int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
  uint32_t lo = __ldg(words);
  int wi = 0;                                 // index of `lo` relative to w0
  uint32_t hi = (w0 + 1 <= last_word) ? __ldg(words + 1) : 0u;
  int off = static_cast<int>(first & 3);      // byte offset of the current pixel inside `lo`
  for (int t = 0; t < cnt; ++t) {
    const uint32_t px = __funnelshift_r(lo, hi, off * 8);     // bytes R, G, B of pixel x0 + t in the low 24 bits
    const int k = coeffs[xo * ksize + t];
    a0 += static_cast<int>(px & 0xffu) * k;
    a1 += static_cast<int>((px >> 8) & 0xffu) * k;
    a2 += static_cast<int>((px >> 16) & 0xffu) * k;
    off += 3;
    if (off >= 4) {
      off -= 4;
      lo = hi;
      ++wi;
      hi = (w0 + wi + 1 <= last_word) ? __ldg(words + wi + 1) : 0u;
    }
  }
  tmp[(static_cast<long>(im) * H + y) * ow4 + xo] =
      static_cast<uint32_t>(clip8(a0)) | (static_cast<uint32_t>(clip8(a1)) << 8) |
      (static_cast<uint32_t>(clip8(a2)) << 16);
}

// Fast path of the horizontal pass for KS <= 5 taps (every scale factor up to 2, e.g. 720p -> 562 x 999): the <= 18
// source bytes of the window are fetched as five aligned words, re-aligned to the window start with four funnel shifts,
// and every tap is cut out at a compile-time position: ~60 instead of ~230 instructions per pixel (the generic loop
// above is bound by instruction issue: 46 us for eight 720p frames at 79 % SM throughput, ncu).
template <int KS>
__global__ void resize_horizontal_fast_kernel(const unsigned char* __restrict__ src, uint32_t* __restrict__ tmp, int H,
                                              int Win, int ow, int ow4, const int* __restrict__ bounds,
                                              const int* __restrict__ coeffs) {
  constexpr int NW = (3 + 3 * (KS - 1)) / 4 + 2;       // words that can hold the window at any alignment, + 1
  const int xo = blockIdx.x * blockDim.x + threadIdx.x;
  if (xo >= ow4) return;
  const int y = blockIdx.y, im = blockIdx.z;
  const int orow = (im * H + y) * ow4;
  if (xo >= ow) {
    tmp[orow + xo] = 0u;
    return;
  }
  const int x0 = bounds[2 * xo], cnt = bounds[2 * xo + 1];
  const int row_byte = (im * H + y) * Win * 3;          // the launcher checks that the frame batch is < 2^31 bytes
  const int first = row_byte + x0 * 3;
  const int w0 = first >> 2;
  const int last_word = (row_byte + Win * 3 - 1) >> 2;
  const uint32_t* words = reinterpret_cast<const uint32_t*>(src);
  uint32_t wd[NW];
#pragma unroll
  for (int i = 0; i < NW; ++i) wd[i] = (w0 + i <= last_word) ? __ldg(words + w0 + i) : 0u;
  const int sh = (first & 3) * 8;
  uint32_t al[NW - 1];                                  // the window as a byte stream starting at byte 0 of al[0]
#pragma unroll
  for (int i = 0; i < NW - 1; ++i) al[i] = __funnelshift_r(wd[i], wd[i + 1], sh);
  int a0 = 1 << (PRECISION_BITS - 1), a1 = a0, a2 = a0;
#pragma unroll
  for (int t = 0; t < KS; ++t) {
    if (t < cnt) {
      const int b = 3 * t;
      const uint32_t px = ((b & 3) == 0) ? al[b >> 2]
                                         : __funnelshift_r(al[b >> 2], al[(b >> 2) + 1 < NW - 1 ? (b >> 2) + 1 : b >> 2],
                                                           (b & 3) * 8);
      const int k = coeffs[xo * KS + t];
      a0 += static_cast<int>(px & 0xffu) * k;
      a1 += static_cast<int>((px >> 8) & 0xffu) * k;
      a2 += static_cast<int>((px >> 16) & 0xffu) * k;
    }
  }
  tmp[orow + xo] = static_cast<uint32_t>(clip8(a0)) | (static_cast<uint32_t>(clip8(a1)) << 8) |
                   (static_cast<uint32_t>(clip8(a2)) << 16);
}

// vertical pass + RGBX -> planar + zero padding: tmp [n][Hin][ow4] words -> dst [n][3][Hp][Wp]; four pixels per thread
// (ow4 = ow rounded up to 4 so that every row of tmp is 16-byte aligned; Wp % 4 == 0)
__global__ void resize_vertical_kernel(const uint32_t* __restrict__ tmp, unsigned char* __restrict__ dst, int Hin,
                                       int oh, int ow, int ow4, int Hp, int Wp, int ksize,
                                       const int* __restrict__ bounds, const int* __restrict__ coeffs) {
  const int xq = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (xq >= Wp) return;
  const int yo = blockIdx.y, im = blockIdx.z;
  uint32_t r[3] = {0u, 0u, 0u};
  if (yo < oh && xq < ow) {
    const int y0 = bounds[2 * yo], cnt = bounds[2 * yo + 1];
    int acc[4][3];
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = 1 << (PRECISION_BITS - 1);
    for (int y = 0; y < cnt; ++y) {
      const int k = coeffs[yo * ksize + y];
      const uint4 v = __ldg(reinterpret_cast<const uint4*>(tmp + (static_cast<long>(im) * Hin + y0 + y) * ow4 + xq));
      const uint32_t px[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        acc[i][0] += static_cast<int>(px[i] & 0xffu) * k;
        acc[i][1] += static_cast<int>((px[i] >> 8) & 0xffu) * k;
        acc[i][2] += static_cast<int>((px[i] >> 16) & 0xffu) * k;
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (xq + i < ow) {
#pragma unroll
        for (int c = 0; c < 3; ++c) r[c] |= static_cast<uint32_t>(clip8(acc[i][c])) << (8 * i);
      }
    }
  }
  const long plane = static_cast<long>(Hp) * Wp;
  unsigned char* o = dst + static_cast<long>(im) * 3 * plane + static_cast<long>(yo) * Wp + xq;
  *reinterpret_cast<uint32_t*>(o) = r[0];
  *reinterpret_cast<uint32_t*>(o + plane) = r[1];
  *reinterpret_cast<uint32_t*>(o + 2 * plane) = r[2];
}

inline long align256(long v) { return (v + 255) & ~255L; }

}  // namespace

long resize_workspace_bytes(int n, int Hin, int Win, int oh, int ow) {
  if (n <= 0 || Hin <= 0 || Win <= 0 || oh <= 0 || ow <= 0) return -1;
  const int kh = resize_ksize(Win, ow), kv = resize_ksize(Hin, oh);
  const long ow4 = (ow + 3) & ~3;
  return align256(static_cast<long>(n) * Hin * ow4 * 4) + align256(static_cast<long>(ow) * (kh + 2) * 4) +
         align256(static_cast<long>(oh) * (kv + 2) * 4);
}

int resize_bilinear_u8_launch(const unsigned char* src, int n, int Hin, int Win, int oh, int ow, unsigned char* dst,
                              int Hp, int Wp, void* workspace, long workspace_bytes, cudaStream_t stream) {
  if (n <= 0 || Hin <= 0 || Win <= 0 || oh <= 0 || ow <= 0 || Hp < oh || Wp < ow) return DVID_ERR_SHAPE;
  if (Hin > 65535 || Hp > 65535 || n > 65535 || (Wp & 3) != 0) return DVID_ERR_SHAPE;
  // aligned word access: frame buffer and output on 4-byte, workspace on 16-byte boundaries; the source is read in
  // whole words, i.e. up to 3 bytes past an odd-sized buffer's end inside its last aligned word
  if ((reinterpret_cast<uintptr_t>(src) & 3) != 0 || (reinterpret_cast<uintptr_t>(dst) & 3) != 0 ||
      (reinterpret_cast<uintptr_t>(workspace) & 15) != 0)
    return DVID_ERR_ARG;
  const int kh = resize_ksize(Win, ow), kv = resize_ksize(Hin, oh);
  if (kh > MAX_KSIZE || kv > MAX_KSIZE) return DVID_ERR_SHAPE;
  if (workspace_bytes < resize_workspace_bytes(n, Hin, Win, oh, ow)) return DVID_ERR_ARG;
  const int ow4 = (ow + 3) & ~3;
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  uint32_t* tmp = reinterpret_cast<uint32_t*>(ws);
  ws += align256(static_cast<long>(n) * Hin * ow4 * 4);
  int* ch = reinterpret_cast<int*>(ws);
  int* bh = ch + static_cast<long>(ow) * kh;
  ws += align256(static_cast<long>(ow) * (kh + 2) * 4);
  int* cv = reinterpret_cast<int*>(ws);
  int* bv = cv + static_cast<long>(oh) * kv;
  resize_coeffs_kernel<<<(ow + oh + 127) / 128, 128, 0, stream>>>(Win, ow, kh, bh, ch, Hin, oh, kv, bv, cv);
  const dim3 hgrid((ow4 + 255) / 256, Hin, n);
  const bool small = static_cast<long>(n) * Hin * Win * 3 < (1L << 31) && static_cast<long>(n) * Hin * ow4 < (1L << 31);
  if (small && kh == 3)
    resize_horizontal_fast_kernel<3><<<hgrid, 256, 0, stream>>>(src, tmp, Hin, Win, ow, ow4, bh, ch);
  else if (small && kh == 5)
    resize_horizontal_fast_kernel<5><<<hgrid, 256, 0, stream>>>(src, tmp, Hin, Win, ow, ow4, bh, ch);
  else
    resize_horizontal_kernel<<<hgrid, 256, 0, stream>>>(src, tmp, Hin, Win, ow, ow4, kh, bh, ch);
  resize_vertical_kernel<<<dim3((Wp / 4 + 127) / 128, Hp, n), 128, 0, stream>>>(tmp, dst, Hin, oh, ow, ow4, Hp, Wp, kv,
                                                                                 bv, cv);
  return check_launch();
}

}  // namespace dvid

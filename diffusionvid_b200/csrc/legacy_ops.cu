// Legacy maskrcnn-benchmark operators of mega_core._C that are NOT on the DiffusionVID hot path but belong to the
// native module the package replaces: roi_align_forward.
//
// Reference: mega_core/csrc/cuda/ROIAlign_cuda.cu:15-62 (bilinear_interpolate), :65-125 (RoIAlignForward), launcher
// :256-300.  Semantics that differ from the detector's own ROIAlign (torchvision aligned=True, roi_dynconv.cu): no -0.5
// pixel shift, roi width/height clamped to >= 1, NCHW fp32 in and out, rois (n,5) = [batch index, x1, y1, x2, y2],
// sampling_ratio <= 0 means an adaptive grid of ceil(roi_size / pooled_size) samples per bin.
//
// The reference runs one thread per output element (n,c,ph,pw).  Here a warp owns one (roi, bin) and its lanes stride
// over the channels: the sample coordinates and the four tap weights of a bin are computed once per sample instead
// of once per channel, and the output is written [n][c][ph][pw] like the reference's.  Integer/fp32 SIMT work bound by
// the uncoalesced NCHW taps (one 4-byte tap per lane per plane); it exists for API completeness, not speed.
#include "dvid_internal.h"

namespace dvid {

namespace {

__global__ void __launch_bounds__(256)
roi_align_legacy_kernel(const float* __restrict__ in, const float* __restrict__ rois, int num_rois, int C, int H,
                        int W, float scale, int PH, int PW, int sampling_ratio, float* __restrict__ out) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const long warp = (static_cast<long>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const long nwarps = (static_cast<long>(gridDim.x) * blockDim.x) >> 5;
  const long total = static_cast<long>(num_rois) * PH * PW;
  for (long item = warp; item < total; item += nwarps) {
    const int pw = static_cast<int>(item % PW);
    const int ph = static_cast<int>((item / PW) % PH);
    const int n = static_cast<int>(item / (static_cast<long>(PW) * PH));
    const float* r = rois + static_cast<long>(n) * 5;
    const int b = static_cast<int>(r[0]);
    const float x1 = __fmul_rn(r[1], scale), y1 = __fmul_rn(r[2], scale);
    const float x2 = __fmul_rn(r[3], scale), y2 = __fmul_rn(r[4], scale);
    const float rw = fmaxf(__fsub_rn(x2, x1), 1.0f), rh = fmaxf(__fsub_rn(y2, y1), 1.0f);
    const float bh = __fdiv_rn(rh, static_cast<float>(PH)), bw = __fdiv_rn(rw, static_cast<float>(PW));
    const int gh = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(__fdiv_rn(rh, static_cast<float>(PH))));
    const int gw = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(__fdiv_rn(rw, static_cast<float>(PW))));
    const float count = static_cast<float>(gh * gw);
    const float* base = in + static_cast<long>(b) * C * H * W;
    for (int c = lane; c < C; c += 32) {
      const float* plane = base + static_cast<long>(c) * H * W;
      float acc = 0.f;
      for (int iy = 0; iy < gh; ++iy) {
        float y = __fadd_rn(__fadd_rn(y1, __fmul_rn(static_cast<float>(ph), bh)),
                            __fdiv_rn(__fmul_rn(static_cast<float>(iy) + 0.5f, bh), static_cast<float>(gh)));
        for (int ix = 0; ix < gw; ++ix) {
          float x = __fadd_rn(__fadd_rn(x1, __fmul_rn(static_cast<float>(pw), bw)),
                              __fdiv_rn(__fmul_rn(static_cast<float>(ix) + 0.5f, bw), static_cast<float>(gw)));
          float yy = y;
          if (yy < -1.0f || yy > static_cast<float>(H) || x < -1.0f || x > static_cast<float>(W)) continue;
          if (yy <= 0.f) yy = 0.f;
          if (x <= 0.f) x = 0.f;
          int yl = static_cast<int>(yy), xl = static_cast<int>(x), yh, xh;
          if (yl >= H - 1) { yh = yl = H - 1; yy = static_cast<float>(yl); } else { yh = yl + 1; }
          if (xl >= W - 1) { xh = xl = W - 1; x = static_cast<float>(xl); } else { xh = xl + 1; }
          const float ly = __fsub_rn(yy, static_cast<float>(yl)), lx = __fsub_rn(x, static_cast<float>(xl));
          const float hy = __fsub_rn(1.0f, ly), hx = __fsub_rn(1.0f, lx);
          const float v1 = __ldg(plane + yl * W + xl), v2 = __ldg(plane + yl * W + xh);
          const float v3 = __ldg(plane + yh * W + xl), v4 = __ldg(plane + yh * W + xh);
          // (w1*v1 + w2*v2 + w3*v3 + w4*v4), left to right like ROIAlign_cuda.cu:58
          const float val = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(hy, hx), v1),
                                                          __fmul_rn(__fmul_rn(hy, lx), v2)),
                                                __fmul_rn(__fmul_rn(ly, hx), v3)),
                                      __fmul_rn(__fmul_rn(ly, lx), v4));
          acc = __fadd_rn(acc, val);
        }
      }
      out[((static_cast<long>(n) * C + c) * PH + ph) * PW + pw] = __fdiv_rn(acc, count);
    }
  }
}

}  // namespace

int roi_align_legacy_launch(const float* in, const float* rois, int num_rois, int C, int H, int W, float scale, int PH,
                            int PW, int sampling_ratio, float* out, cudaStream_t stream) {
  if (num_rois < 0 || C <= 0 || H <= 0 || W <= 0 || PH <= 0 || PW <= 0) return DVID_ERR_SHAPE;
  if (num_rois == 0) return DVID_OK;
  const long items = static_cast<long>(num_rois) * PH * PW;      // one warp each
  long blocks = (items + 7) / 8;
  const long cap = static_cast<long>(num_sms()) * 16;
  if (blocks > cap) blocks = cap;
  launch_pdl(roi_align_legacy_kernel, dim3(static_cast<unsigned>(blocks)), dim3(256), 0, stream, in, rois, num_rois, C,
             H, W, scale, PH, PW, sampling_ratio, out);
  return check_launch();
}

}  // namespace dvid

// extern "C" surface of libdvid_b200.so: argument validation + dispatch to the kernel launchers. See include/dvid_b200.h.
#include "../../include/dvid_b200.h"
#include "dvid_internal.h"

#define DVID_ABI_VERSION 1

extern "C" {

int dvid_abi_version(void) { return DVID_ABI_VERSION; }
int dvid_num_sms(void) { return dvid::num_sms(); }

int dvid_conv2d_nhwc_f16(const void* in, const void* weight, const float* bias, const void* resid, void* out, int n,
                         int h, int w, int cin, int cout, int R, int S, int stride, int pad, int resid_shift,
                         int relu, void* stream) {
  if (!in || !weight || !out) return DVID_ERR_ARG;
  return dvid::conv_gemm_launch(in, weight, bias, resid, out, nullptr, n, h, w, cin, cout, R, S, stride, pad,
                                resid_shift, relu, 1, 0, static_cast<cudaStream_t>(stream));
}

int dvid_gemm_f16(const void* a, const void* w, const float* bias, const void* resid, void* out_f16,
                  float* out_f32_partials, int m, int n, int k, int relu, int splits, int* splits_used,
                  void* stream) {
  if (!a || !w || (!out_f16 && !out_f32_partials)) return DVID_ERR_ARG;
  if (m <= 0 || n <= 0 || k <= 0) return DVID_ERR_SHAPE;
  int used = 1;
  if (out_f32_partials) {
    const int total_kb = (k + 63) / 64;
    int s = splits < 1 ? 1 : (splits > total_kb ? total_kb : splits);
    const int per = (total_kb + s - 1) / s;
    used = (total_kb + per - 1) / per;
  }
  if (splits_used) *splits_used = used;
  return dvid::conv_gemm_launch(a, w, bias, resid, out_f32_partials ? nullptr : out_f16, out_f32_partials, 1, 1, m, k,
                                n, 1, 1, 1, 0, 0, relu, splits, 0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

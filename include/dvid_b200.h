/*
 * dvid_b200 — C ABI of the B200-native DiffusionVID inference hot path (libdvid_b200.so).
 *
 * Every entry point takes plain device pointers, sizes and a CUDA stream (cudaStream_t passed as void*); no torch
 * types cross this boundary. All functions are asynchronous on `stream`, never synchronise, never allocate device
 * memory, and return 0 on success or a DVID_ERR_* code (they never exit the process — the reference's FPS launcher
 * calls exit(-1) on a launch error, mega_core/csrc/cuda/fps.cu:181-185).
 *
 * Layout conventions: activations NHWC fp16 ("half"), weights [Cout][R*S*Cin] fp16, box coordinates / logits /
 * noise fp32, indices int32 unless stated. Citations are into /root/reference (sdroh1027/DiffusionVID @ 8375542).
 */
#ifndef DVID_B200_H
#define DVID_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define DVID_API __attribute__((visibility("default")))
#else
#define DVID_API
#endif

#define DVID_OK 0
#define DVID_ERR_SHAPE 1
#define DVID_ERR_CUDA 2
#define DVID_ERR_DRIVER 3
#define DVID_ERR_ARG 4

/* Library / build identification: returns the ABI version (bumped on any signature change). */
DVID_API int dvid_abi_version(void);
/* Number of SMs of the current device (grid sizing for persistent kernels). */
DVID_API int dvid_num_sms(void);

/* ---------------------------------------------------------------------------------------------------------------
 * Dense contractions (tcgen05 + TMA).
 *
 * dvid_conv2d_nhwc_f16 replaces the cuDNN convolutions behind detectron2's ResNet/FPN that the reference builds at
 * mega_core/modeling/detector/diffusion_det.py:219 and calls at :427 (FrozenBN folded into weight/bias by the host).
 *   out[n,y,x,co] = act( bias[co] + sum_{r,s,ci} in[n, y*stride+r-pad, x*stride+s-pad, ci] * weight[co,(r*S+s)*Cin+ci]
 *                        + resid[n, y>>resid_shift, x>>resid_shift, co] )
 * resid_shift=1 implements FPN's nearest x2 top-down addition. bias/resid may be NULL. relu: 0/1.
 * Requirements: Cin % 8 == 0, Cout % 8 == 0, stride in {1,2}.
 */
DVID_API int dvid_conv2d_nhwc_f16(const void* in, const void* weight, const float* bias, const void* resid, void* out,
                         int n, int h, int w, int cin, int cout, int R, int S, int stride, int pad,
                         int resid_shift, int relu, void* stream);

/* dvid_gemm_f16 replaces torch.nn.functional.linear (cuBLAS) for every nn.Linear of the decoder
 * (mega_core/modeling/roi_heads/box_head/box_head.py:218-223,447-491,675-684):
 *   out_f16[m,n] = act( bias[n] + sum_k a[m,k] * w[n,k] + resid[m,n] )           (out_f32_partials == NULL)
 *   out_f32_partials[s,m,n] = sum_{k in split s} a[m,k] * w[n,k]                 (split-K; reduce with dvid_row_post)
 * K % 8 == 0, N % 8 == 0. `splits` is a request; the number actually used is returned in *splits_used (may be NULL).
 */
DVID_API int dvid_gemm_f16(const void* a, const void* w, const float* bias, const void* resid, void* out_f16,
                  float* out_f32_partials, int m, int n, int k, int relu, int splits, int* splits_used,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVID_B200_H */

// Internal declarations shared by the .cu translation units of libdvid_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define DVID_OK 0
#define DVID_ERR_SHAPE 1   /* unsupported / inconsistent shape arguments */
#define DVID_ERR_CUDA 2    /* a CUDA runtime call or kernel launch failed */
#define DVID_ERR_DRIVER 3  /* driver entry point (tensor-map encode) unavailable or failed */
#define DVID_ERR_ARG 4     /* null pointer / bad enum */

namespace dvid {

int num_sms();

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides);

int conv_gemm_launch(const void* in, const void* weight, const float* bias, const void* resid, void* out,
                     float* out_f32, int n, int h, int w, int cin, int cout, int R, int S, int stride, int pad,
                     int resid_shift, int relu, int splits, int force_bn, cudaStream_t stream);

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? DVID_OK : DVID_ERR_CUDA; }

}  // namespace dvid

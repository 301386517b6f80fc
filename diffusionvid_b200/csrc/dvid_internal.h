// Internal declarations shared by the .cu translation units of libdvid_b200.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define DVID_OK 0
#define DVID_ERR_SHAPE 1   /* unsupported / inconsistent shape arguments */
#define DVID_ERR_CUDA 2    /* a CUDA runtime call or kernel launch failed */
#define DVID_ERR_DRIVER 3  /* driver entry point (tensor-map encode) unavailable or failed */
#define DVID_ERR_ARG 4     /* null pointer / bad enum */

namespace dvid {

int num_sms();
int conv_streamk_enable(int on);   // conv_gemm.cu: allocate the stream-K workspace (outside capture) / switch it

int make_tmap_f16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box, const uint32_t* elem_strides);

int conv_gemm_launch(const void* in, const void* weight, const float* bias, const void* resid, void* out,
                     float* out_f32, int n, int h, int w, int cin, int cout, int R, int S, int stride, int pad,
                     int resid_shift, int relu, int splits, int force_bn, cudaStream_t stream,
                     const uint64_t* a_strides_bytes = nullptr);
int stem_conv_launch(const void* in_haloed, const void* weight, const float* bias, void* out, int n, int H, int W,
                     int cout, int relu, cudaStream_t stream);

int attention_launch(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                     long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                     cudaStream_t stream);

int attention_tc_launch(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                        long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                        cudaStream_t stream);

int roi_align_launch(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                     int num_boxes, int boxes_per_frame, void* roi_out, float* mean_f32, void* mean_f16,
                     cudaStream_t stream);
int roi_dynconv_launch(const void* const* feats, const int* hs, const int* ws, const float* scales,
                       const float* boxes, int num_boxes, int boxes_per_frame, const void* roi_in, const void* params,
                       const float* g1, const float* b1, const float* g2, const float* b2, void* out,
                       cudaStream_t stream);

int roi_dynconv_tc_launch(const void* const* feats, const int* hs, const int* ws, const float* scales,
                          const float* boxes, int num_boxes, int boxes_per_frame, const void* roi_in,
                          const void* params_t, const float* g1, const float* b1, const float* g2, const float* b2,
                          void* out, cudaStream_t stream);

int preprocess_launch(const void* img, int is_u8, void* out, int n, int H, int W, int halo, int Hp, int Wp,
                      const float* mean, const float* std, cudaStream_t stream);
long resize_workspace_bytes(int n, int Hin, int Win, int oh, int ow);
int resize_bilinear_u8_launch(const unsigned char* src, int n, int Hin, int Win, int oh, int ow, unsigned char* dst,
                              int Hp, int Wp, void* workspace, long workspace_bytes, cudaStream_t stream);
int maxpool_launch(const void* in, void* out, int n, int H, int W, int C, cudaStream_t stream);
int row_post_launch(const float* partials, int splits, long split_stride, const void* in_f16, const float* bias,
                    const float* ln1_g, const float* ln1_b, int relu1, const float* resid, const float* ln2_g,
                    const float* ln2_b, int act2, int act2_f16_only, float* out_f32, void* out_f16,
                    const float* mod_scale, const float* mod_shift, int rows_per_group, int scale_stride,
                    int shift_stride, int shift_per_row, void* out_mod_f16, int M, cudaStream_t stream);
int small_linear_launch(const float* a, const void* w, const float* bias, float* out, int m, int n, int k, int act_in,
                        int act_out, cudaStream_t stream);
int time_sinusoid_launch(const float* t, const float* freq, float* out, int m, cudaStream_t stream);
int head_final_launch(const float* logit_part, int ldl, const float* cls_bias, int C, const float* delta_part, int ldd,
                      const float* delta_bias, const float* boxes_in, float* logits_out, float* boxes_out, int M,
                      cudaStream_t stream);
int noise_to_boxes_launch(const float* x, float* boxes, int M, float scale, float W, float H, cudaStream_t stream);
int ddim_step_launch(const float* logits, int C, const float* coord, const float* x_t, const float* eps,
                     const float* fill, float* x_next, float* boxes_next, int* num_kept, int frames, int N,
                     float scale, float W, float H, float sqrt_recip_a, float sqrt_recipm1_a, float sqrt_a_next,
                     float c_coef, float sigma, cudaStream_t stream);

int topk_scores_launch(const float* logits, const float* boxes, int frames, int N, int C, int k, float* out_boxes,
                       float* out_scores, int* out_labels, int cap, int slot0, cudaStream_t stream);
int topk_mask_launch(const float* logits, int frames, int N, int C, int k1, int k2, unsigned char* mask1,
                     unsigned char* mask2, cudaStream_t stream);
int gather_masked_rows_launch(const float* src, const unsigned char* mask, int frames, int N, int k, float* dst,
                              cudaStream_t stream);
int nms_launch(const float* boxes, const float* scores, const int* labels, const int* counts, int n, int cap,
               int frames, float thr, int plus_one, int ge, int ascending_out, float clip_w, float clip_h,
               long long* keep_idx, float* out_boxes, float* out_scores, int* out_labels, int* out_count,
               void* workspace, size_t workspace_bytes, cudaStream_t stream);
int cdist_launch(const float* x, float* out, int n, int d, cudaStream_t stream);
int fps_launch(int b, int n, int m, const float* dist, float* temp, int* idx, cudaStream_t stream);

int roi_align_legacy_launch(const float* in, const float* rois, int num_rois, int C, int H, int W, float scale, int PH,
                            int PW, int sampling_ratio, float* out, cudaStream_t stream);

int vid_match_launch(const float* pred_boxes, const int* pred_labels, const int* order, const int* pred_off,
                     const float* gt_boxes, const int* gt_labels, const unsigned char* gt_ignore, const int* gt_off,
                     int n_images, float thr, double empty_weight, unsigned char* gt_taken, unsigned char* hit,
                     double* weight, cudaStream_t stream);

int head_tail_launch(const void* fc, const void* cls_w, const float* cls_g, const float* cls_b, const void* logit_w,
                     const float* logit_bias, int C, const void* const* reg_w, const float* const* reg_g,
                     const float* const* reg_b, const void* delta_w, const float* delta_bias, const float* boxes_in,
                     float* logits_out, float* boxes_out, int M, cudaStream_t stream);

int swin_rows_launch(float* X, int write_x, const void* add, int add_mode, const float* gamma, const float* beta,
                     void* out16, float* out32, int out_mode, int B, int H, int W, int C, int shift,
                     cudaStream_t stream);
int swin_merge_launch(const float* X, int B, int H, int W, int C, const float* gamma, const float* beta, void* out,
                      cudaStream_t stream);
int swin_patch_gather_launch(const void* img, int is_u8, void* out, int B, int H, int W, const float* mean,
                             const float* std, cudaStream_t stream);
int swin_window_attention_launch(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads,
                                 int shift, cudaStream_t stream);

int jpeg_info(const unsigned char* data, long nbytes, int* width, int* height);
int jpeg_decode_rgb(const unsigned char* data, long nbytes, unsigned char* dst_hwc, int width, int height,
                    cudaStream_t stream);

int gemm256_row_launch(const void* a, const void* w, const float* bias, const float* resid, const float* gamma,
                       const float* beta, int act, float* out_f32, void* out_f16, int M, cudaStream_t stream);

int swin_window_attention_tc_launch(const void* qkv, const float* bias, void* out, int B, int H, int W, int C, int heads,
                                    int shift, cudaStream_t stream);

// ---------------------------------------------------------------------------------------------------------------
// Programmatic dependent launch: every kernel of the library is launched with the programmatic-stream-serialization
// attribute and starts with pdl_prologue(): `griddepcontrol.launch_dependents` lets the NEXT kernel of the stream be
// scheduled (and run its own prologue: barrier init, TMEM allocation, descriptor prefetch) while this one drains,
// `griddepcontrol.wait` blocks until the PREVIOUS kernel has completed and its writes are visible.  No global memory
// may be touched before the wait.  The hot path is ~5000 short dependent kernels per clip, so the launch-to-launch
// bubble matters.  DVID_PDL=0 launches without the attribute (the two instructions are then no-ops).
bool pdl_enabled();

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_prologue() {
  pdl_trigger();
  pdl_wait();
}

template <typename... KArgs, typename... Args>
inline void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                       Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

inline int check_launch() { return cudaGetLastError() == cudaSuccess ? DVID_OK : DVID_ERR_CUDA; }

}  // namespace dvid

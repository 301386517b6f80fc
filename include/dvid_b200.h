/*
 * dvid_b200 - C ABI of the B200-native DiffusionVID inference hot path (libdvid_b200.so).
 *
 * This library replaces, for the DiffusionDet/DiffusionVID model of sdroh1027/DiffusionVID, (a) the cuDNN/cuBLAS/ATen/
 * torchvision kernels the reference reaches through PyTorch and detectron2 and (b) its native extension mega_core/_C
 * (mega_core/csrc/vision.cpp:10-27).  Every entry point takes plain device pointers, sizes and a CUDA stream
 * (cudaStream_t passed as void*); no torch types cross this boundary.  All functions are asynchronous on `stream`,
 * never synchronise, never allocate device memory, and return 0 (DVID_OK) or a DVID_ERR_* code - they never exit the
 * process (the reference's FPS launcher calls exit(-1) on a launch error, mega_core/csrc/cuda/fps.cu:181-185).
 *
 * Layout conventions: activations NHWC fp16 ("half"), weights [Cout][R*S*Cin] fp16, box coordinates / logits / noise /
 * object features fp32, indices int32 unless stated.  Citations are into /root/reference (DiffusionVID @ 8375542);
 * "SURVEY A<n>" refers to the restated third-party semantics in SURVEY.md Appendix A.
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
 * Dense contractions (tcgen05 + TMA + TMEM), csrc/conv_gemm.cu.
 *
 * dvid_conv2d_nhwc_f16 replaces the cuDNN convolutions behind detectron2's ResNet/FPN that the reference builds at
 * mega_core/modeling/detector/diffusion_det.py:219 and calls at :427 (FrozenBN folded into weight/bias by the host).
 *   out[n,y,x,co] = act( bias[co] + sum_{r,s,ci} in[n, y*stride+r-pad, x*stride+s-pad, ci] * weight[co,(r*S+s)*Cin+ci]
 *                        + resid[n, y>>resid_shift, x>>resid_shift, co] )
 * resid_shift=1 implements FPN's nearest x2 top-down addition. bias/resid may be NULL.
 * relu: activation 0 none / 1 ReLU / 2 GELU(erf) (same meaning in dvid_gemm_f16).
 * Requirements: Cin % 8 == 0, Cout % 8 == 0, stride in {1,2}.
 */
DVID_API int dvid_conv2d_nhwc_f16(const void* in, const void* weight, const float* bias, const void* resid, void* out,
                         int n, int h, int w, int cin, int cout, int R, int S, int stride, int pad,
                         int resid_shift, int relu, void* stream);

/* Stem: 7x7 / stride 2 / pad 3 convolution of the 3-channel image (detectron2 BasicStem) + ReLU.
 * `in_haloed`: output of dvid_preprocess with halo 3: [n][H+6][W+6][8] fp16.  `weight`: [cout][7][8][8] fp16 with
 * weight[co][r][s][c] = w[co][c][r][s] * bn_scale for s<7, c<3 and zero elsewhere.  out: [n][H/2][W/2][cout]. */
DVID_API int dvid_stem_conv_f16(const void* in_haloed, const void* weight, const float* bias, void* out, int n, int H, int W,
                       int cout, int relu, void* stream);

/* dvid_gemm_f16 replaces torch.nn.functional.linear (cuBLAS) for every nn.Linear of the decoder
 * (mega_core/modeling/roi_heads/box_head/box_head.py:218-223,447-491,675-684):
 *   out_f16[m,n] = act( bias[n] + sum_k a[m,k] * w[n,k] + resid[m,n] )           (out_f32_partials == NULL)
 *   out_f32_partials[s,m,n] = sum_{k in split s} a[m,k] * w[n,k]                 (split-K; reduce with dvid_row_post)
 * K % 8 == 0, N % 8 == 0. `splits` is a request; the number actually used is returned in *splits_used (may be NULL).
 */
/* Stream-K scheduling for dvid_conv2d_nhwc_f16 / dvid_gemm_f16 layers whose tile count leaves >= 1/4 wave idle (e.g. the
 * 152-tile res4 convolutions on 148 SMs).  enable != 0 allocates the workspace (one fp32 accumulator tile + flag per
 * SM, 19.4 MB) on first use - call it outside stream capture - and switches the mode on; 0 switches it off.  One
 * workspace serves all launches, so keep it off when convolutions run concurrently on several streams.  Results are
 * the same fp32 sums in a different association (k-blocks of a tile summed in two parts). Off by default. */
DVID_API int dvid_conv_streamk(int enable);

DVID_API int dvid_gemm_f16(const void* a, const void* w, const float* bias, const void* resid, void* out_f16,
                  float* out_f32_partials, int m, int n, int k, int relu, int splits, int* splits_used,
                  void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Image-side memory-bound kernels, csrc/rowops.cu.
 */
/* normalizer (diffusion_det.py:301-303, applied :422) + NCHW fp32 -> NHWC fp16 (8 channels, 3 real) with a zero halo.
 * img [n][3][H][W] in [0,1]; out [n][Hp][Wp][8]; mean/std: 3 host floats (already divided by 255). */
DVID_API int dvid_preprocess(const float* img, void* out, int n, int H, int W, int halo, int Hp, int Wp, const float* mean,
                    const float* std, void* stream);
/* Clip-loader variant: img [n][3][H][W] uint8 as decoded; fuses the reference's ToTensor
 * (mega_core/data/transforms/transforms.py:295-297 = torchvision to_tensor: u8 -> fp32 / 255, applied by
 * transforms/build.py:75-83) in front of the normalizer.  Output bit-identical to dvid_preprocess(to_tensor(img)). */
DVID_API int dvid_preprocess_u8(const unsigned char* img, void* out, int n, int H, int W, int halo, int Hp, int Wp,
                       const float* mean, const float* std, void* stream);
/* max_pool2d(kernel 3, stride 2, pad 1) of the stem, NHWC fp16, C % 8 == 0. */
DVID_API int dvid_maxpool3x3s2_nhwc_f16(const void* in, void* out, int n, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Decoder (DynamicHead) kernels.
 */
/* softmax(Q K^T / sqrt(32)) V for head dim 32 - the core of torch.nn.MultiheadAttention at box_head.py:516,626 (per
 * frame self-attention) and :371 (global cross-attention).  fp16 in/out with element strides: row strides *_rs,
 * batch strides *_bs; head h reads columns [32h, 32h+32) of each row. */
DVID_API int dvid_attention_hd32(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                        long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                        void* stream);

/* The same contract with both contractions on tcgen05 (S and O accumulators in TMEM, softmax out of TMEM, 128 queries
 * per CTA, keys in chunks of 128; csrc/attention_tc.cu).  All strides must be multiples of 8 elements. */
DVID_API int dvid_attention_hd32_tc(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq,
                           int lk, long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs,
                           long o_bs, void* stream);

/* detectron2 ROIPooler -> torchvision roi_align(aligned=True, 7x7, sampling_ratio 2) over 3 FPN levels (call sites
 * box_head.py:507,617).  feats: 3 host pointers to device NHWC fp16 maps [frames][h_l][w_l][256];
 * boxes [num_boxes][4] xyxy fp32, box b belongs to frame b / boxes_per_frame.  roi_out [num_boxes][49][256] fp16
 * (position-major, box_head.py:512) may be NULL; mean_* [num_boxes][256] = mean over the 49 positions
 * (pro_features seed, box_head.py:509-510) may be NULL. */
DVID_API int dvid_roi_align(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                   int num_boxes, int boxes_per_frame, void* roi_out, float* mean_f32, void* mean_f16, void* stream);

/* Fused [ROIAlign +] DynamicConv bmm chain (box_head.py:698-704): out[b] = relu(LN256(relu(LN64(roi_b @ P1_b)) @ P2_b)),
 * params [num_boxes][32768] fp16 = dynamic_layer output (P1 [256][64] then P2 [64][256], box_head.py:693-696).
 * roi_in == NULL: the ROI tile is gathered in-kernel from feats/boxes (fused ROIAlign); else read from roi_in.
 * out [num_boxes][49][256] fp16 (input of out_layer). */
DVID_API int dvid_roi_dynconv(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                     int num_boxes, int boxes_per_frame, const void* roi_in, const void* params, const float* ln1_g,
                     const float* ln1_b, const float* ln2_g, const float* ln2_b, void* out, void* stream);

/* The same operation with both per-box contractions on tcgen05 (TMEM accumulators, LayerNorm out of TMEM, generated
 * weights by TMA): params_t [num_boxes][32768] fp16 holds the dynamic_layer output TRANSPOSED per box - P1^T [64][256]
 * (params_t[j*256+i] = P1[i][j]) then P2^T [256][64] (params_t[16384+i*64+j] = P2[j][i]) - i.e. the rows of the
 * dynamic_layer weight / bias are permuted once when the weights are packed (box_head.py:693-696 fixes the layout of
 * the reference's Linear; a row permutation of a Linear permutes its outputs and nothing else). */
DVID_API int dvid_roi_dynconv_tc(const void* const* feats, const int* hs, const int* ws, const float* scales,
                        const float* boxes, int num_boxes, int boxes_per_frame, const void* roi_in,
                        const void* params_t, const float* ln1_g, const float* ln1_b, const float* ln2_g,
                        const float* ln2_b, void* out, void* stream);

/* 256 -> 256 Linear fused with its row epilogue (csrc/gemm_row.cu): y = a @ w^T + bias [+ resid] -> [LayerNorm(ln_g, ln_b)]
 * -> out_f32 = y, out_f16 = act(y) (act: 0 none, 1 ReLU, 2 SiLU).  a [m][256] fp16, w [256][256] fp16, resid / out_f32
 * fp32 [m][256].  Replaces nn.Linear + residual + nn.LayerNorm at box_head.py:516-518 / :626-628 (self_attn.out_proj,
 * norm1) and the two Linears around the SiLU at :371 / :644 (global attention out_proj, c_mlp). */
DVID_API int dvid_gemm256_row(const void* a, const void* w, const float* bias, const float* resid, const float* ln_g,
                     const float* ln_b, int act, float* out_f32, void* out_f16, int m, void* stream);

/* Row kernel for 256-wide rows: y = sum_s partials[s] (or in_f16) + bias -> [LN1] -> [ReLU] -> [+resid] -> [LN2] ->
 * act2 (0 none / 1 ReLU / 2 SiLU; on the fp16 output only if act2_f16_only) -> out_f32 / out_f16; optional time /
 * condition modulation out_mod_f16 = y * (scale[row / rows_per_group] + 1) + shift (box_head.py:533-536, :643-647).
 * Implements every LayerNorm / residual / activation between the decoder GEMMs (box_head.py:517-529,540-543,707-709). */
DVID_API int dvid_row_post(const float* partials, int splits, long split_stride, const void* in_f16, const float* bias,
                  const float* ln1_g, const float* ln1_b, int relu1, const float* resid, const float* ln2_g,
                  const float* ln2_b, int act2, int act2_f16_only, float* out_f32, void* out_f16,
                  const float* mod_scale, const float* mod_shift, int rows_per_group, int scale_stride,
                  int shift_stride, int shift_per_row, void* out_mod_f16, int M, void* stream);

/* out[m][n] = act_out(bias[n] + sum_k act_in(a[m][k]) * w[n][k]) for m <= 8 rows: time_mlp (box_head.py:218-223) and
 * block_time_mlp (:464,:602).  a fp32, w fp16 [n][k], act_in 1 = SiLU, act_out 1 = GELU(erf). */
DVID_API int dvid_small_linear(const float* a, const void* w, const float* bias, float* out, int m, int n, int k, int act_in,
                      int act_out, void* stream);
/* SinusoidalPositionEmbeddings (box_head.py:729-741), dim 256: out[m][256] = [sin(t*freq), cos(t*freq)]. */
DVID_API int dvid_time_sinusoid(const float* t, const float* freq, float* out, int m, void* stream);

/* class_logits / bboxes_delta bias + RCNNHead.apply_deltas (box_head.py:544-590). */
DVID_API int dvid_head_final(const float* logit_part, int ldl, const float* cls_bias, int C, const float* delta_part, int ldd,
                    const float* delta_bias, const float* boxes_in, float* logits_out, float* boxes_out, int M,
                    void* stream);

/* Fused tail of RCNNHead.forward / RCNNHead_cond.forward (box_head.py:538-548, :649-664) for NUM_CLS = 1, NUM_REG = 3:
 * cls tower (Linear no-bias + LayerNorm + ReLU) -> class_logits (+bias); reg tower x3 -> bboxes_delta (+bias) ->
 * apply_deltas (:550-590).  One CTA per 128 rows keeps the activations in shared memory across the six tcgen05 GEMMs.
 * fc [M][256] fp16 (time/cond-modulated object features); cls_w / reg_w* [256][256] fp16; logit_w [32][256] fp16 (rows
 * >= C zero); delta_w [16][256] fp16 (rows >= 4 zero); LayerNorm params and biases fp32; boxes_in [M][4] xyxy;
 * logits_out [M][C], boxes_out [M][4] fp32. */
DVID_API int dvid_head_tail(const void* fc, const void* cls_w, const float* cls_ln_g, const float* cls_ln_b,
                   const void* logit_w, const float* logit_bias, int C, const void* reg_w0, const void* reg_w1,
                   const void* reg_w2, const float* reg_ln_g0, const float* reg_ln_b0, const float* reg_ln_g1,
                   const float* reg_ln_b1, const float* reg_ln_g2, const float* reg_ln_b2, const void* delta_w,
                   const float* delta_bias, const float* boxes_in, float* logits_out, float* boxes_out, int M,
                   void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Diffusion loop, csrc/rowops.cu + csrc/postproc.cu.
 */
/* DiffusionDet.model_predictions input map (diffusion_det.py:657-660): noise-space cxcywh -> absolute xyxy. */
DVID_API int dvid_noise_to_boxes(const float* x, float* boxes, int M, float scale, float W, float H, void* stream);
/* One DDIM step with box renewal, one CTA per frame (diffusion_det.py:559-596 and :668-676).  Scalars are the fp32
 * schedule constants the host derives in float64 exactly as the reference does (:578-584). */
DVID_API int dvid_ddim_step(const float* logits, int C, const float* coord, const float* x_t, const float* eps,
                   const float* fill, float* x_next, float* boxes_next, int* num_kept, int frames, int N, float scale,
                   float W, float H, float sqrt_recip_a, float sqrt_recipm1_a, float sqrt_a_next, float c_coef,
                   float sigma, void* stream);
/* DiffusionDet.inference per-frame top-k of the N*C sigmoid scores (diffusion_det.py:772-784); results are written at
 * slot0 of per-frame candidate buffers of capacity `cap` (the ensemble concatenation of :608-610). */
DVID_API int dvid_topk_scores(const float* logits, const float* boxes, int frames, int N, int C, int k, float* out_boxes,
                     float* out_scores, int* out_labels, int cap, int slot0, void* stream);
/* Per-frame top-k1 / top-k2 masks of the max logit (box_head.py:304-311) and the masked row gather (:315-317). */
DVID_API int dvid_topk_mask(const float* logits, int frames, int N, int C, int k1, int k2, unsigned char* mask1,
                   unsigned char* mask2, void* stream);
DVID_API int dvid_gather_masked_rows(const float* src, const unsigned char* mask, int frames, int N, int k, float* dst,
                            void* stream);
/* Greedy NMS, one CTA per frame, n <= 1024 candidates, sweep on the device.  labels != NULL: torchvision batched_nms
 * coordinate trick (diffusion_det.py:617,793).  plus_one/ge/ascending_out select the legacy mega_core._C.nms
 * semantics (mega_core/csrc/cpu/nms_cpu.cpp:5-65, cuda/nms.cu:13-67).  clip_w>0 applies BoxList.clip_to_image
 * (mega_core/structures/bounding_box.py:214-224) to out_boxes.  keep_idx is int64 like the reference's return.
 * workspace (optional, device, >= frames * DVID_NMS_WORKSPACE_PER_FRAME bytes): with it the O(n^2) suppression matrix
 * is computed by a separate kernel spread over the whole GPU (sort -> matrix -> sweep, three launches); without it
 * one CTA per frame does everything.  Results are identical. */
#define DVID_NMS_WORKSPACE_PER_FRAME (1024 * 8 + 1024 * 16 + 1024 * 16 * 8)
DVID_API int dvid_nms(const float* boxes, const float* scores, const int* labels, const int* counts, int n, int cap,
             int frames, float thr, int plus_one, int ge, int ascending_out, float clip_w, float clip_h,
             long long* keep_idx, float* out_boxes, float* out_scores, int* out_labels, int* out_count,
             void* workspace, long workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Global memory management (diffusion_det.py:841-896).
 */
/* torch.cdist(x, x, p=2) by direct differences, fp32 (diffusion_det.py:880). */
DVID_API int dvid_cdist_f32(const float* x, float* out, int n, int d, void* stream);
/* Drop-in for mega_core._C.furthest_point_sampling (mega_core/csrc/fps.h:15-36, cuda/fps.cu:25-185): dist (b,n,n) fp32,
 * temp (b,n) pre-filled with 1e10, idx (b,m) int32; identical picks including the kernel's tie-breaking. */
DVID_API int dvid_furthest_point_sampling(int b, int n, int m, const float* dist, float* temp, int* idx, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Legacy operators of mega_core._C outside the DiffusionVID path, csrc/legacy_ops.cu.
 */
/* Drop-in for mega_core._C.roi_align_forward (mega_core/csrc/ROIAlign.h:11-27, cuda/ROIAlign_cuda.cu:65-125,256-300):
 * the maskrcnn-benchmark ROIAlign - NO half-pixel shift, roi size clamped to >= 1, sampling_ratio <= 0 = adaptive grid.
 * input [N][C][H][W] fp32, rois [num_rois][5] fp32 = (batch index, x1, y1, x2, y2), out [num_rois][C][ph][pw] fp32. */
DVID_API int dvid_roi_align_legacy_forward(const float* input, const float* rois, int num_rois, int channels,
                                           int height, int width, float spatial_scale, int pooled_height,
                                           int pooled_width, int sampling_ratio, float* out, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * ImageNet-VID evaluator, csrc/vid_match.cu.
 */
/* The greedy matching of mega_core/data/datasets/evaluation/vid/vid_eval.py:167-291 (calc_detection_vid_prec_rec,
 * incl. the motion-specific "ignored ground truth" rules :233-264) for ALL images in one launch.  Packed inputs:
 * pred_boxes [Np][4] fp32 xyxy, pred_labels [Np], pred_off [n_images+1] (image i owns detections pred_off[i] ..
 * pred_off[i+1]-1), order [Np] = detection indices, image by image, each image in descending score order (stable);
 * gt_boxes [Ng][4], gt_labels [Ng], gt_ignore [Ng] (1 = outside the motion range), gt_off [n_images+1];
 * gt_taken [Ng] zero-initialised scratch.  Outputs, indexed by POSITION in `order`: hit [Np] (1 = matched) and
 * weight [Np] fp64 = the reference's pred_ignore (0 regular, 1 dropped, fraction = partial false positive;
 * empty_weight for a detection whose class has no ground truth in the image). */
DVID_API int dvid_vid_match(const float* pred_boxes, const int* pred_labels, const int* order, const int* pred_off,
                            const float* gt_boxes, const int* gt_labels, const unsigned char* gt_ignore,
                            const int* gt_off, int n_images, float iou_thresh, double empty_weight,
                            unsigned char* gt_taken, unsigned char* hit, double* weight, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * Swin Transformer backbone (mega_core/modeling/backbone/swintransformer.py), csrc/swin.cu.  The linear layers
 * (qkv / proj / fc1+GELU / fc2 / reduction / patch-embed projection) are dvid_gemm_f16 calls.
 */
/* Row kernel over the fp32 residual stream x [B][H][W][C] (C % 128 == 0, C <= 2048):
 *   v = x[token] (0 if x == NULL) + add[...]      add fp16 rows; add_mode 0 none / 1 token order / 2 shifted-window order
 *                                                 (window_reverse + roll(+shift) + crop, swintransformer.py:259-270)
 *   x[token] = v if write_x;  y = LayerNorm(v; gamma, beta, eps 1e-5) (y = v if gamma == NULL)
 *   out_f32[token] = y;  out_f16[row] = y with out_mode 1 token order / 2 shifted-window order including zero rows
 *   for the padding to a multiple of 7 (F.pad + roll(-shift) + window_partition, :236-252).  shift: 0 or 3. */
DVID_API int dvid_swin_rows(float* x, int write_x, const void* add, int add_mode, const float* gamma, const float* beta,
                   void* out_f16, float* out_f32, int out_mode, int B, int H, int W, int C, int shift, void* stream);
/* PatchMerging (:279-317): out_f16 [B*ceil(H/2)*ceil(W/2)][4C] = LayerNorm_4C(cat of the 2x2 neighbours, zero padded). */
DVID_API int dvid_swin_patch_merge(const float* x, int B, int H, int W, int C, const float* gamma, const float* beta,
                          void* out_f16, void* stream);
/* normalizer (diffusion_det.py:301-303) + 4x4 patch extraction for PatchEmbed (:422-461): img [B][3][H][W] fp32 ->
 * out_f16 [B*(H/4)*(W/4)][64], k = c*16 + py*4 + px, k >= 48 zero.  mean/std: 3 host floats (already / 255). */
DVID_API int dvid_swin_patch_gather(const float* img, void* out_f16, int B, int H, int W, const float* mean,
                           const float* std, void* stream);
/* Clip loader, first stage: the reference's test-time Resize (mega_core/data/transforms/transforms.py:31-67 ->
 * torchvision F.resize on a PIL image = Pillow Image.resize(BILINEAR): antialiased two-pass triangle filter in 22-bit
 * fixed point, libImaging/Resample.c) on decoded frames.  src_hwc [n][Hin][Win][3] uint8 -> dst_chw [n][3][Hp][Wp]
 * uint8 with the oh x ow result in the top-left corner and zeros elsewhere (the padding to_image_list adds,
 * structures/image_list.py:36-66).  Output bytes equal Pillow's.  `workspace`: device scratch of at least
 * dvid_resize_workspace_bytes(...) bytes (horizontal-pass image + coefficient tables, computed on the device). */
DVID_API long dvid_resize_workspace_bytes(int n, int Hin, int Win, int oh, int ow);
DVID_API int dvid_resize_bilinear_u8(const unsigned char* src_hwc, int n, int Hin, int Win, int oh, int ow,
                            unsigned char* dst_chw, int Hp, int Wp, void* workspace, long workspace_bytes,
                            void* stream);
/* uint8 frames (ToTensor fused, see dvid_preprocess_u8); W % 4 == 0 keeps the 4-pixel loads aligned. */
DVID_API int dvid_swin_patch_gather_u8(const unsigned char* img, void* out_f16, int B, int H, int W, const float* mean,
                              const float* std, void* stream);
/* WindowAttention core (:145-176): softmax(q*scale k^T + bias + shift mask) v per (window, head), head dim 32.
 * qkv [windows*49][3C] fp16 in shifted-window order, bias [heads][49][49] fp32 (table gathered by
 * relative_position_index), out_f16 [windows*49][C].  The SW-MSA mask of BasicLayer.forward (:387-406) is derived from
 * the token coordinates.  H, W: token grid of the stage (unpadded). */
DVID_API int dvid_swin_window_attention(const void* qkv, const float* bias, void* out_f16, int B, int H, int W, int C,
                               int heads, int shift, void* stream);
/* The same contract on tcgen05: two windows (2 x 64 padded rows) per 128-row tile, scores and outputs in TMEM, bias +
 * shift mask + softmax out of TMEM (csrc/swin_attention_tc.cu). */
DVID_API int dvid_swin_window_attention_tc(const void* qkv, const float* bias, void* out_f16, int B, int H, int W, int C,
                                  int heads, int shift, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * JPEG decode in front of the clip loader, csrc/jpeg_decode.cu: nvJPEG (CUDA toolkit library, opened with
 * dlopen on first use; DVID_ERR_DRIVER when it is not installed).  Replaces PIL's Image.open(...).convert("RGB") of the
 * reference's datasets (mega_core/data/datasets/vid.py) for baseline / progressive JPEG files.
 */
/* width / height of the JPEG in `data` (HOST memory, nbytes long). */
DVID_API int dvid_jpeg_info(const unsigned char* data, long nbytes, int* width, int* height);
/* Decode into dst_hwc (DEVICE memory, [height][width][3] uint8, RGB interleaved) on `stream`; `data` is host memory. */
DVID_API int dvid_jpeg_decode_rgb(const unsigned char* data, long nbytes, unsigned char* dst_hwc, int width, int height,
                         void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DVID_B200_H */

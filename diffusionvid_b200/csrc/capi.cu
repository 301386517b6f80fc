// extern "C" surface of libdvid_b200.so: argument validation + dispatch to the kernel launchers. See include/dvid_b200.h.
#include "../../include/dvid_b200.h"
#include "dvid_internal.h"

#define DVID_ABI_VERSION 16

static inline cudaStream_t S(void* s) { return static_cast<cudaStream_t>(s); }

extern "C" {

int dvid_abi_version(void) { return DVID_ABI_VERSION; }
int dvid_num_sms(void) { return dvid::num_sms(); }

int dvid_conv2d_nhwc_f16(const void* in, const void* weight, const float* bias, const void* resid, void* out, int n,
                         int h, int w, int cin, int cout, int R, int S_, int stride, int pad, int resid_shift,
                         int relu, void* stream) {
  if (!in || !weight || !out) return DVID_ERR_ARG;
  return dvid::conv_gemm_launch(in, weight, bias, resid, out, nullptr, n, h, w, cin, cout, R, S_, stride, pad,
                                resid_shift, relu, 1, 0, S(stream), nullptr);
}

int dvid_stem_conv_f16(const void* in_haloed, const void* weight, const float* bias, void* out, int n, int H, int W,
                       int cout, int relu, void* stream) {
  if (!in_haloed || !weight || !out) return DVID_ERR_ARG;
  return dvid::stem_conv_launch(in_haloed, weight, bias, out, n, H, W, cout, relu, S(stream));
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
                                n, 1, 1, 1, 0, 0, relu, splits, 0, S(stream));
}

int dvid_roi_align_legacy_forward(const float* input, const float* rois, int num_rois, int channels, int height,
                                  int width, float spatial_scale, int pooled_height, int pooled_width,
                                  int sampling_ratio, float* out, void* stream) {
  if (num_rois > 0 && (!input || !rois || !out)) return DVID_ERR_ARG;
  return dvid::roi_align_legacy_launch(input, rois, num_rois, channels, height, width, spatial_scale, pooled_height,
                                       pooled_width, sampling_ratio, out, S(stream));
}

int dvid_vid_match(const float* pred_boxes, const int* pred_labels, const int* order, const int* pred_off,
                   const float* gt_boxes, const int* gt_labels, const unsigned char* gt_ignore, const int* gt_off,
                   int n_images, float iou_thresh, double empty_weight, unsigned char* gt_taken, unsigned char* hit,
                   double* weight, void* stream) {
  if (n_images > 0 && (!pred_off || !gt_off || !hit || !weight)) return DVID_ERR_ARG;   // the packed arrays may be empty
  return dvid::vid_match_launch(pred_boxes, pred_labels, order, pred_off, gt_boxes, gt_labels, gt_ignore, gt_off,
                                n_images, iou_thresh, empty_weight, gt_taken, hit, weight, S(stream));
}

int dvid_conv_streamk(int enable) { return dvid::conv_streamk_enable(enable); }

int dvid_preprocess(const float* img, void* out, int n, int H, int W, int halo, int Hp, int Wp, const float* mean,
                    const float* std, void* stream) {
  if (!img || !out || !mean || !std) return DVID_ERR_ARG;
  return dvid::preprocess_launch(img, 0, out, n, H, W, halo, Hp, Wp, mean, std, S(stream));
}

long dvid_resize_workspace_bytes(int n, int Hin, int Win, int oh, int ow) {
  return dvid::resize_workspace_bytes(n, Hin, Win, oh, ow);
}

int dvid_resize_bilinear_u8(const unsigned char* src_hwc, int n, int Hin, int Win, int oh, int ow,
                            unsigned char* dst_chw, int Hp, int Wp, void* workspace, long workspace_bytes,
                            void* stream) {
  if (!src_hwc || !dst_chw || !workspace) return DVID_ERR_ARG;
  return dvid::resize_bilinear_u8_launch(src_hwc, n, Hin, Win, oh, ow, dst_chw, Hp, Wp, workspace, workspace_bytes,
                                         S(stream));
}

int dvid_preprocess_u8(const unsigned char* img, void* out, int n, int H, int W, int halo, int Hp, int Wp,
                       const float* mean, const float* std, void* stream) {
  if (!img || !out || !mean || !std) return DVID_ERR_ARG;
  return dvid::preprocess_launch(img, 1, out, n, H, W, halo, Hp, Wp, mean, std, S(stream));
}

int dvid_maxpool3x3s2_nhwc_f16(const void* in, void* out, int n, int H, int W, int C, void* stream) {
  if (!in || !out) return DVID_ERR_ARG;
  return dvid::maxpool_launch(in, out, n, H, W, C, S(stream));
}

int dvid_attention_hd32(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                        long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                        void* stream) {
  if (!q || !k || !v || !o) return DVID_ERR_ARG;
  return dvid::attention_launch(q, k, v, o, batch, heads, lq, lk, q_rs, k_rs, v_rs, o_rs, q_bs, k_bs, v_bs, o_bs,
                                S(stream));
}

int dvid_attention_hd32_tc(const void* q, const void* k, const void* v, void* o, int batch, int heads, int lq, int lk,
                           long q_rs, long k_rs, long v_rs, long o_rs, long q_bs, long k_bs, long v_bs, long o_bs,
                           void* stream) {
  if (!q || !k || !v || !o) return DVID_ERR_ARG;
  return dvid::attention_tc_launch(q, k, v, o, batch, heads, lq, lk, q_rs, k_rs, v_rs, o_rs, q_bs, k_bs, v_bs, o_bs,
                                   S(stream));
}

int dvid_roi_align(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                   int num_boxes, int boxes_per_frame, void* roi_out, float* mean_f32, void* mean_f16, void* stream) {
  if (!feats || !hs || !ws || !scales || !boxes) return DVID_ERR_ARG;
  return dvid::roi_align_launch(feats, hs, ws, scales, boxes, num_boxes, boxes_per_frame, roi_out, mean_f32, mean_f16,
                                S(stream));
}

int dvid_roi_dynconv(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                     int num_boxes, int boxes_per_frame, const void* roi_in, const void* params, const float* ln1_g,
                     const float* ln1_b, const float* ln2_g, const float* ln2_b, void* out, void* stream) {
  if (!params || !ln1_g || !ln1_b || !ln2_g || !ln2_b || !out) return DVID_ERR_ARG;
  if (!roi_in && (!feats || !hs || !ws || !scales || !boxes)) return DVID_ERR_ARG;
  static const void* const null_feats[3] = {nullptr, nullptr, nullptr};
  static const int zeros[3] = {0, 0, 0};
  static const float zf[3] = {0.f, 0.f, 0.f};
  if (roi_in) {
    if (!feats) feats = null_feats;
    if (!hs) hs = zeros;
    if (!ws) ws = zeros;
    if (!scales) scales = zf;
  }
  return dvid::roi_dynconv_launch(feats, hs, ws, scales, boxes, num_boxes, boxes_per_frame, roi_in, params, ln1_g,
                                  ln1_b, ln2_g, ln2_b, out, S(stream));
}

int dvid_roi_dynconv_tc(const void* const* feats, const int* hs, const int* ws, const float* scales, const float* boxes,
                        int num_boxes, int boxes_per_frame, const void* roi_in, const void* params_t,
                        const float* ln1_g, const float* ln1_b, const float* ln2_g, const float* ln2_b, void* out,
                        void* stream) {
  if (!params_t || !ln1_g || !ln1_b || !ln2_g || !ln2_b || !out) return DVID_ERR_ARG;
  if (!roi_in && (!feats || !hs || !ws || !scales || !boxes)) return DVID_ERR_ARG;
  static const void* const null_feats[3] = {nullptr, nullptr, nullptr};
  static const int zeros[3] = {0, 0, 0};
  static const float zf[3] = {0.f, 0.f, 0.f};
  if (roi_in) {
    if (!feats) feats = null_feats;
    if (!hs) hs = zeros;
    if (!ws) ws = zeros;
    if (!scales) scales = zf;
  }
  return dvid::roi_dynconv_tc_launch(feats, hs, ws, scales, boxes, num_boxes, boxes_per_frame, roi_in, params_t, ln1_g,
                                     ln1_b, ln2_g, ln2_b, out, S(stream));
}

int dvid_row_post(const float* partials, int splits, long split_stride, const void* in_f16, const float* bias,
                  const float* ln1_g, const float* ln1_b, int relu1, const float* resid, const float* ln2_g,
                  const float* ln2_b, int act2, int act2_f16_only, float* out_f32, void* out_f16,
                  const float* mod_scale, const float* mod_shift, int rows_per_group, int scale_stride,
                  int shift_stride, int shift_per_row, void* out_mod_f16, int M, void* stream) {
  return dvid::row_post_launch(partials, splits, split_stride, in_f16, bias, ln1_g, ln1_b, relu1, resid, ln2_g, ln2_b,
                               act2, act2_f16_only, out_f32, out_f16, mod_scale, mod_shift, rows_per_group,
                               scale_stride, shift_stride, shift_per_row, out_mod_f16, M, S(stream));
}

int dvid_small_linear(const float* a, const void* w, const float* bias, float* out, int m, int n, int k, int act_in,
                      int act_out, void* stream) {
  if (!a || !w || !out) return DVID_ERR_ARG;
  return dvid::small_linear_launch(a, w, bias, out, m, n, k, act_in, act_out, S(stream));
}

int dvid_time_sinusoid(const float* t, const float* freq, float* out, int m, void* stream) {
  if (!t || !freq || !out) return DVID_ERR_ARG;
  return dvid::time_sinusoid_launch(t, freq, out, m, S(stream));
}

int dvid_head_final(const float* logit_part, int ldl, const float* cls_bias, int C, const float* delta_part, int ldd,
                    const float* delta_bias, const float* boxes_in, float* logits_out, float* boxes_out, int M,
                    void* stream) {
  if (!logit_part || !cls_bias || !delta_part || !delta_bias || !boxes_in || !logits_out || !boxes_out)
    return DVID_ERR_ARG;
  return dvid::head_final_launch(logit_part, ldl, cls_bias, C, delta_part, ldd, delta_bias, boxes_in, logits_out,
                                 boxes_out, M, S(stream));
}

int dvid_noise_to_boxes(const float* x, float* boxes, int M, float scale, float W, float H, void* stream) {
  if (!x || !boxes) return DVID_ERR_ARG;
  return dvid::noise_to_boxes_launch(x, boxes, M, scale, W, H, S(stream));
}

int dvid_ddim_step(const float* logits, int C, const float* coord, const float* x_t, const float* eps,
                   const float* fill, float* x_next, float* boxes_next, int* num_kept, int frames, int N, float scale,
                   float W, float H, float sqrt_recip_a, float sqrt_recipm1_a, float sqrt_a_next, float c_coef,
                   float sigma, void* stream) {
  if (!logits || !coord || !x_t || !eps || !fill || !x_next || !boxes_next) return DVID_ERR_ARG;
  return dvid::ddim_step_launch(logits, C, coord, x_t, eps, fill, x_next, boxes_next, num_kept, frames, N, scale, W, H,
                                sqrt_recip_a, sqrt_recipm1_a, sqrt_a_next, c_coef, sigma, S(stream));
}

int dvid_topk_scores(const float* logits, const float* boxes, int frames, int N, int C, int k, float* out_boxes,
                     float* out_scores, int* out_labels, int cap, int slot0, void* stream) {
  if (!logits || !boxes || !out_boxes || !out_scores || !out_labels) return DVID_ERR_ARG;
  return dvid::topk_scores_launch(logits, boxes, frames, N, C, k, out_boxes, out_scores, out_labels, cap, slot0,
                                  S(stream));
}

int dvid_topk_mask(const float* logits, int frames, int N, int C, int k1, int k2, unsigned char* mask1,
                   unsigned char* mask2, void* stream) {
  if (!logits || !mask1 || !mask2) return DVID_ERR_ARG;
  return dvid::topk_mask_launch(logits, frames, N, C, k1, k2, mask1, mask2, S(stream));
}

int dvid_gather_masked_rows(const float* src, const unsigned char* mask, int frames, int N, int k, float* dst,
                            void* stream) {
  if (!src || !mask || !dst) return DVID_ERR_ARG;
  return dvid::gather_masked_rows_launch(src, mask, frames, N, k, dst, S(stream));
}

int dvid_nms(const float* boxes, const float* scores, const int* labels, const int* counts, int n, int cap,
             int frames, float thr, int plus_one, int ge, int ascending_out, float clip_w, float clip_h,
             long long* keep_idx, float* out_boxes, float* out_scores, int* out_labels, int* out_count,
             void* workspace, long workspace_bytes, void* stream) {
  if (!boxes || !scores || !out_count) return DVID_ERR_ARG;
  return dvid::nms_launch(boxes, scores, labels, counts, n, cap, frames, thr, plus_one, ge, ascending_out, clip_w,
                          clip_h, keep_idx, out_boxes, out_scores, out_labels, out_count, workspace,
                          workspace_bytes > 0 ? static_cast<size_t>(workspace_bytes) : 0, S(stream));
}

int dvid_cdist_f32(const float* x, float* out, int n, int d, void* stream) {
  if (!x || !out) return DVID_ERR_ARG;
  return dvid::cdist_launch(x, out, n, d, S(stream));
}

int dvid_furthest_point_sampling(int b, int n, int m, const float* dist, float* temp, int* idx, void* stream) {
  if (!dist || !temp || !idx) return DVID_ERR_ARG;
  return dvid::fps_launch(b, n, m, dist, temp, idx, S(stream));
}

int dvid_head_tail(const void* fc, const void* cls_w, const float* cls_ln_g, const float* cls_ln_b, const void* logit_w,
                   const float* logit_bias, int C, const void* reg_w0, const void* reg_w1, const void* reg_w2,
                   const float* reg_ln_g0, const float* reg_ln_b0, const float* reg_ln_g1, const float* reg_ln_b1,
                   const float* reg_ln_g2, const float* reg_ln_b2, const void* delta_w, const float* delta_bias,
                   const float* boxes_in, float* logits_out, float* boxes_out, int M, void* stream) {
  if (!fc || !cls_w || !cls_ln_g || !cls_ln_b || !logit_w || !logit_bias || !reg_w0 || !reg_w1 || !reg_w2 ||
      !reg_ln_g0 || !reg_ln_b0 || !reg_ln_g1 || !reg_ln_b1 || !reg_ln_g2 || !reg_ln_b2 || !delta_w || !delta_bias ||
      !boxes_in || !logits_out || !boxes_out)
    return DVID_ERR_ARG;
  const void* rw[3] = {reg_w0, reg_w1, reg_w2};
  const float* rg[3] = {reg_ln_g0, reg_ln_g1, reg_ln_g2};
  const float* rb[3] = {reg_ln_b0, reg_ln_b1, reg_ln_b2};
  return dvid::head_tail_launch(fc, cls_w, cls_ln_g, cls_ln_b, logit_w, logit_bias, C, rw, rg, rb, delta_w, delta_bias,
                                boxes_in, logits_out, boxes_out, M, S(stream));
}

int dvid_swin_rows(float* x, int write_x, const void* add, int add_mode, const float* gamma, const float* beta,
                   void* out_f16, float* out_f32, int out_mode, int B, int H, int W, int C, int shift, void* stream) {
  return dvid::swin_rows_launch(x, write_x, add, add_mode, gamma, beta, out_f16, out_f32, out_mode, B, H, W, C, shift,
                                S(stream));
}

int dvid_swin_patch_merge(const float* x, int B, int H, int W, int C, const float* gamma, const float* beta,
                          void* out_f16, void* stream) {
  if (!x || !gamma || !beta || !out_f16) return DVID_ERR_ARG;
  return dvid::swin_merge_launch(x, B, H, W, C, gamma, beta, out_f16, S(stream));
}

int dvid_swin_patch_gather(const float* img, void* out_f16, int B, int H, int W, const float* mean, const float* std,
                           void* stream) {
  if (!img || !out_f16 || !mean || !std) return DVID_ERR_ARG;
  return dvid::swin_patch_gather_launch(img, 0, out_f16, B, H, W, mean, std, S(stream));
}

int dvid_swin_patch_gather_u8(const unsigned char* img, void* out_f16, int B, int H, int W, const float* mean,
                              const float* std, void* stream) {
  if (!img || !out_f16 || !mean || !std) return DVID_ERR_ARG;
  return dvid::swin_patch_gather_launch(img, 1, out_f16, B, H, W, mean, std, S(stream));
}

int dvid_swin_window_attention(const void* qkv, const float* bias, void* out_f16, int B, int H, int W, int C,
                               int heads, int shift, void* stream) {
  if (!qkv || !bias || !out_f16) return DVID_ERR_ARG;
  return dvid::swin_window_attention_launch(qkv, bias, out_f16, B, H, W, C, heads, shift, S(stream));
}

int dvid_jpeg_info(const unsigned char* data, long nbytes, int* width, int* height) {
  if (!data || nbytes <= 0 || !width || !height) return DVID_ERR_ARG;
  return dvid::jpeg_info(data, nbytes, width, height);
}

int dvid_jpeg_decode_rgb(const unsigned char* data, long nbytes, unsigned char* dst_hwc, int width, int height,
                         void* stream) {
  if (!data || nbytes <= 0 || !dst_hwc || width <= 0 || height <= 0) return DVID_ERR_ARG;
  return dvid::jpeg_decode_rgb(data, nbytes, dst_hwc, width, height, S(stream));
}

int dvid_gemm256_row(const void* a, const void* w, const float* bias, const float* resid, const float* ln_g,
                     const float* ln_b, int act, float* out_f32, void* out_f16, int m, void* stream) {
  if (!a || !w || (!out_f32 && !out_f16) || (ln_g && !ln_b)) return DVID_ERR_ARG;
  return dvid::gemm256_row_launch(a, w, bias, resid, ln_g, ln_b, act, out_f32, out_f16, m, S(stream));
}

int dvid_swin_window_attention_tc(const void* qkv, const float* bias, void* out_f16, int B, int H, int W, int C,
                                  int heads, int shift, void* stream) {
  if (!qkv || !bias || !out_f16) return DVID_ERR_ARG;
  return dvid::swin_window_attention_tc_launch(qkv, bias, out_f16, B, H, W, C, heads, shift, S(stream));
}

}  // extern "C"

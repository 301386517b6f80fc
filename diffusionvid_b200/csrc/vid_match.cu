// Greedy detection <-> ground-truth matching of the ImageNet-VID evaluator on the GPU ("GPU
// IoU-matching for AP").
//
// Reference: mega_core/data/datasets/evaluation/vid/vid_eval.py:167-291 (calc_detection_vid_prec_rec): per image and
// class, detections in descending score order take the unassigned ground-truth box of highest IoU >= thresh; with
// motion-specific evaluation (:172-181,192-197,233-264) boxes outside the motion range are "ignored" and change how a
// detection is weighed.  The reference walks Python lists (classes inside images); the evaluator of this package
// (evaluation.py) flattens every image's boxes into packed tensors and this kernel matches ALL images in one launch.
//
// One warp per image.  Detections are visited sequentially in the given order (the assignment is order dependent);
// the lanes stride over the image's ground-truth boxes: IoU, class test, candidate test, then warp reductions for the
// best IoU and the tie rule.  The IoU follows evaluation._vid_iou operation by operation in round-to-nearest fp32
// (integer-box convention: +1 on x2/y2, then the legacy +1 widths of boxlist_ops.py:83-88), so thresholds and ties
// fall exactly where the CPU evaluator puts them.  Integer / fp32 SIMT work, a few MB for the whole VID val set.
#include "dvid_internal.h"

namespace dvid {

namespace {

__device__ __forceinline__ float vid_iou(const float4 a, const float4 b) {
  const float ax2 = __fadd_rn(a.z, 1.f), ay2 = __fadd_rn(a.w, 1.f);
  const float bx2 = __fadd_rn(b.z, 1.f), by2 = __fadd_rn(b.w, 1.f);
  const float area_a = __fmul_rn(__fadd_rn(__fsub_rn(ax2, a.x), 1.f), __fadd_rn(__fsub_rn(ay2, a.y), 1.f));
  const float area_b = __fmul_rn(__fadd_rn(__fsub_rn(bx2, b.x), 1.f), __fadd_rn(__fsub_rn(by2, b.y), 1.f));
  const float w = fmaxf(__fadd_rn(__fsub_rn(fminf(ax2, bx2), fmaxf(a.x, b.x)), 1.f), 0.f);
  const float h = fmaxf(__fadd_rn(__fsub_rn(fminf(ay2, by2), fmaxf(a.y, b.y)), 1.f), 0.f);
  const float inter = __fmul_rn(w, h);
  return __fdiv_rn(inter, __fsub_rn(__fadd_rn(area_a, area_b), inter));
}

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_min_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_max_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(128)
vid_match_kernel(const float4* __restrict__ pred_boxes, const int* __restrict__ pred_labels,
                 const int* __restrict__ order, const int* __restrict__ pred_off, const float4* __restrict__ gt_boxes,
                 const int* __restrict__ gt_labels, const unsigned char* __restrict__ gt_ignore,
                 const int* __restrict__ gt_off, int n_images, float thr, double empty_weight,
                 unsigned char* __restrict__ gt_taken, unsigned char* __restrict__ hit, double* __restrict__ weight) {
  pdl_prologue();
  const int lane = threadIdx.x & 31;
  const int img = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (img >= n_images) return;
  const int p0 = pred_off[img], p1 = pred_off[img + 1];
  const int g0 = gt_off[img], g1 = gt_off[img + 1];
  for (int j = p0; j < p1; ++j) {            // position j of the output = j-th detection of the image by score
    const int p = order[j];
    const float4 pb = __ldg(pred_boxes + p);
    const int pl = __ldg(pred_labels + p);
    // pass 1: class census, best candidate IoU, best overlaps with ignored / regular boxes of the class
    int n_own = 0, n_ign = 0;
    float best = -2.f, ig = -1.f, nig = -1.f;
    for (int g = g0 + lane; g < g1; g += 32) {
      if (gt_labels[g] != pl) continue;
      const float iou = vid_iou(pb, __ldg(gt_boxes + g));
      const bool ign = gt_ignore[g] != 0;
      ++n_own;
      n_ign += ign ? 1 : 0;
      if (ign) ig = fmaxf(ig, iou); else nig = fmaxf(nig, iou);
      if (!gt_taken[g] && iou >= thr) best = fmaxf(best, iou);
    }
    n_own = warp_sum_i(n_own);
    unsigned char h = 0;
    double w;
    if (n_own == 0) {                         // no ground truth of this class in the image (vid_eval.py:219-222)
      w = empty_weight;
    } else {
      best = warp_max(best);
      if (best >= thr) {
        // pass 2, the reference's ascending scan over the ties (:236-252): the first tie that is not ignored, else the
        // last tie
        int first_reg = 0x7fffffff, last_tie = -1;
        for (int g = g0 + lane; g < g1; g += 32) {
          if (gt_labels[g] != pl || gt_taken[g]) continue;
          if (vid_iou(pb, __ldg(gt_boxes + g)) != best) continue;
          last_tie = g;
          if (!gt_ignore[g]) first_reg = min(first_reg, g);
        }
        first_reg = warp_min_i(first_reg);
        last_tie = warp_max_i(last_tie);
        const int k = first_reg != 0x7fffffff ? first_reg : last_tie;
        __syncwarp();
        if (lane == 0) gt_taken[k] = 1;
        __syncwarp();
        h = 1;
        w = gt_ignore[k] ? 1.0 : 0.0;
      } else {                                // unmatched (:258-264)
        ig = warp_max(ig);
        nig = warp_max(nig);
        n_ign = warp_sum_i(n_ign);
        w = nig > ig ? 0.0 : (ig > nig ? 1.0 : static_cast<double>(n_ign) / static_cast<double>(n_own));
      }
    }
    if (lane == 0) {
      hit[j] = h;
      weight[j] = w;
    }
  }
}

}  // namespace

int vid_match_launch(const float* pred_boxes, const int* pred_labels, const int* order, const int* pred_off,
                     const float* gt_boxes, const int* gt_labels, const unsigned char* gt_ignore, const int* gt_off,
                     int n_images, float thr, double empty_weight, unsigned char* gt_taken, unsigned char* hit,
                     double* weight, cudaStream_t stream) {
  if (n_images < 0) return DVID_ERR_SHAPE;
  if (n_images == 0) return DVID_OK;
  const unsigned blocks = static_cast<unsigned>((static_cast<long>(n_images) + 3) / 4);
  launch_pdl(vid_match_kernel, dim3(blocks), dim3(128), 0, stream, reinterpret_cast<const float4*>(pred_boxes),
             pred_labels, order, pred_off, reinterpret_cast<const float4*>(gt_boxes), gt_labels, gt_ignore, gt_off,
             n_images, thr, empty_weight, gt_taken, hit, weight);
  return check_launch();
}

}  // namespace dvid

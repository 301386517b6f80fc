"""`mega_core._C` look-alike: the native operator module the reference builds from mega_core/csrc (pybind names in
csrc/vision.cpp:10-27), re-exported over libdvid_b200.so so `mega_core/layers/{fps,nms}.py` keep working when their
`from mega_core import _C` is pointed here.

Only the operators that exist for sm_100a are provided:
  furthest_point_sampling(b, n, m, points, temp, idx) -> int      csrc/fps.h:15-36 (the one _C op DiffusionDet calls)
  nms(dets, scores, threshold) -> LongTensor                       csrc/nms.h:10-28, CUDA semantics of cuda/nms.cu
  roi_align_forward(input, rois, scale, ph, pw, sampling_ratio)    csrc/ROIAlign.h:11-27, cuda/ROIAlign_cuda.cu:65-125
The remaining legacy R-CNN ops (roi_align_backward, roi_pool_*, sigmoid_focalloss_*, deform_*) are training-time or
belong to other detectors and raise NotImplementedError here.  CUDA tensors only: there is no CPU path
in this package.
"""
import torch

from . import ops
from ._lib import DvidError


def furthest_point_sampling(b, n, m, points_tensor, temp_tensor, idx_tensor):
    """Writes `idx_tensor` (b,m) int32 in place; returns 1, or -1 for empty / CPU input like the reference."""
    if not points_tensor.is_cuda or points_tensor.numel() == 0:
        return -1
    return ops.furthest_point_sampling(b, n, m, points_tensor.contiguous(), temp_tensor, idx_tensor)


def nms(dets, scores, threshold):
    """Legacy maskrcnn-benchmark NMS: +1 pixel IoU, suppress IoU > threshold (cuda/nms.cu:13-67), kept ORIGINAL indices
    in ascending order (cuda/nms.cu:127-130).  dets (n,4) fp32 CUDA, scores (n) fp32 CUDA, n <= 1024."""
    if not dets.is_cuda:
        raise DvidError("diffusionvid_b200._C_shim.nms: CUDA tensors only (no CPU path)")
    n = dets.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=dets.device)
    r = ops.nms(dets.float().contiguous().view(1, n, 4), scores.float().contiguous().view(1, n), thr=float(threshold),
                plus_one=True, ge=False, ascending_out=True, want_compact=False)
    c = int(r["count"][0].item())
    return r["keep"][0, :c]


def roi_align_forward(input, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """Legacy maskrcnn-benchmark ROIAlign forward (csrc/cuda/ROIAlign_cuda.cu:65-125): no half-pixel shift, roi size
    clamped to >= 1, NCHW.  input (N,C,H,W) fp32 CUDA, rois (n,5) = (batch index, x1, y1, x2, y2) -> (n,C,ph,pw).
    The reference wraps the call in apex `amp.float_function` (layers/roi_align.py:57), so inputs arrive as fp32; other
    floating dtypes are converted here the same way."""
    if not input.is_cuda or not rois.is_cuda:
        raise DvidError("diffusionvid_b200._C_shim.roi_align_forward: CUDA tensors only (no CPU path)")
    out = ops.roi_align_legacy_forward(input.float().contiguous(), rois.float().contiguous(), float(spatial_scale),
                                       int(pooled_height), int(pooled_width), int(sampling_ratio))
    return out.to(input.dtype)


def _absent(name):
    def f(*a, **k):
        raise NotImplementedError(
            "mega_core._C.%s is a legacy R-CNN operator that the DiffusionVID path never calls; it is not part of "
            "libdvid_b200.so (SURVEY.md 2.2)" % name)
    f.__name__ = name
    return f


for _n in ("roi_align_backward", "roi_pool_forward", "roi_pool_backward",
           "sigmoid_focalloss_forward", "sigmoid_focalloss_backward", "deform_conv_forward",
           "deform_conv_backward_input", "deform_conv_backward_parameters", "modulated_deform_conv_forward",
           "modulated_deform_conv_backward", "deform_psroi_pooling_forward", "deform_psroi_pooling_backward"):
    globals()[_n] = _absent(_n)

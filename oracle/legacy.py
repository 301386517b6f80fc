"""numpy restatement of the reference's legacy native NMS (mega_core._C.nms).  TEST INFRASTRUCTURE ONLY.

Follows mega_core/csrc/cpu/nms_cpu.cpp:5-65 (CPU: areas and overlaps with the legacy +1 pixel convention, suppress when
`ovr >= threshold`, return the surviving ORIGINAL indices in ascending order) and mega_core/csrc/cuda/nms.cu:13-21,
:112-130 (CUDA: same IoU, suppress when `> threshold`).  Pinned by the reference's own tests/test_nms.py vectors
(tests/golden/nms_vectors.json, extracted by tests/golden/extract_nms_vectors.py which executes that test file)."""
import numpy as np


def nms_legacy(dets, scores, threshold, ge=True):
    dets = np.asarray(dets, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    areas = (x2 - x1 + np.float32(1)) * (y2 - y1 + np.float32(1))
    order = np.argsort(-scores, kind="stable")
    suppressed = np.zeros(n, dtype=np.uint8)
    thr = np.float32(threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        for _j in range(_i + 1, n):
            j = order[_j]
            if suppressed[j]:
                continue
            w = max(np.float32(0), min(x2[i], x2[j]) - max(x1[i], x1[j]) + np.float32(1))
            h = max(np.float32(0), min(y2[i], y2[j]) - max(y1[i], y1[j]) + np.float32(1))
            inter = np.float32(w * h)
            ovr = inter / (areas[i] + areas[j] - inter)
            if (ovr >= thr) if ge else (ovr > thr):
                suppressed[j] = 1
    return np.nonzero(suppressed == 0)[0].astype(np.int64)


def roi_align_legacy(inp, rois, spatial_scale, pooled_h, pooled_w, sampling_ratio):
    """numpy restatement of the reference's legacy ROIAlign forward (mega_core._C.roi_align_forward):
    mega_core/csrc/cuda/ROIAlign_cuda.cu:15-62 (bilinear_interpolate) and :65-125 (RoIAlignForward) - no half-pixel shift,
    roi width/height clamped to >= 1, adaptive sampling grid ceil(roi/pooled) when sampling_ratio <= 0, fp32 throughout.
    inp (N,C,H,W) float32, rois (n,5) float32 = (batch index, x1, y1, x2, y2) -> (n,C,ph,pw) float32.
    Pinned in tests/test_oracle_ops.py against the installed torchvision.ops.roi_align(aligned=False), the same
    caffe2-derived algorithm the reference's kernel was taken from."""
    f32 = np.float32
    inp = np.asarray(inp, dtype=f32)
    rois = np.asarray(rois, dtype=f32)
    N, C, H, W = inp.shape
    n = rois.shape[0]
    out = np.zeros((n, C, pooled_h, pooled_w), dtype=f32)
    scale = f32(spatial_scale)
    for i in range(n):
        b = int(rois[i, 0])
        x1, y1, x2, y2 = (f32(rois[i, k] * scale) for k in (1, 2, 3, 4))
        rw = max(f32(x2 - x1), f32(1))
        rh = max(f32(y2 - y1), f32(1))
        bh = f32(rh / f32(pooled_h))
        bw = f32(rw / f32(pooled_w))
        gh = sampling_ratio if sampling_ratio > 0 else int(np.ceil(rh / f32(pooled_h)))
        gw = sampling_ratio if sampling_ratio > 0 else int(np.ceil(rw / f32(pooled_w)))
        for ph in range(pooled_h):
            for pw in range(pooled_w):
                acc = np.zeros((C,), dtype=f32)
                for iy in range(gh):
                    y = f32(f32(y1 + f32(f32(ph) * bh)) + f32(f32(f32(iy + 0.5) * bh) / f32(gh)))
                    for ix in range(gw):
                        x = f32(f32(x1 + f32(f32(pw) * bw)) + f32(f32(f32(ix + 0.5) * bw) / f32(gw)))
                        yy, xx = y, x
                        if yy < -1.0 or yy > H or xx < -1.0 or xx > W:
                            continue
                        yy = max(yy, f32(0))
                        xx = max(xx, f32(0))
                        yl, xl = int(yy), int(xx)
                        if yl >= H - 1:
                            yh = yl = H - 1
                            yy = f32(yl)
                        else:
                            yh = yl + 1
                        if xl >= W - 1:
                            xh = xl = W - 1
                            xx = f32(xl)
                        else:
                            xh = xl + 1
                        ly, lx = f32(yy - f32(yl)), f32(xx - f32(xl))
                        hy, hx = f32(f32(1) - ly), f32(f32(1) - lx)
                        val = f32(f32(f32(f32(hy * hx) * inp[b, :, yl, xl] + f32(hy * lx) * inp[b, :, yl, xh])
                                      + f32(ly * hx) * inp[b, :, yh, xl]) + f32(ly * lx) * inp[b, :, yh, xh])
                        acc = (acc + val).astype(f32)
                out[i, :, ph, pw] = acc / f32(gh * gw)
    return out

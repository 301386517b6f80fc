"""CPU restatement of the frame resize of the reference's test transform.  TEST INFRASTRUCTURE ONLY (see
oracle/__init__.py).

Reference: `Resize(min_size, max_size)` (mega_core/data/transforms/transforms.py:31-67) computes the output size and
calls torchvision `F.resize(pil_image, size)`, i.e. Pillow's `Image.resize(size, BILINEAR)` - an antialiased two-pass
(horizontal, then vertical) triangle filter evaluated in 8-bit fixed point.  Pillow is a third-party dependency
(not in /root/reference; no pin in the reference, INSTALL.md installs whatever pip brings); its algorithm
(libImaging/Resample.c: precompute_coeffs, normalize_coeffs_8bpc, ImagingResampleHorizontal/Vertical_8bpc) is restated
here from its published source and PINNED against the Pillow installed in this container (12.2.0) in
tests/test_oracle_resize.py: bit-exact on every tested shape.

    support = max(scale, 1), ksize = 2 * ceil(support) + 1
    for every output coordinate: center = (xx + .5) * scale; window [int(center - support + .5), int(center + support
    + .5)) clipped to the image; triangle weights w(x) = max(0, 1 - |x + .5 - center| / max(scale, 1)), normalised in
    double precision, then rounded to 22-bit fixed point; a pass accumulates from 2^21 (round half up), shifts right by 22
    and clamps to [0, 255].
"""
import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2


def get_size(image_size, min_size=600, max_size=1000):
    """transforms.py:38-59 for one min_size: image_size (w, h) -> (oh, ow)."""
    w, h = image_size
    size = min_size
    if max_size is not None:
        mn, mx = float(min(w, h)), float(max(w, h))
        if mx / mn * size > max_size:
            size = int(round(max_size * mn / mx))
    if (w <= h and w == size) or (h <= w and h == size):
        return (h, w)
    if w < h:
        return (int(size * h / w), size)
    return (size, int(size * w / h))


def coefficients(in_size, out_size):
    """-> (bounds int32 [out, 2] = (first input index, count), coeffs int32 [out, ksize]) of one axis."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 1.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    ss = 1.0 / fscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        x = np.arange(xmax, dtype=np.float64)
        w = np.maximum(0.0, 1.0 - np.abs((x + xmin - center + 0.5) * ss))
        ww = 0.0
        for v in w:                      # same left-to-right double accumulation as the C loop
            ww += v
        if ww != 0.0:
            w = w / ww
        kk[xx, :xmax] = w
        bounds[xx] = (xmin, xmax)
    fixed = np.where(kk < 0, np.trunc(-0.5 + kk * (1 << PRECISION_BITS)), np.trunc(0.5 + kk * (1 << PRECISION_BITS)))
    return bounds, fixed.astype(np.int32)


def _pass(img, bounds, coeffs, axis):
    """one resampling pass over `axis` of a uint8 array; int64 accumulation (values fit 32 bits like in C)."""
    src = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + src.shape[1:], dtype=np.uint8)
    for xx in range(bounds.shape[0]):
        x0, n = int(bounds[xx, 0]), int(bounds[xx, 1])
        acc = np.full(src.shape[1:], 1 << (PRECISION_BITS - 1), dtype=np.int64)
        for x in range(n):
            acc += src[x0 + x] * int(coeffs[xx, x])
        out[xx] = np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


def resize_bilinear_u8(img_hwc, oh, ow):
    """Pillow Image.resize((ow, oh), BILINEAR) of a uint8 [H, W, C] array: horizontal pass (skipped when the width is
    unchanged), then vertical pass (skipped when the height is unchanged)."""
    h, w = img_hwc.shape[:2]
    out = img_hwc
    if ow != w:
        out = _pass(out, *coefficients(w, ow), axis=1)
    if oh != h:
        out = _pass(out, *coefficients(h, oh), axis=0)
    return out

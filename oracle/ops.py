"""Operator-level restatements (pure PyTorch / numpy, CPU) of the third-party and native ops on the DiffusionVID hot
path.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Each function cites what it restates.  detectron2 / torchvision sources are not under /root/reference; their
semantics are taken from SURVEY.md Appendix A and pinned against the installed torchvision in tests/test_oracle_ops.py.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------------------------------ ROIAlign / ROIPooler
def roi_align(feat, rois, out_size, scale, sampling_ratio):
    """torchvision.ops.roi_align(feat, rois, out_size, scale, sampling_ratio, aligned=True) restated.

    Called by the reference through detectron2 ROIPooler at mega_core/modeling/roi_heads/box_head/box_head.py:507,617.
    Same bilinear routine as mega_core/csrc/cuda/ROIAlign_cuda.cu:15-62 plus the aligned=True half-pixel shift and no
    clamp of the roi size.  feat (B,C,H,W) fp32, rois (K,5) [batch,x1,y1,x2,y2] -> (K,C,out,out) fp32.
    """
    K = rois.shape[0]
    B, C, H, W = feat.shape
    P = out_size
    S = sampling_ratio
    if K == 0:
        return feat.new_zeros((0, C, P, P))
    bidx = rois[:, 0].long()
    x1 = rois[:, 1] * scale - 0.5
    y1 = rois[:, 2] * scale - 0.5
    x2 = rois[:, 3] * scale - 0.5
    y2 = rois[:, 4] * scale - 0.5
    bin_w = (x2 - x1) / P
    bin_h = (y2 - y1) / P
    ar = torch.arange(P * S, device=feat.device)
    pbin = (ar // S).to(feat.dtype)[None, :]                               # ph for each of the P*S sample rows
    frac = ((ar % S).to(feat.dtype) + 0.5)[None, :]                        # iy + .5
    # torchvision: y = roi_start_h + ph * bin_size_h + (iy + .5f) * bin_size_h / roi_bin_grid_h  (same op order)
    ys = y1[:, None] + pbin * bin_h[:, None] + frac * bin_h[:, None] / S   # (K, P*S)
    xs = x1[:, None] + pbin * bin_w[:, None] + frac * bin_w[:, None] / S

    def prep(v, size):
        invalid = (v < -1.0) | (v > size)
        v = v.clamp(min=0)
        low = v.floor().long()
        top = low >= size - 1
        low = torch.where(top, torch.full_like(low, size - 1), low)
        high = torch.where(top, torch.full_like(low, size - 1), low + 1)
        v = torch.where(top, low.to(v.dtype), v)
        l = v - low.to(v.dtype)
        h = 1.0 - l
        return low, high, l, h, invalid

    y_lo, y_hi, ly, hy, y_bad = prep(ys, H)
    x_lo, x_hi, lx, hx, x_bad = prep(xs, W)
    hy = torch.where(y_bad, torch.zeros_like(hy), hy)
    ly = torch.where(y_bad, torch.zeros_like(ly), ly)
    hx = torch.where(x_bad, torch.zeros_like(hx), hx)
    lx = torch.where(x_bad, torch.zeros_like(lx), lx)

    out = feat.new_zeros((K, C, P, P))
    # chunk over rois to bound memory: gather (k, C, PS, PS) at a time
    step = max(1, int(2 ** 24 // (C * (P * S) ** 2)))
    for s in range(0, K, step):
        e = min(K, s + step)
        f = feat[bidx[s:e]]                                               # (k,C,H,W)
        k = e - s

        def g(yi, xi):
            idx = (yi[:, :, None] * W + xi[:, None, :]).view(k, 1, -1).expand(k, C, -1)
            return f.reshape(k, C, H * W).gather(2, idx).view(k, C, P * S, P * S)

        v = (hy[s:e, None, :, None] * hx[s:e, None, None, :]) * g(y_lo[s:e], x_lo[s:e]) + \
            (hy[s:e, None, :, None] * lx[s:e, None, None, :]) * g(y_lo[s:e], x_hi[s:e]) + \
            (ly[s:e, None, :, None] * hx[s:e, None, None, :]) * g(y_hi[s:e], x_lo[s:e]) + \
            (ly[s:e, None, :, None] * lx[s:e, None, None, :]) * g(y_hi[s:e], x_hi[s:e])
        out[s:e] = v.view(k, C, P, S, P, S).sum(dim=(3, 5)) / (S * S)
    return out


def assign_levels(boxes, min_level=3, max_level=5, canonical_size=224, canonical_level=4):
    """detectron2 poolers.assign_boxes_to_levels: level index (0-based from min_level) per box."""
    size = torch.sqrt((boxes[:, 2] - boxes[:, 0]) * (boxes[:, 3] - boxes[:, 1]))
    lvl = torch.floor(canonical_level + torch.log2(size / canonical_size + 1e-8))
    lvl = torch.clamp(lvl, min=min_level, max=max_level)
    return lvl.to(torch.int64) - min_level


def roi_pooler(feats, boxes, scales=(1 / 8., 1 / 16., 1 / 32.), out_size=7, sampling_ratio=2):
    """detectron2 ROIPooler.forward as configured at box_head.py:250-271.

    feats: list of (B,C,H_l,W_l); boxes: (B,N,4) absolute xyxy -> (B*N, C, 7, 7), row = b*N + n.
    """
    B, N = boxes.shape[:2]
    flat = boxes.reshape(-1, 4)
    bcol = torch.arange(B, dtype=flat.dtype, device=flat.device).repeat_interleave(N)[:, None]
    rois = torch.cat([bcol, flat], dim=1)
    lvl = assign_levels(flat, 3, 3 + len(feats) - 1)
    out = feats[0].new_zeros((B * N, feats[0].shape[1], out_size, out_size))
    for l, (f, sc) in enumerate(zip(feats, scales)):
        idx = torch.nonzero(lvl == l).squeeze(1)
        if idx.numel():
            out[idx] = roi_align(f, rois[idx], out_size, sc, sampling_ratio)
    return out


# ------------------------------------------------------------------------------------------------ attention
def mha(query, key, value, in_w, in_b, out_w, out_b, nheads, lin=None):
    """torch.nn.MultiheadAttention(E, nheads, dropout=0, batch_first=False).forward(...)[0] restated.

    query (L,Bt,E), key/value (S,Bt,E).  `lin(x, w, b)` lets the caller swap the projection arithmetic (fp16 emulation).
    Used by the reference at box_head.py:516,626 (self-attention) and :371 (global cross-attention).
    """
    if lin is None:
        lin = F.linear
    L, Bt, E = query.shape
    S = key.shape[0]
    hd = E // nheads
    q = lin(query.reshape(L * Bt, E), in_w[:E], in_b[:E]).view(L, Bt, nheads, hd)
    k = lin(key.reshape(S * Bt, E), in_w[E:2 * E], in_b[E:2 * E]).view(S, Bt, nheads, hd)
    v = lin(value.reshape(S * Bt, E), in_w[2 * E:], in_b[2 * E:]).view(S, Bt, nheads, hd)
    q = q * (1.0 / math.sqrt(hd))
    att = torch.einsum("lbhd,sbhd->bhls", q, k)
    att = torch.softmax(att, dim=-1)
    ctx = torch.einsum("bhls,sbhd->lbhd", att, v).reshape(L * Bt, E)
    return lin(ctx, out_w, out_b).view(L, Bt, E), ctx.view(L, Bt, E)


# ------------------------------------------------------------------------------------------------ NMS
def box_iou_row(box, others):
    """IoU of one box vs many, torchvision convention (no +1): inter / (a + b - inter)."""
    lt = torch.maximum(box[:2], others[:, :2])
    rb = torch.minimum(box[2:], others[:, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[:, 0] * wh[:, 1]
    a = (box[2] - box[0]) * (box[3] - box[1])
    b = (others[:, 2] - others[:, 0]) * (others[:, 3] - others[:, 1])
    return inter / (a + b - inter)


def nms(boxes, scores, thr):
    """torchvision.ops.nms restated: greedy, descending score, suppress IoU > thr; returns kept original
    indices in descending-score order.  Ties in score are broken by ascending index (stable sort) - the canonical
    order this repo fixes (torchvision's CUDA sort is not stable; SURVEY.md 8c contract 3)."""
    n = boxes.shape[0]
    if n == 0:
        return torch.zeros((0,), dtype=torch.int64)
    order = torch.sort(scores, descending=True, stable=True)[1]
    b = boxes[order]
    alive = torch.ones(n, dtype=torch.bool)
    keep = []
    for i in range(n):
        if not alive[i]:
            continue
        keep.append(i)
        if i + 1 < n:
            iou = box_iou_row(b[i], b[i + 1:])
            alive[i + 1:] &= ~(iou > thr)
    return order[torch.tensor(keep, dtype=torch.int64)]


def batched_nms(boxes, scores, idxs, thr):
    """detectron2.layers.batched_nms -> torchvision.ops.batched_nms, coordinate-offset branch (the one taken on CUDA at
    the sizes of diffusion_det.py:617,793): boxes.float() + idxs * (boxes.max() + 1), then nms."""
    boxes = boxes.float()
    if boxes.numel() == 0:
        return torch.zeros((0,), dtype=torch.int64)
    max_coord = boxes.max()
    offsets = idxs.to(boxes) * (max_coord + torch.tensor(1).to(boxes))
    return nms(boxes + offsets[:, None], scores, thr)


# ------------------------------------------------------------------------------------------------ farthest point sampling
def fps_block_size(n):
    """opt_n_threads, mega_core/csrc/cuda/fps.cu:11-15."""
    p = int(math.log(float(n)) / math.log(2.0))
    return max(min(1 << p, 1024), 1)


def fps(dist, m):
    """Literal emulation of furthest_point_sampling_kernel (mega_core/csrc/cuda/fps.cu:25-142) for b=1.

    dist: (n,n) float32 numpy; returns int32 (m,) picks.  temp starts at 1e10 (diffusion_det.py:893); first pick is 0;
    each round: temp = min(temp, dist[old]); per-thread strided scan with strict '>' (lowest k of the thread wins),
    then the shared-memory tree where slot t takes slot t+s only if strictly greater (lowest slot wins ties).
    """
    dist = np.asarray(dist, dtype=np.float32)
    n = dist.shape[0]
    bs = fps_block_size(n)
    temp = np.full((n,), 1e10, dtype=np.float32)
    idx = np.zeros((m,), dtype=np.int32)
    old = 0
    ks = np.arange(n)
    slot = ks % bs
    for j in range(1, m):
        temp = np.minimum(dist[old], temp)
        best = np.full((bs,), -1.0, dtype=np.float32)
        besti = np.zeros((bs,), dtype=np.int64)
        # per-thread scan in increasing k: strict > keeps the first maximum
        for r in range((n + bs - 1) // bs):
            k = ks[r * bs:(r + 1) * bs]
            v = temp[k]
            t = slot[k]
            upd = v > best[t]
            besti[t] = np.where(upd, k, besti[t])
            best[t] = np.where(upd, v, best[t])
        s = bs // 2
        while s >= 1:
            v1, v2 = best[:s].copy(), best[s:2 * s]
            i1, i2 = besti[:s].copy(), besti[s:2 * s]
            best[:s] = np.maximum(v1, v2)
            besti[:s] = np.where(v2 > v1, i2, i1)
            s //= 2
        old = int(besti[0])
        idx[j] = old
    return idx


def cdist_l2(x):
    """torch.cdist(x, x, p=2.0) as called at diffusion_det.py:880, computed by direct differences in fp32
    (compute_mode='donot_use_mm_for_euclid_dist'): the repo's kernel computes the same sum((a-b)^2) then sqrt."""
    return torch.cdist(x, x, p=2.0, compute_mode="donot_use_mm_for_euclid_dist")

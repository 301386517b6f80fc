"""Shared helpers for end-to-end parity checks.

The detector contains discontinuous steps (FPN level assignment by floor(log2 sqrt(area)), per-frame top-k, the 0.5
renewal threshold, NMS, farthest-point sampling).  An fp16-ulp difference in one box can flip one of them and move that
box - and, through the global memory, nudge the rest - so end-to-end results are compared as *sets with tolerance*:
every oracle detection is matched to a product detection of the same label whose box is within `box_tol` (fraction
of the image size) and whose score is within `score_tol`; the test asserts on the matched fraction.  Operator- and
head-level tests (tests/test_gpu_ops.py, tests/test_gpu_model.py) carry the tight per-element tolerances."""
import torch


def match_fraction(got_boxes, got_scores, got_labels, ref_boxes, ref_scores, ref_labels, img_max, box_tol=1e-3,
                   score_tol=2e-3):
    """fraction of reference detections that have a counterpart in `got` (greedy one-to-one)."""
    n = ref_scores.numel()
    if n == 0:
        return 1.0 if got_scores.numel() == 0 else 0.0
    used = torch.zeros(got_scores.numel(), dtype=torch.bool)
    hit = 0
    for i in range(n):
        ok = (got_labels == ref_labels[i]) & ~used
        ok &= (got_boxes - ref_boxes[i]).abs().max(dim=1)[0] <= box_tol * img_max
        ok &= (got_scores - ref_scores[i]).abs() <= score_tol
        idx = torch.nonzero(ok)
        if idx.numel():
            used[idx[0, 0]] = True
            hit += 1
    return hit / n


def match_fraction_two_sided(got_boxes, got_scores, got_labels, ref_boxes, ref_scores, ref_labels, img_max,
                             box_tol=1e-3, score_tol=2e-3):
    """min(reference -> product, product -> reference): spurious extra detections (NMS under-suppression, stale rows
    past `count`) lower the score just like missing ones do."""
    fwd = match_fraction(got_boxes, got_scores, got_labels, ref_boxes, ref_scores, ref_labels, img_max, box_tol,
                         score_tol)
    bwd = match_fraction(ref_boxes, ref_scores, ref_labels, got_boxes, got_scores, got_labels, img_max, box_tol,
                         score_tol)
    return min(fwd, bwd)


def rows_within(a, b, tol):
    """fraction of rows of a/b (.., D) whose max-abs difference is <= tol."""
    d = (a - b).abs().reshape(-1, a.shape[-1]).max(dim=1)[0]
    return (d <= tol).float().mean().item()


# ---------------------------------------------------------------------------------------------------- step-level helpers
def quantiles(d):
    """p50 / p99 / p99.9 / max of a tensor of absolute differences."""
    d = d.flatten().float()
    if d.numel() == 0:
        return dict(p50=0.0, p99=0.0, p999=0.0, max=0.0)
    sub = d[torch.randperm(d.numel(), device=d.device)[:4_000_000]] if d.numel() > 4_000_000 else d
    qs = torch.quantile(sub, torch.tensor([0.5, 0.99, 0.999], device=d.device))
    return dict(p50=qs[0].item(), p99=qs[1].item(), p999=qs[2].item(), max=d.max().item())


def product_step(m, lv, boxes, t, B, N):
    """One DDIM step's head chain of the product, head by head (the body of DiffusionDet._decode's loop):
    [(input boxes, logits, output boxes, object features fp32)] for head_series[0..2] and head_series_cond[0]."""
    outs = []
    pro32 = pro16 = None
    for e in m._pk["heads"]:
        ins = boxes
        lg, boxes, pro32, pro16 = m._head(e, lv, boxes.contiguous(), pro32, pro16, t)
        outs.append((ins, lg, boxes, pro32))
    if m._pk["cond"]:
        cond16 = m._global_context(pro16.contiguous(), B * N)
        for e in m._pk["cond"]:
            ins = boxes
            lg, boxes, pro32, pro16 = m._head(e, lv, boxes.contiguous(), pro32.contiguous(), pro16.contiguous(), t,
                                              shift_rows=m._cond_shift(e, cond16, B * N))
            outs.append((ins, lg, boxes, pro32))
    return outs


def oracle_step(om, o, feats, boxes, t, B, mem):
    """The same chain through the oracle (oracle.model as `om`, OracleDiffusionVID `o`)."""
    temb = om.time_embedding(o.c, torch.full((B,), t, dtype=torch.long, device=boxes.device))
    outs = []
    pro = None
    for i in range(o.cfg["num_heads"]):
        ins = boxes
        lg, boxes, pro = om.rcnn_head(o.c, "head.head_series.%d." % i, feats, boxes, pro, temb, o.cfg)
        outs.append((ins, lg, boxes, pro))
    if o.cfg["global_enable"] and o.cfg["num_heads_local"] > 0:
        attn_ = om.global_attention(o.c, pro, mem, o.cfg)
        for hi in range(o.cfg["num_heads_local"]):
            ins = boxes
            lg, boxes, pro = om.rcnn_head(o.c, "head.head_series_cond.%d." % hi, feats, boxes, pro, temb, o.cfg,
                                          cond=attn_)
            outs.append((ins, lg, boxes, pro))
    return outs


def step_deltas(po, oo_, levels_fn, size):
    """Per head of a step: FPN-level disagreements accumulated so far, and box / logit differences of the boxes that were
    pooled from the same levels in both implementations (arithmetic differences only)."""
    n = po[0][1].shape[0] * po[0][1].shape[1]
    dev = po[0][1].device
    flipped = torch.zeros(n, dtype=torch.bool)
    per_head = []
    for (pin, plg, pbx, _), (oin, olg, obx, _) in zip(po, oo_):
        flipped |= levels_fn(pin.reshape(-1, 4).float().cpu()) != levels_fn(oin.reshape(-1, 4).float().cpu())
        ok = ~flipped.to(dev)
        dbox = (pbx - obx.to(dev)).abs().reshape(n, 4).max(dim=1)[0] / size
        dlog = (plg - olg.to(dev)).abs().reshape(n, -1).max(dim=1)[0]
        per_head.append(dict(level_flips_so_far=int(flipped.sum()), box_same_level=quantiles(dbox[ok]),
                             logit_same_level=quantiles(dlog[ok]),
                             box_flipped_max=float(dbox[~ok].max()) if bool((~ok).any()) else 0.0))
    return per_head

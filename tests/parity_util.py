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

"""ImageNet-VID detection metrics for the inference driver: AP50 / mAP and CorLoc over a list of
predicted and ground-truth BoxLists, as `mega_core/data/datasets/evaluation/vid/vid_eval.py` computes them for the
DiffusionVID configs (`do_vid_evaluation` :14-81 with `motion_specific=False`, `eval_detection_vid` :132-164,
`calc_detection_vid_prec_rec` :167-291, `calc_detection_vid_ap` :294-353, `corloc_eval_detection_vid` :355-436).

Organisation differs from the reference (which walks classes inside images inside Python lists): every image's
detections are matched in ONE pass in global score order against a label-masked IoU matrix (a detection only sees
ground truth of its own class, so the per-class greedy assignment is unchanged), the per-detection (score, class, hit)
records of all images are concatenated once, and precision/recall/AP are computed per class from one stable sort.

Conventions kept from the reference because they change numbers:
  * VID boxes are integer-typed: x2, y2 get +1 before the IoU (:224-228) and the IoU itself uses the legacy +1
    width/height (`boxlist_ops.py:83-88`), so a box is effectively w+2 wide.
  * a detection is assigned to the unassigned ground truth of highest IoU >= thresh, the first one on ties (:236-252).
  * classes that occur only in predictions have no recall -> AP = nan and are skipped by the mean (:281-289, :160).
  * CorLoc looks at the single top-scoring detection of an image and counts once per ground-truth BOX of a class
    (:377-417), not once per image.
Motion-specific AP (`motion_specific=True`, needs the dataset's motion-IoU .mat file) is not implemented.
"""
import os

import numpy as np
import torch

F32 = torch.float32


def _vid_iou(a, b):
    """IoU [len(a), len(b)] of xyxy float32 boxes under the VID integer-box convention (see module docstring)."""
    a = a.clone(); b = b.clone()
    a[:, 2:] += 1
    b[:, 2:] += 1
    area_a = (a[:, 2] - a[:, 0] + 1) * (a[:, 3] - a[:, 1] + 1)
    area_b = (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
    lt = torch.max(a[:, None, :2], b[:, :2])
    rb = torch.min(a[:, None, 2:], b[:, 2:])
    wh = (rb - lt + 1).clamp(min=0)
    inter = wh[:, :, 0] * wh[:, :, 1]
    return inter / (area_a[:, None] + area_b - inter)


def _fields(bl, device):
    return (bl.bbox.to(device, F32).reshape(-1, 4), bl.get_field("labels").to(device).long().reshape(-1),
            bl.get_field("scores").to(device, F32).reshape(-1) if bl.has_field("scores") else None)


def match_detections(pred_boxlists, gt_boxlists, iou_thresh=0.5, device="cpu"):
    """Greedy VID matching.  Returns (scores, labels, hits) of all detections concatenated in image order (within an
    image: descending score) and the number of ground-truth boxes per class (dict)."""
    if len(pred_boxlists) != len(gt_boxlists):
        raise ValueError("Length of gt and pred lists need to be same.")
    scores, labels, hits = [], [], []
    n_pos = {}
    for pred, gt in zip(pred_boxlists, gt_boxlists):
        pb, pl, ps = _fields(pred, device)
        gb, gl, _ = _fields(gt, device)
        for l, c in zip(*[t.tolist() for t in torch.unique(gl, return_counts=True)]):
            n_pos[l] = n_pos.get(l, 0) + c
        for l in torch.unique(pl).tolist():
            n_pos.setdefault(l, 0)
        if pb.shape[0] == 0:
            continue
        order = torch.sort(ps, descending=True, stable=True)[1]
        pb, pl, ps = pb[order], pl[order], ps[order]
        hit = torch.zeros(pb.shape[0], dtype=torch.bool, device=pb.device)
        if gb.shape[0] > 0:
            iou = _vid_iou(pb, gb)
            iou = torch.where(pl[:, None] == gl[None, :], iou, torch.full_like(iou, -1.0))   # own class only
            iou = iou.cpu()
            free = torch.ones(gb.shape[0], dtype=torch.bool)
            for j in range(iou.shape[0]):
                cand = torch.where(free & (iou[j] >= iou_thresh), iou[j], torch.full_like(iou[j], -1.0))
                k = int(torch.argmax(cand))                  # first maximum, like the reference's scan
                if cand[k] >= iou_thresh:
                    free[k] = False
                    hit[j] = True
        scores.append(ps.cpu()); labels.append(pl.cpu()); hits.append(hit.cpu())
    cat = lambda xs, dt: torch.cat(xs) if xs else torch.zeros(0, dtype=dt)
    return cat(scores, F32), cat(labels, torch.int64), cat(hits, torch.bool), n_pos


def precision_recall(scores, labels, hits, n_pos):
    """Per-class precision / recall arrays (index = class id; None where the class never occurs / has no ground
    truth), vid_eval.py:265-291 with no ignored boxes."""
    n_cls = (max(n_pos) + 1) if n_pos else 0
    prec, rec = [None] * n_cls, [None] * n_cls
    scores, hits = scores.numpy().astype(np.float64), hits.numpy()
    labels = labels.numpy()
    for l in n_pos:
        sel = labels == l
        order = np.argsort(-scores[sel], kind="stable")
        h = hits[sel][order]
        tp = np.cumsum(h)
        fp = np.cumsum(~h)
        prec[l] = tp / (fp + tp + np.spacing(1))
        if n_pos[l] > 0:
            rec[l] = tp / n_pos[l]
    return prec, rec


def average_precision(prec, rec, use_07_metric=False):
    """vid_eval.py:294-353: area under the monotone precision envelope (or the VOC07 11-point mean)."""
    ap = np.full(len(prec), np.nan)
    for l, (p, r) in enumerate(zip(prec, rec)):
        if p is None or r is None:
            continue
        p = np.nan_to_num(p)
        if use_07_metric:
            ap[l] = sum((p[r >= t].max() if (r >= t).any() else 0.0) / 11 for t in np.arange(0.0, 1.1, 0.1))
        else:
            mpre = np.concatenate(([0.0], p, [0.0]))
            mrec = np.concatenate(([0.0], r, [1.0]))
            mpre = np.maximum.accumulate(mpre[::-1])[::-1]
            i = np.where(mrec[1:] != mrec[:-1])[0]
            ap[l] = np.sum((mrec[i + 1] - mrec[i]) * mpre[i + 1])
    return ap


def eval_detection_vid(pred_boxlists, gt_boxlists, iou_thresh=0.5, use_07_metric=False, device="cpu"):
    """-> {"ap": per-class AP array (nan where undefined, index 0 = background), "map": nanmean}."""
    prec, rec = precision_recall(*match_detections(pred_boxlists, gt_boxlists, iou_thresh, device))
    ap = average_precision(prec, rec, use_07_metric)
    return {"ap": ap, "map": float(np.nanmean(ap)) if len(ap) else float("nan")}


def corloc_eval_detection_vid(pred_boxlists, gt_boxlists, iou_thresh=0.5, device="cpu"):
    """-> ({class: CorLoc}, mean CorLoc); see the module docstring for the counting rule."""
    if len(pred_boxlists) != len(gt_boxlists):
        raise ValueError("Length of gt and pred lists need to be same.")
    n_gt, n_ok = {}, {}
    for pred, gt in zip(pred_boxlists, gt_boxlists):
        pb, pl, ps = _fields(pred, device)
        gb, gl, _ = _fields(gt, device)
        top = int(torch.argmax(ps)) if ps.numel() else -1
        for l in gl.tolist():
            n_gt[l] = n_gt.get(l, 0) + 1
            n_ok.setdefault(l, 0)
            if top >= 0 and int(pl[top]) == l:
                if bool((_vid_iou(pb[top:top + 1], gb[gl == l]) > iou_thresh).any()):
                    n_ok[l] += 1
    corloc = {l: n_ok[l] / n_gt[l] for l in n_gt}
    return corloc, (sum(corloc.values()) / len(corloc) if corloc else float("nan"))


def do_vid_evaluation(dataset, predictions, output_folder=None, logger=None, device="cpu"):
    """vid_eval.py:14-81 for the branch the DiffusionVID configs take (box_only=False, motion_specific=False):
    predictions are resized to the original image size, AP50 per class + mAP + CorLoc are logged and written to
    `result.txt`.  `dataset` provides get_img_info(i) -> {"width","height"}, get_groundtruth(i) -> BoxList and
    map_class_id_to_class_name(i)."""
    preds, gts = [], []
    for i, p in enumerate(predictions):
        info = dataset.get_img_info(i)
        preds.append(p.resize((info["width"], info["height"])))
        gts.append(dataset.get_groundtruth(i))
    res = eval_detection_vid(preds, gts, 0.5, False, device)
    corloc, corloc_avg = corloc_eval_detection_vid(preds, gts, 0.5, device)
    s = "AP50 | motion={:>6s} = {:0.4f}\n".format("all", res["map"]) + "Category AP:\n"
    for i, ap in enumerate(res["ap"]):
        if i > 0:
            s += "{:<16}: {:.4f}\n".format(dataset.map_class_id_to_class_name(i), ap)
    s += "Mean CorLoc: {:.4f}\nCategory CorLoc:\n".format(corloc_avg)
    for l, v in corloc.items():
        s += "{:<16}: {:.4f}\n".format(dataset.map_class_id_to_class_name(l), v)
    if logger is not None:
        logger.info("\n" + s)
    if output_folder:
        with open(os.path.join(output_folder, "result.txt"), "w") as f:
            f.write(s)
    return dict(res, corloc=corloc, corloc_avg=corloc_avg, text=s)

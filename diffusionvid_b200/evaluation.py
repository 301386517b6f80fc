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
`device="cuda"` runs the matching of all images as ONE kernel launch (csrc/vid_match.cu, `_match_detections_gpu`):
40 000 frames/s with packing against 370 frames/s for the loop below (tools/bench_vid_match.py), identical records.
Motion-specific AP (`motion_specific=True`, vid_eval.py:39-44,142-149,172-181,192-197,233-264,279-283): every ground
truth box carries a "motion IoU" (how much it moves over +-10 frames; shipped with the dataset as
vid_groundtruth_motion_iou.mat); for a range [lo, hi] the boxes outside it are *ignored*: they do not count as
positives, a detection matched to one is neither true nor false positive, an unmatched detection is dropped when its
best overlap is with an ignored box, counted fully when it is with a regular one and with the image's ignored fraction
on a tie, and detections on images without ground truth of their class weigh `empty_weight` = the share of all boxes
inside the range.  `load_motion_ious` reads the .mat file the way the reference does.
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


def load_motion_ious(mat_file):
    """vid_eval.py:142-147: the dataset's per-image lists of ground-truth motion IoUs (empty entries -> 0)."""
    import scipy.io as sio
    m = sio.loadmat(mat_file)["motion_iou"]
    return [[(m[i][0][j][0] if len(m[i][0][j]) != 0 else 0) for j in range(len(m[i][0]))] for i in range(len(m))]


def match_detections(pred_boxlists, gt_boxlists, iou_thresh=0.5, device="cpu", motion_ious=None,
                     motion_range=(0.0, 1.0)):
    """Greedy VID matching.  Returns (scores, labels, hits, ignore) of all detections concatenated in image order
    (within an image: descending score) and the number of countable ground-truth boxes per class (dict).
    `ignore` is the reference's `pred_ignore` (vid_eval.py:170,221,254-263): 0 = regular, 1 = dropped, in between =
    fractional false-positive weight; all zeros without `motion_ious`."""
    if len(pred_boxlists) != len(gt_boxlists):
        raise ValueError("Length of gt and pred lists need to be same.")
    lo, hi = motion_range
    if motion_ious is None:
        motion_ious = [None] * len(gt_boxlists)
        empty_weight = 0.0
    else:
        allm = np.concatenate([np.asarray(m, dtype=np.float64).reshape(-1) for m in motion_ious], axis=0)
        empty_weight = float(((allm >= lo) & (allm <= hi)).sum()) / float(len(allm))
        if empty_weight == 1:
            empty_weight = 0.0
    if torch.device(device).type == "cuda":
        return _match_detections_gpu(pred_boxlists, gt_boxlists, iou_thresh, torch.device(device), motion_ious, lo, hi,
                                     empty_weight)
    scores, labels, hits, ignores = [], [], [], []
    n_pos = {}
    for pred, gt, miou in zip(pred_boxlists, gt_boxlists, motion_ious):
        pb, pl, ps = _fields(pred, device)
        gb, gl, _ = _fields(gt, device)
        ign = torch.zeros(gb.shape[0], dtype=torch.bool)
        if miou is not None and len(miou):                     # `if motion_iou:` on the per-image list (:192)
            mi = torch.as_tensor(np.asarray(miou, dtype=np.float64))
            ign = (mi < lo) | (mi > hi)
        glc = gl.cpu()
        for l in torch.unique(glc).tolist():
            own = glc == l
            n_pos[l] = n_pos.get(l, 0) + int(own.sum()) - int((own & ign).sum())
        for l in torch.unique(pl).tolist():
            n_pos.setdefault(l, 0)
        if pb.shape[0] == 0:
            continue
        order = torch.sort(ps, descending=True, stable=True)[1]
        pb, pl, ps = pb[order], pl[order], ps[order]
        hit = torch.zeros(pb.shape[0], dtype=torch.bool)
        pig = torch.zeros(pb.shape[0], dtype=torch.float64)
        plc = pl.cpu()
        if gb.shape[0] > 0:
            iou = _vid_iou(pb, gb)
            iou = torch.where(pl[:, None] == gl[None, :], iou, torch.full_like(iou, -1.0)).cpu()   # own class only
            free = torch.ones(gb.shape[0], dtype=torch.bool)
        for j in range(pb.shape[0]):
            own = glc == plc[j] if gb.shape[0] > 0 else torch.zeros(0, dtype=torch.bool)
            n_own = int(own.sum())
            if n_own == 0:                                       # no ground truth of this class in the image (:219-222)
                pig[j] = empty_weight
                continue
            row = iou[j]
            cand = free & own & (row >= iou_thresh)
            if bool(cand.any()):
                best = row[cand].max()
                ties = torch.nonzero(cand & (row == best)).flatten().tolist()
                # the reference's ascending scan (:236-252): a later tie replaces the current pick while that is ignored
                k = ties[0]
                for t in ties[1:]:
                    if bool(ign[k]):
                        k = t
                free[k] = False
                hit[j] = True
                pig[j] = float(ign[k])
            else:                                                # unmatched (:258-264)
                ig = row[own & ign].max().item() if bool((own & ign).any()) else -1.0
                nig = row[own & ~ign].max().item() if bool((own & ~ign).any()) else -1.0
                pig[j] = 0.0 if nig > ig else (1.0 if ig > nig else float((own & ign).sum()) / float(n_own))
        scores.append(ps.cpu()); labels.append(plc); hits.append(hit); ignores.append(pig)
    cat = lambda xs, dt: torch.cat(xs) if xs else torch.zeros(0, dtype=dt)
    return (cat(scores, F32), cat(labels, torch.int64), cat(hits, torch.bool), cat(ignores, torch.float64)), n_pos


def _match_detections_gpu(pred_boxlists, gt_boxlists, iou_thresh, device, motion_ious, lo, hi, empty_weight):
    """`match_detections` with every image matched by ONE launch of the dvid_vid_match kernel (csrc/vid_match.cu): the
    BoxLists are packed into flat tensors (boxes, labels, per-image offsets), the per-image descending-score order is
    two stable device sorts, the positives per class are two bincounts.  Same records, bit for bit, as the CPU loop
    (tests/test_gpu_ops.py::test_vid_match_equals_the_cpu_evaluator)."""
    from . import ops
    i32 = torch.int32
    n_img = len(pred_boxlists)
    pcnt = torch.tensor([len(p) for p in pred_boxlists], dtype=torch.int64)
    gcnt = torch.tensor([len(g) for g in gt_boxlists], dtype=torch.int64)
    cat = lambda xs, shape, dt: (torch.cat(xs) if xs else torch.zeros(shape, dtype=dt))
    pb = cat([p.bbox.reshape(-1, 4).to(F32) for p in pred_boxlists], (0, 4), F32).to(device).contiguous()
    pl = cat([p.get_field("labels").reshape(-1).long() for p in pred_boxlists], (0,), torch.int64).to(device)
    ps = cat([p.get_field("scores").reshape(-1).to(F32) for p in pred_boxlists], (0,), F32).to(device)
    gb = cat([g.bbox.reshape(-1, 4).to(F32) for g in gt_boxlists], (0, 4), F32).to(device).contiguous()
    gl = cat([g.get_field("labels").reshape(-1).long() for g in gt_boxlists], (0,), torch.int64).to(device)
    ign = []
    for g, miou in zip(gt_boxlists, motion_ious):
        if miou is not None and len(miou):
            mi = np.asarray(miou, dtype=np.float64).reshape(-1)
            if mi.shape[0] != len(g):
                raise ValueError("motion IoU list and ground-truth BoxList differ in length")
            ign.append(torch.from_numpy((mi < lo) | (mi > hi)))
        else:
            ign.append(torch.zeros(len(g), dtype=torch.bool))
    gi = cat(ign, (0,), torch.bool).to(device)
    poff = torch.zeros(n_img + 1, dtype=torch.int64)
    poff[1:] = torch.cumsum(pcnt, 0)
    goff = torch.zeros(n_img + 1, dtype=torch.int64)
    goff[1:] = torch.cumsum(gcnt, 0)
    img_of = torch.repeat_interleave(torch.arange(n_img), pcnt).to(device)
    o1 = torch.sort(ps, descending=True, stable=True)[1]
    order = o1[torch.sort(img_of[o1], stable=True)[1]]
    hit, weight = ops.vid_match(pb, pl.to(i32), order.to(i32), poff.to(device, i32), gb, gl.to(i32),
                                gi.to(torch.uint8), goff.to(device, i32), iou_thresh, empty_weight)
    # countable positives per class: ground truth inside the range; classes seen only in predictions count 0
    n_pos = {}
    n_cls = int(max(gl.max().item() if gl.numel() else -1, pl.max().item() if pl.numel() else -1)) + 1
    if n_cls > 0:
        g_all = torch.bincount(gl, minlength=n_cls).cpu()
        g_reg = torch.bincount(gl[~gi], minlength=n_cls).cpu()
        p_all = torch.bincount(pl, minlength=n_cls).cpu()
        for l in range(n_cls):
            if g_all[l] > 0:
                n_pos[l] = int(g_reg[l])
            elif p_all[l] > 0:
                n_pos[l] = 0
    return (ps[order].cpu(), pl[order].cpu(), hit.bool().cpu(), weight.cpu()), n_pos


def precision_recall(records, n_pos):
    """Per-class precision / recall arrays (index = class id; None where the class never occurs / has no countable
    ground truth), vid_eval.py:265-291."""
    scores, labels, hits, ignores = records
    n_cls = (max(n_pos) + 1) if n_pos else 0
    prec, rec = [None] * n_cls, [None] * n_cls
    scores, hits = scores.numpy().astype(np.float64), hits.numpy()
    labels, ignores = labels.numpy(), ignores.numpy()
    for l in n_pos:
        sel = labels == l
        order = np.argsort(-scores[sel], kind="stable")
        h = hits[sel][order]
        w = ignores[sel][order].copy()
        tps = np.logical_and(h, w != 1)
        fps = np.logical_and(~h, w != 1)
        w[w == 0] = 1
        tp = np.cumsum(tps)
        fp = np.cumsum(fps * w)
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


MOTION_RANGES = (("all", (0.0, 1.0)), ("fast", (0.0, 0.7)), ("medium", (0.7, 0.9)), ("slow", (0.9, 1.0)))   # :40-41


def eval_detection_vid(pred_boxlists, gt_boxlists, iou_thresh=0.5, use_07_metric=False, device="cpu",
                       motion_ious=None, motion_range=(0.0, 1.0)):
    """-> {"ap": per-class AP array (nan where undefined, index 0 = background), "map": nanmean}.  With `motion_ious`
    (per image: one value per ground-truth box) the boxes outside `motion_range` are ignored (motion-specific AP)."""
    prec, rec = precision_recall(*match_detections(pred_boxlists, gt_boxlists, iou_thresh, device, motion_ious,
                                                   motion_range))
    ap = average_precision(prec, rec, use_07_metric)
    return {"ap": ap, "map": float(np.nanmean(ap)) if len(ap) else float("nan")}


def eval_detection_vid_motion(pred_boxlists, gt_boxlists, motion_ious, iou_thresh=0.5, use_07_metric=False,
                              device="cpu"):
    """vid_eval.py:39-51 with motion_specific=True: {"all" | "fast" | "medium" | "slow": {"ap", "map"}}."""
    return {name: eval_detection_vid(pred_boxlists, gt_boxlists, iou_thresh, use_07_metric, device, motion_ious, rng)
            for name, rng in MOTION_RANGES}


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


def do_vid_evaluation(dataset, predictions, output_folder=None, logger=None, device="cpu", motion_ious=None):
    """vid_eval.py:14-81 (box_only=False): predictions are resized to the original image size, AP50 per class + mAP +
    CorLoc are logged and written to `result.txt`.  `motion_ious` (list per image, or the path of the dataset's
    vid_groundtruth_motion_iou.mat) switches on the motion-specific report (`motion_specific=True`: one AP50 line each
    for all / fast / medium / slow).  `dataset` provides get_img_info(i) -> {"width","height"}, get_groundtruth(i) ->
    BoxList and map_class_id_to_class_name(i)."""
    preds, gts = [], []
    for i, p in enumerate(predictions):
        info = dataset.get_img_info(i)
        preds.append(p.resize((info["width"], info["height"])))
        gts.append(dataset.get_groundtruth(i))
    if isinstance(motion_ious, (str, bytes)):
        motion_ious = load_motion_ious(motion_ious)
    if motion_ious is not None:
        by_motion = eval_detection_vid_motion(preds, gts, motion_ious, 0.5, False, device)
        res = by_motion["all"]
    else:
        res = eval_detection_vid(preds, gts, 0.5, False, device)
        by_motion = {"all": res}
    corloc, corloc_avg = corloc_eval_detection_vid(preds, gts, 0.5, device)
    s = "".join("AP50 | motion={:>6s} = {:0.4f}\n".format(name, r["map"]) for name, r in by_motion.items())
    s += "Category AP:\n"
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
    return dict(res, corloc=corloc, corloc_avg=corloc_avg, text=s, motion={k: v["map"] for k, v in by_motion.items()})

#!/usr/bin/env python
"""Generate tests/golden/vid_eval_vectors.json by RUNNING the reference's own VID evaluator on this container's CPU.

Loads, by path and unmodified, /root/reference/mega_core/structures/bounding_box.py, boxlist_ops.py and
data/datasets/evaluation/vid/vid_eval.py.  The only stand-in is ``mega_core.layers.nms`` (imported by boxlist_ops.py for
boxlist_nms, never called by the evaluator).  For every seeded scenario the synthetic predictions / ground truth and
the reference's outputs - eval_detection_vid (per-class AP, mAP; motion_specific=False, the branch
mega_core/engine/inference.py takes for DiffusionVID) and corloc_eval_detection_vid - are stored.

Run:  python tests/golden/make_golden_eval.py     (needs /root/reference)
"""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("DVID_REFERENCE", "/root/reference")


def _load(name, rel):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, rel))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def scenario(seed, n_img, n_cls, max_gt, max_pred, size=(640, 360), quantise_scores=False, jitter=12.0):
    """Ground truth: random boxes; predictions: jittered copies of some gt boxes (true positives of varying IoU, some
    duplicates, some with a wrong label) plus random false positives; some images without gt or without predictions."""
    g = torch.Generator().manual_seed(seed)
    W, H = size
    imgs = []
    for i in range(n_img):
        n_gt = int(torch.randint(0, max_gt + 1, (1,), generator=g))
        xy = torch.rand(n_gt, 2, generator=g) * torch.tensor([W * 0.6, H * 0.6])
        wh = torch.rand(n_gt, 2, generator=g) * torch.tensor([W * 0.35, H * 0.35]) + 8
        gt = torch.cat([xy, xy + wh], 1).round()
        gl = torch.randint(1, n_cls + 1, (n_gt,), generator=g)
        preds, pl = [], []
        for k in range(n_gt):
            for _ in range(int(torch.randint(0, 3, (1,), generator=g))):
                preds.append(gt[k] + torch.randn(4, generator=g) * jitter)
                wrong = torch.rand(1, generator=g).item() < 0.15
                pl.append(int(torch.randint(1, n_cls + 1, (1,), generator=g)) if wrong else int(gl[k]))
        n_fp = int(torch.randint(0, max_pred + 1, (1,), generator=g))
        for _ in range(n_fp):
            a = torch.rand(2, generator=g) * torch.tensor([W * 0.7, H * 0.7])
            preds.append(torch.cat([a, a + torch.rand(2, generator=g) * 120 + 4]))
            pl.append(int(torch.randint(1, n_cls + 1, (1,), generator=g)))
        if i % 7 == 3:
            preds, pl = [], []
        pb = torch.stack(preds) if preds else torch.zeros(0, 4)
        sc = torch.rand(len(pl), generator=g)
        if quantise_scores:
            sc = (sc * 8).round() / 8          # many exact score ties
        imgs.append(dict(size=[W, H], gt_boxes=gt.tolist(), gt_labels=gl.tolist(), pred_boxes=pb.tolist(),
                         pred_labels=pl, pred_scores=sc.tolist()))
    return imgs


def main():
    for name in ("mega_core", "mega_core.structures", "mega_core.layers"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["mega_core.layers"].nms = None
    bb = _load("mega_core.structures.bounding_box", "mega_core/structures/bounding_box.py")
    _load("mega_core.structures.boxlist_ops", "mega_core/structures/boxlist_ops.py")
    ev = _load("ref_vid_eval", "mega_core/data/datasets/evaluation/vid/vid_eval.py")

    out = {"reference": "sdroh1027/DiffusionVID mega_core/data/datasets/evaluation/vid/vid_eval.py (unmodified, run on CPU)",
           "scenarios": []}
    specs = [dict(seed=1, n_img=40, n_cls=5, max_gt=4, max_pred=6),
             dict(seed=2, n_img=60, n_cls=30, max_gt=3, max_pred=10),
             dict(seed=3, n_img=25, n_cls=3, max_gt=6, max_pred=4, jitter=30.0),
             dict(seed=4, n_img=30, n_cls=4, max_gt=5, max_pred=5, jitter=2.0),
             dict(seed=5, n_img=12, n_cls=2, max_gt=1, max_pred=0)]
    for sp in specs:
        imgs = scenario(**sp)
        preds, gts = [], []
        for im in imgs:
            p = bb.BoxList(torch.tensor(im["pred_boxes"], dtype=torch.float32).reshape(-1, 4), tuple(im["size"]), "xyxy")
            p.add_field("labels", torch.tensor(im["pred_labels"], dtype=torch.int64))
            p.add_field("scores", torch.tensor(im["pred_scores"], dtype=torch.float32))
            t = bb.BoxList(torch.tensor(im["gt_boxes"], dtype=torch.float32).reshape(-1, 4), tuple(im["size"]), "xyxy")
            t.add_field("labels", torch.tensor(im["gt_labels"], dtype=torch.int64))
            preds.append(p); gts.append(t)
        with redirect_stdout(io.StringIO()):
            res = ev.eval_detection_vid(pred_boxlists=preds, gt_boxlists=gts, iou_thresh=0.5,
                                        motion_ranges=[[0.0, 1.0]], motion_specific=False, use_07_metric=False)
            res07 = ev.eval_detection_vid(pred_boxlists=preds, gt_boxlists=gts, iou_thresh=0.5,
                                          motion_ranges=[[0.0, 1.0]], motion_specific=False, use_07_metric=True)
            corloc, corloc_avg, _ = ev.corloc_eval_detection_vid(pred_boxlists=preds, gt_boxlists=gts, iou_thresh=0.5)
        ap = [None if np.isnan(a) else float(a) for a in res[0]["ap"]]
        ap07 = [None if np.isnan(a) else float(a) for a in res07[0]["ap"]]
        out["scenarios"].append(dict(spec=sp, images=imgs, ap=ap, map=float(res[0]["map"]), ap07=ap07,
                                     map07=float(res07[0]["map"]),
                                     corloc={str(int(k)): float(v) for k, v in corloc.items()},
                                     corloc_avg=float(corloc_avg)))
        print("scenario seed=%d: mAP %.6f  mAP07 %.6f  CorLoc %.6f" % (sp["seed"], res[0]["map"], res07[0]["map"], corloc_avg))
    with open(os.path.join(HERE, "vid_eval_vectors.json"), "w") as f:
        json.dump(out, f)
    print("wrote", os.path.join(HERE, "vid_eval_vectors.json"), os.path.getsize(os.path.join(HERE, "vid_eval_vectors.json")), "bytes")


if __name__ == "__main__":
    main()

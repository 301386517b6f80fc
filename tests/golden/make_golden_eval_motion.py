#!/usr/bin/env python
"""Generate tests/golden/vid_eval_motion_vectors.json by RUNNING the reference's own motion-specific VID evaluator
(mega_core/data/datasets/evaluation/vid/vid_eval.py::calc_detection_vid_prec_rec + calc_detection_vid_ap, unmodified,
loaded by path) on this container's CPU, for the four motion ranges of do_vid_evaluation (vid_eval.py:39-41).

`eval_detection_vid(motion_specific=True)` itself reads the dataset's vid_groundtruth_motion_iou.mat from a hard-coded
relative path (:142-147); the synthetic scenarios here carry their own per-box motion IoUs, so the two functions it calls
are driven directly with the same arguments it passes (:151-159).  Scenarios come from make_golden_eval.scenario.

Run:  python tests/golden/make_golden_eval_motion.py     (needs /root/reference)
"""
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden_eval import _load, scenario  # noqa: E402

RANGES = [("all", [0.0, 1.0]), ("fast", [0.0, 0.7]), ("medium", [0.7, 0.9]), ("slow", [0.9, 1.0])]


def main():
    for name in ("mega_core", "mega_core.structures", "mega_core.layers"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["mega_core.layers"].nms = None
    bb = _load("mega_core.structures.bounding_box", "mega_core/structures/bounding_box.py")
    _load("mega_core.structures.boxlist_ops", "mega_core/structures/boxlist_ops.py")
    ev = _load("ref_vid_eval", "mega_core/data/datasets/evaluation/vid/vid_eval.py")
    out = {"reference": "sdroh1027/DiffusionVID vid_eval.py calc_detection_vid_prec_rec/_ap (unmodified, CPU), "
                        "motion ranges of do_vid_evaluation", "scenarios": []}
    specs = [dict(seed=11, n_img=40, n_cls=5, max_gt=4, max_pred=6),
             dict(seed=12, n_img=50, n_cls=30, max_gt=3, max_pred=8),
             dict(seed=13, n_img=30, n_cls=3, max_gt=6, max_pred=4, jitter=30.0),
             dict(seed=14, n_img=20, n_cls=2, max_gt=5, max_pred=3, jitter=4.0)]
    for sp in specs:
        imgs = scenario(**sp)
        g = torch.Generator().manual_seed(1000 + sp["seed"])
        preds, gts, motion = [], [], []
        for im in imgs:
            p = bb.BoxList(torch.tensor(im["pred_boxes"], dtype=torch.float32).reshape(-1, 4), tuple(im["size"]), "xyxy")
            p.add_field("labels", torch.tensor(im["pred_labels"], dtype=torch.int64))
            p.add_field("scores", torch.tensor(im["pred_scores"], dtype=torch.float32))
            t = bb.BoxList(torch.tensor(im["gt_boxes"], dtype=torch.float32).reshape(-1, 4), tuple(im["size"]), "xyxy")
            t.add_field("labels", torch.tensor(im["gt_labels"], dtype=torch.int64))
            preds.append(p); gts.append(t)
            # motion IoUs spread over the three bands, a few exactly on the band edges
            m = torch.rand(len(im["gt_labels"]), generator=g)
            edge = torch.rand(len(im["gt_labels"]), generator=g)
            m = torch.where(edge < 0.1, torch.full_like(m, 0.7), torch.where(edge > 0.9, torch.full_like(m, 0.9), m))
            im["motion_iou"] = [float(v) for v in m.double().tolist()]
            motion.append(im["motion_iou"])
        res = {}
        for name, rng in RANGES:
            with redirect_stdout(io.StringIO()):
                prec, rec = ev.calc_detection_vid_prec_rec(pred_boxlists=preds, gt_boxlists=gts, motion_ious=motion,
                                                           iou_thresh=0.5, motion_range=rng)
                ap = ev.calc_detection_vid_ap(prec, rec, use_07_metric=False)
            res[name] = dict(ap=[None if np.isnan(a) else float(a) for a in ap], map=float(np.nanmean(ap)))
        out["scenarios"].append(dict(spec=sp, images=imgs, motion=res))
        print("scenario seed=%d: " % sp["seed"] + "  ".join("%s %.6f" % (k, v["map"]) for k, v in res.items()))
    path = os.path.join(HERE, "vid_eval_motion_vectors.json")
    with open(path, "w") as f:
        json.dump(out, f)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()

"""VID AP50 / CorLoc metrics (diffusionvid_b200/evaluation.py) against golden vectors produced by the
reference's own evaluator (tests/golden/make_golden_eval.py runs mega_core/data/datasets/evaluation/vid/vid_eval.py
unmodified on CPU): per-class AP, mAP (VOC area and VOC07 11-point) and CorLoc must agree to 1e-12 on five seeded
scenarios (5..30 classes, images without ground truth / without detections, duplicate and mislabelled detections,
loose and tight localisation)."""
import json
import math
import os

import numpy as np
import pytest
import torch

from diffusionvid_b200 import evaluation as ev
from diffusionvid_b200.structures import BoxList

HERE = os.path.dirname(os.path.abspath(__file__))


def _scenarios():
    with open(os.path.join(HERE, "golden", "vid_eval_vectors.json")) as f:
        return json.load(f)["scenarios"]


def _boxlists(images):
    preds, gts = [], []
    for im in images:
        p = BoxList(torch.tensor(im["pred_boxes"], dtype=torch.float32).reshape(-1, 4), tuple(im["size"]), "xyxy")
        p.add_field("labels", torch.tensor(im["pred_labels"], dtype=torch.int64))
        p.add_field("scores", torch.tensor(im["pred_scores"], dtype=torch.float32))
        t = BoxList(torch.tensor(im["gt_boxes"], dtype=torch.float32).reshape(-1, 4), tuple(im["size"]), "xyxy")
        t.add_field("labels", torch.tensor(im["gt_labels"], dtype=torch.int64))
        preds.append(p); gts.append(t)
    return preds, gts


@pytest.mark.parametrize("idx", range(5))
def test_ap_and_corloc_match_the_reference_evaluator(idx):
    sc = _scenarios()[idx]
    preds, gts = _boxlists(sc["images"])
    for key, use07 in (("ap", False), ("ap07", True)):
        res = ev.eval_detection_vid(preds, gts, 0.5, use_07_metric=use07)
        want = sc[key]
        assert len(res["ap"]) == len(want)
        for a, w in zip(res["ap"], want):
            if w is None:
                assert math.isnan(a)
            else:
                assert abs(a - w) <= 1e-12, (key, a, w)
        assert abs(res["map"] - sc["map07" if use07 else "map"]) <= 1e-12
    corloc, avg = ev.corloc_eval_detection_vid(preds, gts, 0.5)
    assert {str(k) for k in corloc} == set(sc["corloc"])
    for k, v in corloc.items():
        assert abs(v - sc["corloc"][str(k)]) <= 1e-12
    assert abs(avg - sc["corloc_avg"]) <= 1e-12


def test_known_answers_and_edge_cases():
    """hand-checkable cases: a perfect detector has AP 1; a detection on an image without ground truth is a false
    positive that lowers precision only after it outranks a true positive; empty inputs; length mismatch."""
    size = (100, 100)

    def bl(boxes, labels, scores=None):
        b = BoxList(torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4), size, "xyxy")
        b.add_field("labels", torch.tensor(labels, dtype=torch.int64))
        if scores is not None:
            b.add_field("scores", torch.tensor(scores, dtype=torch.float32))
        return b
    gts = [bl([[10, 10, 50, 50]], [1]), bl([[20, 20, 60, 60], [5, 5, 15, 15]], [1, 2])]
    perfect = [bl([[10, 10, 50, 50]], [1], [0.9]), bl([[20, 20, 60, 60], [5, 5, 15, 15]], [1, 2], [0.8, 0.7])]
    res = ev.eval_detection_vid(perfect, gts)
    assert np.isnan(res["ap"][0]) and res["ap"][1] == pytest.approx(1.0) and res["ap"][2] == pytest.approx(1.0)
    assert res["map"] == pytest.approx(1.0)
    # a higher-scoring false positive on a third image without ground truth: class-1 precision 0, 1/2, 2/3 at recall
    # 0, 1/2, 1 -> the monotone envelope is 2/3 everywhere -> AP 2/3
    gts3 = gts + [bl([], [])]
    preds3 = perfect + [bl([[0, 0, 30, 30]], [1], [0.95])]
    res3 = ev.eval_detection_vid(preds3, gts3)
    assert res3["ap"][1] == pytest.approx(2.0 / 3.0, abs=1e-9)
    # duplicate detection of the same box: the second is a false positive
    dup = [bl([[10, 10, 50, 50], [11, 11, 50, 50]], [1, 1], [0.9, 0.8]), bl([], [], [])]
    r = ev.eval_detection_vid(dup, gts)
    assert r["ap"][1] == pytest.approx(0.5)              # recall reaches 1/2 at precision 1
    corloc, avg = ev.corloc_eval_detection_vid(perfect, gts)
    assert corloc[1] == pytest.approx(1.0) and corloc[2] == pytest.approx(0.0)    # image 2's top box is class 1
    assert avg == pytest.approx(0.5)
    empty = ev.eval_detection_vid([], [])
    assert len(empty["ap"]) == 0 and math.isnan(empty["map"])
    with pytest.raises(ValueError):
        ev.eval_detection_vid(perfect, gts[:1])


def test_do_vid_evaluation_writes_result_file(tmp_path):
    sc = _scenarios()[0]
    preds, gts = _boxlists(sc["images"])

    class DS:
        def get_img_info(self, i):
            return {"width": sc["images"][i]["size"][0], "height": sc["images"][i]["size"][1]}

        def get_groundtruth(self, i):
            return gts[i]

        def map_class_id_to_class_name(self, i):
            return "class%d" % i
    out = ev.do_vid_evaluation(DS(), preds, str(tmp_path))
    assert abs(out["map"] - sc["map"]) <= 1e-12
    text = open(os.path.join(str(tmp_path), "result.txt")).read()
    assert text.startswith("AP50 | motion=   all = %.4f" % sc["map"]) and "Mean CorLoc: %.4f" % sc["corloc_avg"] in text


def _motion_scenarios():
    with open(os.path.join(HERE, "golden", "vid_eval_motion_vectors.json")) as f:
        return json.load(f)["scenarios"]


@pytest.mark.parametrize("idx", range(4))
def test_motion_specific_ap_matches_the_reference_evaluator(idx):
    """vid_eval.py:39-44 (motion_specific=True): all / fast / medium / slow AP50 with ignored ground truth, golden
    vectors from the reference's own calc_detection_vid_prec_rec (tests/golden/make_golden_eval_motion.py)."""
    sc = _motion_scenarios()[idx]
    preds, gts = _boxlists(sc["images"])
    motion = [im["motion_iou"] for im in sc["images"]]
    res = ev.eval_detection_vid_motion(preds, gts, motion)
    assert list(res) == ["all", "fast", "medium", "slow"]
    for name, want in sc["motion"].items():
        got = res[name]
        assert len(got["ap"]) == len(want["ap"])
        for a, w in zip(got["ap"], want["ap"]):
            if w is None:
                assert math.isnan(a)
            else:
                assert abs(a - w) <= 1e-12, (name, a, w)
        assert abs(got["map"] - want["map"]) <= 1e-12


def test_motion_report_lines(tmp_path):
    sc = _motion_scenarios()[0]
    preds, gts = _boxlists(sc["images"])

    class DS:
        def get_img_info(self, i):
            return {"width": sc["images"][i]["size"][0], "height": sc["images"][i]["size"][1]}

        def get_groundtruth(self, i):
            return gts[i]

        def map_class_id_to_class_name(self, i):
            return "class%d" % i
    out = ev.do_vid_evaluation(DS(), preds, str(tmp_path), motion_ious=[im["motion_iou"] for im in sc["images"]])
    lines = out["text"].splitlines()
    for i, name in enumerate(("all", "fast", "medium", "slow")):
        assert lines[i] == "AP50 | motion={:>6s} = {:0.4f}".format(name, sc["motion"][name]["map"])

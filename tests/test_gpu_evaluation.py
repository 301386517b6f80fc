"""The VID evaluator's matching on the GPU (csrc/vid_match.cu behind dvid_vid_match): one launch over
all images must reproduce the CPU evaluator's per-detection records EXACTLY - and with them the golden AP numbers the
reference's own vid_eval.py produced (tests/golden/vid_eval_vectors.json, vid_eval_motion_vectors.json)."""
import math

import numpy as np
import pytest
import torch

from diffusionvid_b200 import evaluation as ev, ops
from diffusionvid_b200.structures import BoxList
from tests.test_evaluation import _boxlists, _motion_scenarios, _scenarios

pytestmark = pytest.mark.gpu


def _same_records(a, b):
    (sa, la, ha, ia), na = a
    (sb, lb, hb, ib), nb = b
    assert na == nb
    assert torch.equal(sa, sb) and torch.equal(la, lb) and torch.equal(ha, hb)
    assert ia.dtype == ib.dtype == torch.float64 and torch.equal(ia, ib)


@pytest.mark.parametrize("idx", range(5))
def test_gpu_matching_reproduces_the_golden_ap(cuda, idx):
    sc = _scenarios()[idx]
    preds, gts = _boxlists(sc["images"])
    l0 = ops.LAUNCHES
    _same_records(ev.match_detections(preds, gts, 0.5, device=cuda), ev.match_detections(preds, gts, 0.5))
    assert ops.LAUNCHES == l0 + 1                              # all images, one launch
    res = ev.eval_detection_vid(preds, gts, 0.5, device=cuda)
    for a, w in zip(res["ap"], sc["ap"]):
        assert (math.isnan(a) if w is None else abs(a - w) <= 1e-12), (a, w)
    assert abs(res["map"] - sc["map"]) <= 1e-12


@pytest.mark.parametrize("idx", range(4))
def test_gpu_matching_motion_specific(cuda, idx):
    sc = _motion_scenarios()[idx]
    preds, gts = _boxlists(sc["images"])
    motion = [im["motion_iou"] for im in sc["images"]]
    for _, rng in ev.MOTION_RANGES:
        _same_records(ev.match_detections(preds, gts, 0.5, cuda, motion, rng),
                      ev.match_detections(preds, gts, 0.5, "cpu", motion, rng))
    res = ev.eval_detection_vid_motion(preds, gts, motion, device=cuda)
    for name, want in sc["motion"].items():
        for a, w in zip(res[name]["ap"], want["ap"]):
            assert (math.isnan(a) if w is None else abs(a - w) <= 1e-12), (name, a, w)
        assert abs(res[name]["map"] - want["map"]) <= 1e-12


def _random_set(seed, n_img, max_gt, max_pred, n_cls):
    """Integer boxes on a coarse grid: many exact IoU ties and duplicates, equal scores, empty images, more ground truth
    than one warp pass (> 32) in some images."""
    g = torch.Generator().manual_seed(seed)
    preds, gts, motion = [], [], []
    for i in range(n_img):
        ng = int(torch.randint(0, max_gt + 1, (1,), generator=g))
        npred = int(torch.randint(0, max_pred + 1, (1,), generator=g))

        def boxes(n):
            xy = torch.randint(0, 6, (n, 2), generator=g).float() * 8
            wh = torch.randint(1, 4, (n, 2), generator=g).float() * 8
            return torch.cat([xy, xy + wh], 1)
        gb = boxes(ng)
        if ng > 2:
            gb[1] = gb[0]                                       # duplicate ground truth: exact ties
        t = BoxList(gb, (64, 64), "xyxy")
        t.add_field("labels", torch.randint(1, n_cls + 1, (ng,), generator=g))
        pb = boxes(npred)
        if ng and npred:
            take = torch.randint(0, ng, (npred,), generator=g)
            pb = torch.where(torch.rand(npred, 1, generator=g) < 0.6, gb[take], pb)
        p = BoxList(pb, (64, 64), "xyxy")
        p.add_field("labels", torch.randint(1, n_cls + 1, (npred,), generator=g))
        p.add_field("scores", torch.randint(0, 8, (npred,), generator=g).float() / 8)      # equal scores: stable order
        preds.append(p); gts.append(t)
        motion.append([] if i % 7 == 3 else (torch.randint(0, 11, (ng,), generator=g).float() / 10).tolist())
    return preds, gts, motion


@pytest.mark.parametrize("seed,max_gt", [(0, 6), (1, 80), (2, 3)])
def test_gpu_matching_equals_the_cpu_loop_on_tie_heavy_random_sets(cuda, seed, max_gt):
    preds, gts, motion = _random_set(seed, 60, max_gt, 40, 3)
    _same_records(ev.match_detections(preds, gts, 0.5, cuda), ev.match_detections(preds, gts, 0.5))
    for rng in ((0.0, 1.0), (0.0, 0.7), (0.7, 0.9), (0.9, 1.0)):
        a = ev.match_detections(preds, gts, 0.5, cuda, motion, rng)
        b = ev.match_detections(preds, gts, 0.5, "cpu", motion, rng)
        _same_records(a, b)
    hits = ev.match_detections(preds, gts, 0.5, cuda)[0][2]
    assert 0 < int(hits.sum()) < hits.numel()                    # the set exercises both outcomes


def test_gpu_matching_empty_inputs(cuda):
    (s, l, h, i), n_pos = ev.match_detections([], [], 0.5, cuda)
    assert s.numel() == 0 and h.numel() == 0 and n_pos == {}
    e = BoxList(torch.zeros(0, 4), (10, 10), "xyxy")
    e.add_field("labels", torch.zeros(0, dtype=torch.int64))
    e.add_field("scores", torch.zeros(0))
    _same_records(ev.match_detections([e], [e], 0.5, cuda), ev.match_detections([e], [e], 0.5))
    with pytest.raises(ValueError):
        ev.match_detections([e], [], 0.5, cuda)


def test_inference_driver_evaluates_on_the_gpu(cuda, tmp_path):
    """engine.inference(evaluate=True) on a CUDA device: the report equals the golden mAP and the matching ran as
    kernel launches (one per eval_detection_vid call), not as the CPU loop."""
    from diffusionvid_b200 import engine
    from diffusionvid_b200.structures import ImageList
    sc = _scenarios()[1]
    preds, gts = _boxlists(sc["images"])

    class DS:
        def get_img_info(self, i):
            return {"width": sc["images"][i]["size"][0], "height": sc["images"][i]["size"][1]}

        def get_groundtruth(self, i):
            return gts[i]

        def map_class_id_to_class_name(self, i):
            return "class%d" % i

    class Loader:
        dataset = DS()

        def __iter__(self):
            for i in range(len(preds)):
                img = ImageList(torch.zeros(1, 3, 8, 8), [(8, 8)])
                yield dict(cur=img, ref_l=[], ref_g=[], frame_id=i), None, [[i]]

    class Model(torch.nn.Module):
        def forward(self, images):
            assert images["cur"].tensors.is_cuda
            return [preds[images["frame_id"]].to(cuda)]
    l0 = ops.LAUNCHES
    out, metrics = engine.inference(Model(), Loader(), device=cuda, output_folder=str(tmp_path), evaluate=True)
    assert len(out) == len(preds) and ops.LAUNCHES == l0 + 1
    assert abs(metrics["map"] - sc["map"]) <= 1e-12
    assert abs(metrics["corloc_avg"] - sc["corloc_avg"]) <= 1e-12

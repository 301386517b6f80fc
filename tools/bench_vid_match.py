"""Evaluator matching, GPU kernel vs the CPU loop (diffusionvid_b200/evaluation.py): synthetic set shaped like ImageNet-VID
val output (100 detections and 1..4 ground-truth boxes per frame).  The CPU loop runs on a sample of the frames."""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import evaluation as ev  # noqa: E402
from diffusionvid_b200.structures import BoxList  # noqa: E402


def make(n_img, seed=0):
    g = torch.Generator().manual_seed(seed)
    preds, gts = [], []
    for _ in range(n_img):
        ng = int(torch.randint(1, 5, (1,), generator=g))
        xy = torch.rand(ng, 2, generator=g) * 400
        gb = torch.cat([xy, xy + 40 + torch.rand(ng, 2, generator=g) * 300], 1).round()
        t = BoxList(gb, (1000, 600), "xyxy")
        t.add_field("labels", torch.randint(1, 31, (ng,), generator=g))
        src = torch.randint(0, ng, (100,), generator=g)
        pb = (gb[src] + torch.randn(100, 4, generator=g) * 25).round()
        p = BoxList(pb, (1000, 600), "xyxy")
        p.add_field("labels", torch.where(torch.rand(100, generator=g) < 0.7, t.get_field("labels")[src],
                                          torch.randint(1, 31, (100,), generator=g)))
        p.add_field("scores", torch.rand(100, generator=g))
        preds.append(p); gts.append(t)
    return preds, gts


def main():
    n_img, n_cpu = 20000, 1000
    preds, gts = make(n_img)
    ev.match_detections(preds[:100], gts[:100], 0.5, "cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    rec_g = ev.match_detections(preds, gts, 0.5, "cuda")
    torch.cuda.synchronize()
    t_gpu = time.perf_counter() - t0
    t0 = time.perf_counter()
    rec_c = ev.match_detections(preds[:n_cpu], gts[:n_cpu], 0.5, "cpu")
    t_cpu = time.perf_counter() - t0
    same = torch.equal(rec_g[0][2][:n_cpu * 100], rec_c[0][2])
    print("VID_MATCH " + json.dumps({"images": n_img, "detections": n_img * 100,
                                     "gpu_s_total_with_packing": round(t_gpu, 3),
                                     "gpu_images_per_s": round(n_img / t_gpu, 1),
                                     "cpu_loop_images": n_cpu, "cpu_loop_s": round(t_cpu, 3),
                                     "cpu_images_per_s": round(n_cpu / t_cpu, 1),
                                     "hits_equal_on_the_cpu_sample": bool(same)}))


if __name__ == "__main__":
    main()

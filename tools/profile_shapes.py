#!/usr/bin/env python
"""Per-shape time breakdown of one clip (eager, single stream, CUDA events around every conv/GEMM launch)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=64); ap.add_argument("--global-frames", type=int, default=24)
ap.add_argument("--T", type=int, default=4); ap.add_argument("--proposals", type=int, default=300)
ap.add_argument("--height", type=int, default=600); ap.add_argument("--width", type=int, default=1000)
ap.add_argument("--backbone", default="r101")
a = ap.parse_args()
from diffusionvid_b200 import model as pm, ops, synth
dev = torch.device("cuda", 0)
hp = dict(bench.HP_BASE, num_proposals=a.proposals, sample_step=a.T, device=str(dev))
if a.backbone == "swinb":
    hp.update(swin=dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)), infer_batch=4, all_frame_interval=4)
m = pm.DiffusionDet(hp); m.load_state_dict(synth.make_state_dict(seed=1234, blocks=hp["blocks"], swin=hp.get("swin")), strict=False); m.to(dev)
m.use_graphs = m.use_streams = False
samples, _ = bench.make_clip_inputs(a, dev, pinned=False)
with torch.no_grad():
    bench.run_clip(m, samples, False)
    ops.PROFILE = {}
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); bench.run_clip(m, samples, False); e1.record(); torch.cuda.synchronize()
prof = ops.PROFILE; ops.PROFILE = None
rows = []
for k, (evs, fl, by) in prof.items():
    t = sum(s.elapsed_time(e) for s, e in evs)
    rows.append((t, k, len(evs), fl, by))
rows.sort(reverse=True)
print("clip wall (eager, instrumented): %.1f ms" % e0.elapsed_time(e1))
print("%-64s %6s %9s %8s %9s %8s" % ("family:shape", "n", "total_ms", "avg_us", "TFLOP/s", "GB/s"))
for t, k, n, fl, by in rows[:60]:
    print("%-64s %6d %9.2f %8.1f %9.1f %8.0f" % (k[:64], n, t, 1e3 * t / n, fl / t / 1e9 if t else 0, by / t / 1e6 if t else 0))

#!/usr/bin/env python
"""Where does one bench step go?  Wall time of the clip vs GPU time of the captured units (CUDA events around every
`_run_unit`) vs host time spent in the calls (perf_counter).  GPU time not covered by units = eager glue kernels,
host time during which the GPU idles = launch/sync overhead.

  python tools/step_timeline.py [--backbone swinb] [--frames 64 --global-frames 24]
"""
import argparse
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--global-frames", type=int, default=24)
    ap.add_argument("--T", type=int, default=4)
    ap.add_argument("--proposals", type=int, default=300)
    ap.add_argument("--height", type=int, default=600)
    ap.add_argument("--width", type=int, default=1000)
    ap.add_argument("--backbone", default="r101", choices=["r101", "swinb"])
    ap.add_argument("--host", action="store_true", help="host inputs / host results (the e2e path)")
    a = ap.parse_args()
    from diffusionvid_b200 import model as pm, synth
    dev = torch.device("cuda", 0)
    hp = dict(bench.HP_BASE, num_proposals=a.proposals, sample_step=a.T, device=str(dev))
    if a.backbone == "swinb":
        hp.update(swin=dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)), infer_batch=4,
                  all_frame_interval=4)
    m = pm.DiffusionDet(hp)
    m.load_state_dict(synth.make_state_dict(seed=1234, blocks=hp["blocks"], swin=hp.get("swin")), strict=False)
    m.to(dev)
    m.host_results = a.host
    samples, _ = bench.make_clip_inputs(a, dev, pinned=a.host)
    units = []
    orig = m._run_unit

    def timed_unit(name, fn, key, tensors, consts):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        r = orig(name, fn, key, tensors, consts)
        t1 = time.perf_counter()
        e1.record()
        units.append((name, e0, e1, t1 - t0))
        return r

    calls = []
    fwd = m._forward_test

    def timed_forward(imgs, ref_l, ref_g, infos):
        t0 = time.perf_counter()
        r = fwd(imgs, ref_l, ref_g, infos)
        calls.append((infos["frame_id"], time.perf_counter() - t0))
        return r

    with torch.no_grad():
        for _ in range(3):
            bench.run_clip(m, samples, a.host)
        torch.cuda.synchronize()
        m._run_unit = timed_unit
        m._forward_test = timed_forward
        t0 = time.perf_counter()
        n, _ = bench.run_clip(m, samples, a.host)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
    by = {}
    for name, e0, e1, th in units:
        d = by.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += e0.elapsed_time(e1)
        d[2] += th * 1e3
    print("clip: %d frames, wall %.2f ms (%.1f frames/s)" % (n, wall, n / wall * 1e3))
    gpu = 0.0
    for name, (c, g, th) in by.items():
        print("  unit %-8s n=%3d  gpu %.2f ms  host-side launch %.2f ms" % (name, c, g, th))
        gpu += g
    key = [t for f, t in calls if t > 2e-4]
    noop = [t for f, t in calls if t <= 2e-4]
    print("  GPU time inside units %.2f ms = %.1f%% of wall; outside %.2f ms" % (gpu, 100 * gpu / wall, wall - gpu))
    print("  host: %d key calls %.2f ms total, %d queue-only calls %.2f ms total, between calls %.2f ms"
          % (len(key), sum(key) * 1e3, len(noop), sum(noop) * 1e3, wall - sum(t for _, t in calls) * 1e3))
    print("  key calls (ms):", " ".join("%.2f" % (t * 1e3) for t in key))


if __name__ == "__main__":
    main()

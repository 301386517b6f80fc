#!/usr/bin/env python
"""Profiling driver: one warm clip, then one clip between cudaProfilerStart/Stop (run under
`ncu --profile-from-start off ...`).  Also usable stand-alone: prints a CUDA-event timing of the profiled clip.

  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
      --log-file gpurun_out/launches.csv python tools/profile_clip.py --frames 16 --global-frames 8
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=16)
    ap.add_argument("--global-frames", type=int, default=8)
    ap.add_argument("--T", type=int, default=4)
    ap.add_argument("--proposals", type=int, default=300)
    ap.add_argument("--height", type=int, default=600)
    ap.add_argument("--width", type=int, default=1000)
    ap.add_argument("--warm", type=int, default=1)
    ap.add_argument("--blocks", default="3,4,23,3")
    ap.add_argument("--no-graphs", action="store_true")
    ap.add_argument("--no-streams", action="store_true")
    ap.add_argument("--frames-per-stream", type=int, default=2)
    ap.add_argument("--backbone", default="r101", choices=["r101", "swinb"])
    a = ap.parse_args()
    from diffusionvid_b200 import model as pm, synth
    dev = torch.device("cuda", 0)
    blocks = tuple(int(x) for x in a.blocks.split(","))
    hp = dict(bench.HP_BASE, num_proposals=a.proposals, sample_step=a.T, device=str(dev), blocks=blocks)
    if a.backbone == "swinb":
        hp.update(swin=dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)), infer_batch=4, all_frame_interval=4)
    m = pm.DiffusionDet(hp)
    m.load_state_dict(synth.make_state_dict(seed=1234, blocks=blocks, swin=hp.get("swin")), strict=False)
    m.to(dev)
    m.use_graphs = not a.no_graphs
    m.use_streams = not a.no_streams
    m.frames_per_stream = a.frames_per_stream
    samples, _ = bench.make_clip_inputs(a, dev, pinned=False)
    with torch.no_grad():
        for _ in range(a.warm + (0 if a.no_graphs else 1)):
            bench.run_clip(m, samples, False)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.cudart().cudaProfilerStart()
        e0.record()
        n, _ = bench.run_clip(m, samples, False)
        e1.record()
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()
    print("profiled clip: %d frames in %.2f ms" % (n, e0.elapsed_time(e1)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py - DiffusionVID inference hot path on B200: frames/sec through DiffusionDet.forward.

Workload (BASELINE.json metric / configs[2]): vid_R_101_DiffusionVID.yaml, N=300 boxes, T=4 DDIM steps, R-101 + FPN,
synthetic 1000x600 clips (padded 608x1024).  One *step* = one whole clip pushed frame by frame through the public
API, exactly as mega_core/engine/inference.py:26-93 drives the reference: L=64 frames (8 key batches) + 24 global
reference frames at the video start (backbone + 3 base stages on all 88 frames, farthest-point-sampled memory,
8 x [T x (3 base heads + global attention + conditioned head) + DDIM update + top-k] + NMS).

  value : frames/s with the clip's images already resident in HBM when the timed region starts.
  e2e   : the same clip fed from pinned HOST memory (H2D copies of every ref frame inside the timed region) and the
          detections read back to the host (D2H) - the number to compare with the reference arm.  The loop is the
          reference's (engine/inference.py:66-78): call, move the outputs to the CPU, store; the model's host-result
          BoxLists are deferred (they wait for their asynchronous D2H copy on first access), so the host runs ahead and
          the next batch's frame uploads overlap the current batch's compute; every detection is read on the host
          before the timed region ends.
  e2e_u8: the same call fed with the frames as decoded 8-bit images (uint8 ImageLists; SURVEY.md 8f-1 clip loader):
          the reference's ToTensor is evaluated by the first GPU kernel, a quarter of the bytes cross PCIe.
  roofline : the dominant kernel family (tcgen05 conv/GEMM), algorithmic FLOPs / CUDA-event time measured live in a
          separate instrumented pass, against MEASURED_PEAKS.json (bf16 dense, sustained).
  cpu_baseline / --impl reference : the fp32 CPU oracle (oracle/model.py, a PyTorch restatement of the reference; the
          reference itself cannot be imported - SURVEY.md 8c) on the host cores, on a bounded sample of the workload.

N>1 (torchrun): video-level sharding as in the reference (VIDTestDistributedSampler): every rank processes its own
clips, no data-path collective; value = total frames of all ranks / max-over-ranks time ("weak" scaling).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--T", type=int, default=4, help="DDIM sampling steps (MODEL.DiffusionDet.SAMPLE_STEP)")
    ap.add_argument("--frames", type=int, default=64, help="frames per clip")
    ap.add_argument("--global-frames", type=int, default=24)
    ap.add_argument("--proposals", type=int, default=300)
    ap.add_argument("--height", type=int, default=600)
    ap.add_argument("--width", type=int, default=1000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-roofline", action="store_true")
    ap.add_argument("--cpu-sample-frames", type=int, default=2)
    ap.add_argument("--backbone", default="r101", choices=["r101", "swinb"],
                    help="r101: vid_R_101_DiffusionVID.yaml (headline); swinb: vid_Swin_B_DiffusionVID.yaml "
                         "(INFER_BATCH 4, ALL_FRAME_INTERVAL 4)")
    ap.add_argument("--frames-per-stream", type=int, default=0,
                    help="frames of a batch per parallel stream branch inside the captured units (0 = model default)")
    ap.add_argument("--shard", default="videos", choices=["videos", "frames"],
                    help="N>1: 'videos' = every rank runs its own clips (reference scheme, weak scaling); 'frames' = "
                         "the frames of ONE clip are dealt to the ranks, one all-gather of memory candidates per "
                         "video + one result all-reduce per key batch (strong scaling)")
    return ap.parse_args()


HP_BASE = dict(num_classes=30, hidden=256, nheads=8, dim_dynamic=64, dim_ff=2048, num_heads=3, num_heads_local=1,
               num_cls=1, num_reg=3, snr_scale=2.0, use_nms=True, infer_batch=8, all_frame_interval=8,
               key_frame_location=0, global_enable=True, mem_size=900, mem_size2=150, topk=(75, 25),
               pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375), blocks=(3, 4, 23, 3))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1400.0), d.get("hbm_gbs", 6650.0), "measured"
    return 1590.0, 6650.0, "fallback"


# ---------------------------------------------------------------------------------------------------- CPU oracle arm
def cpu_oracle_run(args, frames_n, global_n, repeats):
    """Times the fp32 oracle on the host cores over a bounded sample clip; returns (frames/s, description)."""
    from diffusionvid_b200 import synth
    from oracle import model as om
    torch.set_num_threads(os.cpu_count() or 1)
    sd = synth.make_state_dict(seed=1234, blocks=HP_BASE["blocks"])
    ocfg = dict(num_proposals=args.proposals, sample_step=args.T)
    frames = synth.make_clip(frames_n, args.height, args.width, seed=1234)
    gidx = [min(frames_n - 1, (i * 7) % frames_n) for i in range(global_n)]
    samples = synth.clip_samples(frames, gidx, args.height, args.width)
    times = []
    for r in range(repeats):
        o = om.OracleDiffusionVID(sd, ocfg, fp16=False, noise=om.NoiseSource(1234 + r, args.proposals))
        t0 = time.perf_counter()
        n_out = 0
        with torch.no_grad():
            for s in samples:
                n_out += len(o.forward(s))
        times.append(time.perf_counter() - t0)
        assert n_out == frames_n
    best = min(times)
    desc = ("%d-frame clip + %d global frames, N=%d, T=%d, %dx%d, R-101+FPN, fp32 PyTorch oracle, %d threads"
            % (frames_n, global_n, args.proposals, args.T, args.width, args.height, torch.get_num_threads()))
    return frames_n / best, desc, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = max(1, args.steps)
    t0 = time.perf_counter()
    _ = cpu_oracle_run(args, args.cpu_sample_frames, 1, max(0, min(args.warmup, 1)))[0] if args.warmup > 0 else None
    fps, desc, times = cpu_oracle_run(args, args.cpu_sample_frames, 1, k)
    ms = 1000.0 * statistics.mean(times)
    line = {"impl": "reference", "metric": "frames/sec", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
            "steps": k, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "vid_R_101_DiffusionVID N=%d T=%d %dx%d (CPU sample: %s)"
                       % (args.proposals, args.T, args.width, args.height, desc)},
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": desc},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "wall_s": time.perf_counter() - t0}
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------------- GPU arm
class ClockSampler:
    def __init__(self, index):
        self.path = tempfile.mktemp(suffix=".csv")
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(index),
                 "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
                 "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
                 "clocks_event_reasons.sw_power_cap", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for ln in open(self.path):
                f = [x.strip() for x in ln.split(",")]
                if len(f) < 7:
                    continue
                try:
                    sm.append(float(f[0])); mx.append(float(f[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            busy = [x for x in sm if x > 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


def make_clip_inputs(args, dev, pinned, u8=False):
    """The per-frame `images` dicts of one clip.  pinned=True: host (pinned) tensors; else device-resident.
    u8=True: the frames as decoded 8-bit images (clip-loader mode, SURVEY.md 8f-1: the reference's ToTensor runs inside
    the first kernel instead of on the host)."""
    from diffusionvid_b200 import structures, synth
    L = args.frames
    frames = synth.make_clip(L, args.height, args.width, seed=1234)
    if u8:
        frames = (frames * 255.0).round().clamp(0, 255).to(torch.uint8)
    gidx = [(i * 7 + 3) % L for i in range(args.global_frames)]
    if pinned:
        store = frames.pin_memory()
    else:
        store = frames.to(dev)
    size = [(args.height, args.width)]
    samples = []
    for f in range(L):
        mo = 3 if getattr(args, "backbone", "r101") == "swinb" else 7      # MODEL.VID.MEGA.MAX_OFFSET
        if f == 0:
            ref_l = list(range(0, min(mo, L - 1) + 1)); ref_g = gidx
        else:
            ref_l = [min(f + mo, L - 1)]; ref_g = []
        samples.append(dict(cur=structures.ImageList(store[f:f + 1], size),
                            ref_l=[structures.ImageList(store[i:i + 1], size) for i in ref_l],
                            ref_g=[structures.ImageList(store[i:i + 1], size) for i in ref_g],
                            frame_id=f, start_id=0, end_id=L - 1, seg_len=L, frame_category=0 if f == 0 else 1,
                            video_id=0))
    img_bytes = frames[0].numel() * frames.element_size()
    n_ref = sum(len(s["ref_l"]) + len(s["ref_g"]) for s in samples)
    return samples, n_ref * img_bytes


def run_clip(model, samples, to_host):
    """One clip through the public call, driven like mega_core/engine/inference.py:66-78 drives the reference: every
    output is moved to the CPU right after the call and stored; the detections are then all read on the host (still
    inside the caller's timed region)."""
    n_out = 0
    d2h = 0
    kept = []
    for s in samples:
        out = model(s)
        if to_host:
            kept.append([o.to("cpu") for o in out])
        n_out += len(out)
    for out in kept:
        for bl in out:
            b = bl.bbox; sc = bl.get_field("scores"); lb = bl.get_field("labels")
            assert not b.is_cuda and b.shape[0] == sc.shape[0] == lb.shape[0]
            d2h += b.numel() * 4 + sc.numel() * 4 + lb.numel() * 8
    return n_out, d2h


def timed(model, samples, steps, to_host, dist, dev):
    from diffusionvid_b200 import ops
    if dist:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    l0 = ops.LAUNCHES
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    frames = 0
    d2h = 0
    for _ in range(steps):
        n, b = run_clip(model, samples, to_host)
        frames += n
        d2h += b
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
        torch.distributed.barrier()
    return ms, frames, ops.LAUNCHES - l0, d2h


def run_ours(args):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the product path has no CPU fallback "
                         "(use --impl reference for the CPU oracle arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if dist:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from diffusionvid_b200 import model as pm, ops, synth

    hp = dict(HP_BASE, num_proposals=args.proposals, sample_step=args.T, device=str(dev))
    if args.backbone == "swinb":
        hp.update(swin=dict(embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32)), infer_batch=4,
                  all_frame_interval=4)
    m = pm.DiffusionDet(hp)
    m.load_state_dict(synth.make_state_dict(seed=1234, blocks=hp["blocks"], swin=hp.get("swin")), strict=False)
    m.to(dev)
    if args.frames_per_stream > 0:
        m.frames_per_stream = args.frames_per_stream
    shard_frames = dist and args.shard == "frames"
    if shard_frames:
        m.set_frame_sharding(rank, world)
    dev_samples, _ = make_clip_inputs(args, dev, pinned=False)
    host_samples, h2d_bytes = make_clip_inputs(args, dev, pinned=True)

    with torch.no_grad():
        for _ in range(max(3, args.warmup)):
            run_clip(m, dev_samples, False)
        sampler = ClockSampler(local) if rank == 0 else None
        ms, frames, launches, _ = timed(m, dev_samples, args.steps, False, dist, dev)
        clocks = sampler.stop() if sampler else None
        m.host_results = True      # e2e: detections are delivered on the host (one packed D2H copy per key batch)
        run_clip(m, host_samples, True)
        io0 = dict(m.io_bytes)
        ms_e2e, frames_e2e, _, d2h = timed(m, host_samples, args.steps, True, dist, dev)
        h2d_bytes = (m.io_bytes["h2d"] - io0["h2d"]) // max(1, args.steps)     # counted from the tensors copied
        d2h = m.io_bytes["d2h"] - io0["d2h"]
        # clip-loader mode: the same clip as decoded 8-bit frames in pinned host memory (ToTensor fused on the GPU)
        u8_samples, _ = make_clip_inputs(args, dev, pinned=True, u8=True)
        for _ in range(2):
            run_clip(m, u8_samples, True)
        io1 = dict(m.io_bytes)
        ms_u8, frames_u8, _, _ = timed(m, u8_samples, args.steps, True, dist, dev)
        h2d_u8 = (m.io_bytes["h2d"] - io1["h2d"]) // max(1, args.steps)
        d2h_u8 = (m.io_bytes["d2h"] - io1["d2h"]) // max(1, args.steps)
        del u8_samples
        m.host_results = False

        roof = None
        if rank == 0 and not args.no_roofline:
            # per-launch CUDA events need eager, single-stream execution (the timed runs above replay CUDA graphs)
            ug, us = m.use_graphs, m.use_streams
            m.use_graphs = m.use_streams = False
            run_clip(m, dev_samples, False)
            ops.PROFILE = {}
            run_clip(m, dev_samples, False)
            torch.cuda.synchronize()
            prof = ops.PROFILE
            ops.PROFILE = None
            m.use_graphs, m.use_streams = ug, us
            fam = {}
            shapes = []
            for k, (evs, flops, nbytes) in prof.items():
                if ":" in k:      # per-shape entry (tools/profile_shapes.py prints these in full)
                    t = sum(s.elapsed_time(e) for s, e in evs)
                    shapes.append((t, k, len(evs), flops))
                    continue
                t = sum(s.elapsed_time(e) for s, e in evs)
                fam[k] = {"ms": t, "launches": len(evs), "tflops": flops / t / 1e9 if t > 0 else 0.0,
                          "gbs": nbytes / t / 1e6 if t > 0 else 0.0}
            peak_tf, peak_bw, src = peaks()
            cg = fam.get("conv_gemm", {"ms": 0.0, "tflops": 0.0, "launches": 0})
            clip_ms = ms / args.steps
            roof = {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit-GEMM conv + all decoder GEMMs)",
                    "achieved": cg["tflops"], "peak": peak_tf, "unit": "TFLOP/s",
                    "frac": cg["tflops"] / peak_tf if peak_tf else None, "peak_source": src, "traffic": None,
                    "launches_per_step": cg["launches"], "avg_launch_us": 1000.0 * cg["ms"] / max(1, cg["launches"]),
                    "share_of_step": cg["ms"] / clip_ms if clip_ms > 0 else None,
                    "note": "achieved = algorithmic FLOPs of all conv_gemm_kernel launches of one clip / their summed "
                            "CUDA-event time (eager instrumented pass); share_of_step relates that eager kernel time "
                            "to the graph-replayed step",
                    "families": {k: {kk: round(vv, 3) for kk, vv in v.items()} for k, v in fam.items()},
                    "top_shapes": [{"shape": k.split(":", 1)[1], "launches": n, "ms": round(t, 3),
                                    "tflops": round(fl / t / 1e9, 1) if t > 0 else 0.0}
                                   for t, k, n, fl in sorted(shapes, reverse=True)[:8]]}
            rd = fam.get("roi_dynconv")
            if rd and peak_bw:
                # the fused ROIAlign + DynamicConv step against the HBM roofline (north_star): algorithmic bytes per box
                # = generated weights in (64 KB) + ROI-tile-equivalent of feature reads and the 49x256 result (2 x 25 KB)
                roof["decoder_step"] = {"kernel": "roi_dynconv_kernel (fused ROIAlign + DynamicConv bmm/LN)",
                                        "bound": "hbm", "achieved": rd["gbs"], "peak": peak_bw, "unit": "GB/s",
                                        "frac": rd["gbs"] / peak_bw, "launches_per_step": rd["launches"],
                                        "avg_launch_us": 1000.0 * rd["ms"] / max(1, rd["launches"]),
                                        "share_of_step": rd["ms"] / clip_ms if clip_ms > 0 else None}
            tf = os.path.join(ROOT, "profiles", "conv_gemm_traffic.json")
            if os.path.exists(tf):       # dram bytes per launch from the committed ncu --set full capture
                with open(tf) as f:
                    tj = json.load(f)
                roof["traffic"] = tj.get("dram_bytes_per_launch")
                roof["traffic_source"] = tj.get("source")

    mult = 1 if shard_frames else world       # frame sharding: all ranks work on the same clip
    total_frames = frames * mult
    value = total_frames / (ms / 1000.0)
    e2e_value = frames_e2e * mult / (ms_e2e / 1000.0)
    e2e_u8_value = frames_u8 * mult / (ms_u8 / 1000.0)
    if rank != 0:
        if dist:
            torch.distributed.destroy_process_group()
        return
    cpu = None
    if not args.no_cpu_baseline and world == 1:
        fps, desc, _ = cpu_oracle_run(args, args.cpu_sample_frames, 1, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "sample": desc}
    line = {"metric": "frames/sec", "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if shard_frames else "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic",
            "config": {"workload": "%s N=%d T=%d fp16, %dx%d clip of %d frames + %d global frames per step, %s "
                                   "across ranks"
                                   % ("vid_Swin_B_DiffusionVID.yaml Swin-B+FPN" if args.backbone == "swinb"
                                      else "vid_R_101_DiffusionVID.yaml R-101+FPN",
                                      args.proposals, args.T, args.width, args.height, args.frames, args.global_frames,
                                      "frame-sharded" if shard_frames else "video-sharded"),
                       "frames_per_step": args.frames, "l2": "inputs larger than L2 (clip %.0f MB fp32, feature maps "
                                                             "52 MB per key batch)" % (args.frames * 7.47)},
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": d2h // max(1, args.steps), "ms_per_step": ms_e2e / args.steps,
                    "input": "fp32 [0,1] ImageLists in pinned host memory (the reference's ToTensor output)",
                    "loop": "mega_core/engine/inference.py:66-78 (call, outputs .to(cpu), store); deferred host results "
                            "(BoxList.deferred) let the host run ahead up to 4 key batches; every detection of the clip "
                            "is read on the host before the clip's timed region ends"},
            "e2e_u8": {"value": e2e_u8_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_u8,
                       "d2h_bytes_per_step": d2h_u8, "ms_per_step": ms_u8 / args.steps,
                       "input": "uint8 ImageLists in pinned host memory (decoded frames; ToTensor fused into "
                                "dvid_preprocess_u8) - same public call"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if dist:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

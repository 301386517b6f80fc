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
  e2e_engine : the e2e clip driven EXACTLY like mega_core/engine/inference.py:26-93 with its timer on: every call first
          moves `cur` and every ref ImageList to the device (:35-40; `cur` is a second upload of a frame that also
          arrives as a ref), the model is called, torch.cuda.synchronize() (:70-73), outputs .to(cpu) (:75).
  cpu_baseline / --impl reference : the fp32 CPU oracle (oracle/model.py, a PyTorch restatement of the reference; the
          reference itself cannot be imported - SURVEY.md 8c) on the host cores, on a bounded sample of the workload:
          one key batch of 8 frames + 13 global frames (975 -> 900 farthest-point sampling active), N=300, T=4.
  parity : the product against the oracle ON THE SAME SAMPLE inside the cpu_baseline leg: detections of that clip
          (two-sided matched fraction), and on two frames the per-head box / logit deltas on identical inputs against
          the fp16-emulating oracle + index-exact top-k/NMS on identical inputs.

N>1 (torchrun): video-level sharding as in the reference (VIDTestDistributedSampler): every rank processes its own
clips, no data-path collective; value = total frames of all ranks / max-over-ranks time ("weak" scaling).  The line
additionally carries BASELINE config 5 - ONE clip spread over the N ranks (strong scaling), in both granularities of
DiffusionDet.set_frame_sharding: `frame_sharded` (frame i of every call on rank i % N: one all-gather of memory
candidates per video + one all-gather of detections per key batch, every rank returns the whole clip) and
`batch_sharded` (key batch k on rank k % N, the global frames dealt frame by frame: the memory all-gather only).
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
    ap.add_argument("--cpu-sample-frames", type=int, default=8, help="key-batch frames of the CPU oracle sample")
    ap.add_argument("--cpu-sample-global", type=int, default=13, help="global frames of the CPU oracle sample")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-frame-sharded", action="store_true", help="N>1: skip the frame-sharded sub-measurement")
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
    frames = synth.make_clip(frames_n + global_n, args.height, args.width, seed=1234)
    samples = synth.clip_samples(frames[:frames_n], [], args.height, args.width)
    samples[0]["ref_g"] = [frames[frames_n + i][None] for i in range(global_n)]      # global frames: distinct images
    times = []
    last = None
    for r in range(repeats):
        o = om.OracleDiffusionVID(sd, ocfg, fp16=False, noise=om.NoiseSource(1234, args.proposals))
        t0 = time.perf_counter()
        res = []
        with torch.no_grad():
            for s in samples:
                res += o.forward(s)
        times.append(time.perf_counter() - t0)
        assert len(res) == frames_n
        last = (o, res)
    best = min(times) if times else float("nan")
    desc = ("%d-frame clip + %d global frames, N=%d, T=%d, %dx%d, R-101+FPN, fp32 PyTorch oracle, %d threads"
            % (frames_n, global_n, args.proposals, args.T, args.width, args.height, torch.get_num_threads()))
    return frames_n / best, desc, times, (sd, samples, last)


def cpu_sample_for_budget(args, steps, budget_s=240.0):
    """--impl reference runs the sample (steps + 1 warm-up) times: pick the largest sample clip whose run still ends
    within a few minutes.  Costs are the measured 16-thread times of the fp32 oracle per sample (profiles/README.md);
    they only steer the choice, the line reports what was actually run."""
    est = [((8, 13), 45.0), ((8, 4), 28.0), ((4, 2), 14.0), ((2, 1), 6.0)]
    want = (args.cpu_sample_frames, args.cpu_sample_global)
    for (f, g), cost in est:
        if f <= want[0] and g <= want[1] and cost * (steps + 1) <= budget_s:
            return f, g
    return 2, 1


# ---------------------------------------------------------------------------------------------------- parity (cpu leg)
def parity_report(args, dev, pack):
    """The product against the oracle on the cpu_baseline sample (part of the cpu_baseline leg: the oracle is used
    here as the CHECKER of the product's outputs, never on the measured path).
      match        two-sided matched fraction of the final detections of the sample's key batch vs the fp32 oracle run
                   that was just timed (box 2e-3 of the image size, score 4e-3); for T > 1 one differing renewal decision
                   re-deals a whole frame's step noise, so the number of such flips is reported next to it
      box_delta /  per-head differences on IDENTICAL inputs (teacher forced) against the fp16-emulating oracle, two
      logit_delta  frames, three base heads + the conditioned head on the oracle's own 900-row memory
      keep_equal   top-k as a set and NMS keep order index-exact when fed the oracle's logits / candidate lists"""
    from diffusionvid_b200 import model as pm, ops, structures
    from oracle import model as om, ops as oo
    from tests.parity_util import match_fraction_two_sided, quantiles
    sd, samples, (o32, ref) = pack
    h, w, N, T = args.height, args.width, args.proposals, args.T
    size = float(max(h, w))
    hp = dict(HP_BASE, num_proposals=N, sample_step=T, device=str(dev))
    m = pm.DiffusionDet(hp)
    m.load_state_dict(sd, strict=False)
    m.to(dev)
    m.noise = om.NoiseSource(1234, N)
    m.debug_trace = True
    L = len(samples)
    got = []
    for s in samples:
        got += m(dict(cur=structures.ImageList(s["cur"].to(dev), [(h, w)]),
                      ref_l=[structures.ImageList(t.to(dev), [(h, w)]) for t in s["ref_l"]],
                      ref_g=[structures.ImageList(t.to(dev), [(h, w)]) for t in s["ref_g"]],
                      frame_id=s["frame_id"], start_id=0, end_id=L - 1, seg_len=L,
                      frame_category=s["frame_category"], video_id=0))
    fr = [match_fraction_two_sided(g.bbox.cpu(), g.get_field("scores").cpu(), g.get_field("labels").cpu(),
                                   r["boxes"], r["scores"], r["labels"], size, box_tol=2e-3, score_tol=4e-3)
          for g, r in zip(got, ref)]
    flips = 0
    for si in range(max(0, T - 1)):
        kp = torch.sigmoid(m.last_trace[("logits", 0, si, 0)]).max(-1)[0].cpu() > 0.5
        ko = torch.sigmoid(o32.trace[("logits", 0, si)]).max(-1)[0] > 0.5
        flips += int((kp != ko).sum())
    out = {"checker": "oracle/model.py on the host (fp32 run = the timed cpu_baseline sample; fp16-emulating run for the "
                      "per-head deltas)",
           "tolerance": {"box": 1e-3, "logit": "fp16 storage: p99.9 <= 2e-2 (north_star's 1e-4 is below one fp16 ulp)",
                         "match": "box 2e-3 of the image size, score 4e-3, same label, two-sided"},
           "match_frac": {"median": sorted(fr)[len(fr) // 2], "mean": sum(fr) / len(fr), "frames": len(fr),
                          "renewal_flips_vs_fp32_oracle": flips, "T": T}}
    # ---- post-processing on identical inputs
    cap = max(1, T - 1) * N
    B = len(ref)
    ens_b = torch.empty((B, cap, 4), device=dev); ens_s = torch.empty((B, cap), device=dev)
    ens_l = torch.empty((B, cap), device=dev, dtype=torch.int32)
    steps = range(T - 1) if T > 1 else range(1)
    topk_equal, want = True, [[] for _ in range(B)]
    for j, si in enumerate(steps):
        lg, bx = o32.trace[("logits", 0, si)], o32.trace[("coord", 0, si)]
        ops.topk_scores(lg.to(dev).contiguous(), bx.to(dev).contiguous(), N, ens_b, ens_s, ens_l, j * N)
        for i in range(B):
            wb, ws, wl = om.topk_scores(lg[i], bx[i], N)
            want[i].append((wb, ws, wl))
            gb = ens_b[i, j * N:(j + 1) * N].cpu(); gl = ens_l[i, j * N:(j + 1) * N].cpu().long()
            key = lambda b, l: sorted(zip(l.tolist(), map(tuple, b.tolist())))      # noqa: E731
            topk_equal &= key(gb, gl) == key(wb, wl)
    # NMS on the ORACLE's candidate lists (its boxes / scores / labels, bit for bit): keep order must be index-exact
    for i in range(B):
        ens_b[i] = torch.cat([e[0] for e in want[i]]).to(dev)
        ens_s[i] = torch.cat([e[1] for e in want[i]]).to(dev)
        ens_l[i] = torch.cat([e[2] for e in want[i]]).to(dev).int()
    r = ops.nms(ens_b, ens_s, ens_l, thr=0.5, clip_wh=(float(w), float(h)))
    cnt = r["count"].cpu().tolist()
    nms_equal = True
    for i in range(B):
        fin = om.finalize_frame(torch.cat([e[0] for e in want[i]]), torch.cat([e[1] for e in want[i]]),
                                torch.cat([e[2] for e in want[i]]), (w, h), True)
        n = fin["scores"].numel()
        nms_equal &= (cnt[i] == n and torch.equal(r["boxes"][i, :n].cpu(), fin["boxes"])
                      and torch.equal(r["scores"][i, :n].cpu(), fin["scores"])
                      and torch.equal(r["labels"][i, :n].cpu().long(), fin["labels"]))
    out["keep_equal"] = {"topk_set_equal": bool(topk_equal), "nms_keep_order_equal": bool(nms_equal), "frames": B,
                         "candidates_per_frame": cap}
    # ---- per-head deltas on identical inputs vs the fp16-emulating oracle (2 frames)
    Bt = min(2, L)
    o16 = om.OracleDiffusionVID(sd, dict(num_proposals=N, sample_step=T), fp16=True, noise=om.NoiseSource(1234, N))
    imgs = torch.cat([s["cur"] for s in samples[:Bt]])
    feats = o16.backbone(imgs)
    got_f = m.extract_features(imgs.to(dev))
    fd = max(((g.float().permute(0, 3, 1, 2).cpu() - f).abs().max() / f.abs().max()).item() for g, f in zip(got_f, feats))
    lv = ops.Levels([f.permute(0, 2, 3, 1).contiguous().half().to(dev) for f in feats])
    whwh = torch.tensor([w, h, w, h], dtype=torch.float32)[None].expand(Bt, -1)
    boxes = o16._x_to_boxes(om.NoiseSource(1234, N).get("img", 0, 0, 0, Bt), whwh)
    temb = om.time_embedding(o16.c, torch.full((Bt,), 999, dtype=torch.long))
    mem = o32.mem[0]
    m._set_memory([mem.to(dev).contiguous(), None])
    m._warm_constants([999])
    dbox, dlog = [], []
    pro = None
    names = ["head.head_series.%d." % i for i in range(3)] + ["head.head_series_cond.0."]
    for i, pre in enumerate(names):
        cond = om.global_attention(o16.c, pro, mem, o16.cfg) if i == 3 else None
        lg_r, bx_r, pro_r = om.rcnn_head(o16.c, pre, feats, boxes, pro, temb, o16.cfg, cond=cond)
        e = m._pk["cond"][0] if i == 3 else m._pk["heads"][i]
        p32 = None if pro is None else pro.to(dev).contiguous()
        p16 = None if pro is None else pro.half().to(dev).contiguous()
        shift = m._cond_shift(e, m._global_context(p16, Bt * N), Bt * N) if i == 3 else None
        lg, bx, _, _ = m._head(e, lv, boxes.to(dev).contiguous(), p32, p16, 999, shift_rows=shift)
        dbox.append((bx.cpu() - bx_r).abs() / size)
        dlog.append((lg.cpu() - lg_r).abs())
        boxes, pro = bx_r, pro_r
    qb, ql = quantiles(torch.cat([d.flatten() for d in dbox])), quantiles(torch.cat([d.flatten() for d in dlog]))
    out["box_delta"] = dict(qb, unit="fraction of the image size", heads=4, boxes=Bt * N, within_tolerance=qb["max"] <= 1e-3)
    out["logit_delta"] = dict(ql, within_tolerance=ql["p999"] <= 2e-2)
    out["feature_delta_rel_max"] = fd
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    k = max(1, args.steps)
    t0 = time.perf_counter()
    fr_n, gl_n = cpu_sample_for_budget(args, k)
    if args.warmup > 0:
        cpu_oracle_run(args, fr_n, gl_n, 1)
    fps, desc, times, _ = cpu_oracle_run(args, fr_n, gl_n, k)
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


def run_clip_engine(model, samples, dev):
    """The clip through the reference's own loop body (mega_core/engine/inference.py:26-93 with its timer): move `cur`
    and every ref ImageList to the device (:35-40), call the model, torch.cuda.synchronize() (:70-73), outputs to the
    CPU (:75), store (:91-93)."""
    n_out = 0
    results = {}
    cpu = torch.device("cpu")
    for i, s in enumerate(samples):
        images = dict(s)
        images["cur"] = images["cur"].to(dev)
        for key in ("ref_l", "ref_g"):
            images[key] = [img.to(dev) for img in images[key]]
        out = model(images)
        torch.cuda.synchronize()
        out = [o.to(cpu) for o in out]
        results.update({i + j: o for j, o in enumerate(out)})
        n_out += len(out)
    for bl in results.values():
        assert bl.bbox.shape[0] == bl.get_field("scores").shape[0]
    return n_out, 0


def timed(model, samples, steps, to_host, dist, dev, runner=None):
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
        n, b = runner(model, samples, dev) if runner else run_clip(model, samples, to_host)
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
    if int(os.environ.get("DVID_MAIN_PRIORITY", "0")) != 0:     # experiment: the caller's (backbone) stream above the decode streams
        torch.cuda.set_stream(torch.cuda.Stream(priority=int(os.environ["DVID_MAIN_PRIORITY"])))
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
        # the clock sampler starts BEFORE the warm-up: nvidia-smi's start-up takes driver locks that stall kernel
        # launches for ~0.1 s, which must not fall into a timed region of a few hundred ms
        sampler = ClockSampler(local) if rank == 0 else None
        for _ in range(max(3, args.warmup)):
            run_clip(m, dev_samples, False)
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
        # optional dead-code elimination (model.skip_unused_base, OFF for every other number of this line): the t=999
        # base stages of local frames, which the reference computes but never reads when SAMPLE_STEP > 1
        dce = None
        if args.T > 1:
            m.skip_unused_base = True
            for _ in range(2):
                run_clip(m, host_samples, True)
            ms_d2, fr_d2, _, _ = timed(m, host_samples, args.steps, True, dist, dev)
            m.host_results = False
            for _ in range(2):
                run_clip(m, dev_samples, False)
            ms_d1, fr_d1, _, _ = timed(m, dev_samples, args.steps, False, dist, dev)
            m.host_results = True
            m.skip_unused_base = False
            dce = {"value": fr_d1 * (1 if shard_frames else world) / (ms_d1 / 1000.0),
                   "e2e": fr_d2 * (1 if shard_frames else world) / (ms_d2 / 1000.0), "unit": "frames/s",
                   "what": "same workload and detections (bit-identical, tests/test_gpu_model.py), without the 3 base-stage "
                           "head evaluations per local frame that diffusion_det.py:438-460 computes and only the "
                           "SAMPLE_STEP == 1 branch (box_head.py:300-302) reads; NOT used for value / e2e / roofline"}
        # the reference's engine loop verbatim (uploads of cur + refs per call, a device synchronisation per call)
        for _ in range(2):
            run_clip_engine(m, host_samples, dev)
        io2 = dict(m.io_bytes)
        ms_eng, frames_eng, _, _ = timed(m, host_samples, args.steps, True, dist, dev, runner=run_clip_engine)
        img_bytes = host_samples[0]["cur"].tensors.numel() * host_samples[0]["cur"].tensors.element_size()
        n_moved = sum(1 + len(s["ref_l"]) + len(s["ref_g"]) for s in host_samples)
        d2h_eng = (m.io_bytes["d2h"] - io2["d2h"]) // max(1, args.steps)
        m.host_results = False

        # N > 1: the SAME clip dealt frame by frame to the ranks (BASELINE config 5)
        fs = bs = None
        if dist and not shard_frames and not args.no_frame_sharded:
            m.set_frame_sharding(rank, world)
            for _ in range(2):
                run_clip(m, dev_samples, False)
            cb0, ne0 = dict(m.comm_bytes), len(m.comm_events)
            ms_fs, frames_fs, _, _ = timed(m, dev_samples, args.steps, False, dist, dev)
            ag_us = [1000.0 * a.elapsed_time(b) for a, b in m.comm_events[ne0:]]
            fs = {"value": frames_fs / (ms_fs / 1000.0), "unit": "frames/s", "scaling": "strong",
                  "ms_per_step": ms_fs / args.steps,
                  "memory_allgather_bytes_per_video": (m.comm_bytes["memory"] - cb0["memory"]) // max(1, args.steps),
                  "memory_allgather_us": statistics.median(ag_us) if ag_us else None,
                  "results_allgather_bytes_per_step": (m.comm_bytes["results"] - cb0["results"]) // max(1, args.steps),
                  "collectives_per_step": "1 all-gather of memory candidates per video + 1 all-gather of detections "
                                          "per key batch (NCCL)"}
            # the same clip sharded by whole key batches: batch k on rank k % world, no per-batch exchange
            m.set_frame_sharding(rank, world, mode="batches")
            for _ in range(2):
                run_clip(m, dev_samples, False)
            cb0 = dict(m.comm_bytes)
            ms_bs, frames_bs, _, _ = timed(m, dev_samples, args.steps, False, dist, dev)
            tot = torch.tensor([frames_bs], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(tot)
            bs = {"value": float(tot.item()) / (ms_bs / 1000.0), "unit": "frames/s", "scaling": "strong",
                  "ms_per_step": ms_bs / args.steps, "frames_per_step_all_ranks": float(tot.item()) / args.steps,
                  "memory_allgather_bytes_per_video": (m.comm_bytes["memory"] - cb0["memory"]) // max(1, args.steps),
                  "collectives_per_step": "1 all-gather of memory candidates per video (NCCL); every rank returns the "
                                          "BoxLists of its own key batches, merged at the end like "
                                          "engine/inference.py:96-116"}
            m.set_frame_sharding(rank, 1)

        # device time of the captured execution units: graph replays, one unit at a time (the decode/backbone overlap is
        # switched off for this pass and for the per-kernel pass below - co-running units stretch each other's duration)
        ov = m.overlap_decode
        m.overlap_decode = False
        run_clip(m, dev_samples, False)
        m.unit_events = {}
        run_clip(m, dev_samples, False)
        torch.cuda.synchronize()
        units = {k: [a.elapsed_time(b) for a, b in v] for k, v in m.unit_events.items()}
        m.unit_events = None

        roof = None
        if rank != 0 or args.no_roofline:
            m.overlap_decode = ov
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
            m.overlap_decode = ov
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
            dec = units.get(("decode", hp["infer_batch"]))
            if dec and args.backbone == "r101":
                # one head evaluation at M = 8 x 300 boxes = 72.6 GFLOP and 82 MB of algorithmic bytes
                # (feature pyramids 52.3 MB + head weights 27.4 MB + object features / boxes, each tensor once, the
                # generated DynamicConv weights NOT counted: fused they never leave the chip); global attention 3.1 GF
                evals = args.T * 4 if args.T > 1 else 1
                dms = statistics.median(dec)
                gf = evals * 72.6 * args.proposals / 300.0 + args.T * 3.1
                roof["decoder_step"] = {
                    "kernel": "decode unit of one 8-frame key batch: %d head evaluations (ROIAlign + self-attention + "
                              "DynamicConv + FFN + towers + apply_deltas) + %d global attentions + DDIM + top-k + NMS, "
                              "one CUDA-graph replay timed with CUDA events" % (evals, args.T),
                    "unit_ms": dms, "head_eval_us": 1000.0 * dms / evals, "bound": "tensor",
                    "schedule": "measured with the units serialised; in the timed runs the decode unit of batch k runs on its "
                                "own stream next to the backbone of batch k+1",
                    "achieved": gf / dms, "peak": peak_tf, "unit": "TFLOP/s", "frac": gf / dms / peak_tf,
                    "hbm": {"algorithmic_bytes_per_head_eval": 82e6, "achieved_gbs": 82e6 * evals / dms / 1e6,
                            "peak_gbs": peak_bw, "frac": 82e6 * evals / dms / 1e6 / peak_bw},
                    "roi_dynconv_kernel": roof.pop("decoder_step", None)}
            # whole-step algorithmic work: backbone GFLOP per frame (R-101+FPN 213.08 hook-counted,
            # Swin-B+FPN 423.4 analytic incl. window padding) on every local + global frame, 3 base head evaluations per
            # frame at extraction, T x (3 + 1) head evaluations + T global attentions per local frame in the DDIM loop
            bb_gf = 423.4 if args.backbone == "swinb" else 213.08
            he_gf = 72.6 / 8.0 * args.proposals / 300.0
            nfr = args.frames + args.global_frames
            dec_evals = (args.T * 4 if args.T > 1 else 1)
            step_gf = nfr * (bb_gf + 3 * he_gf) + args.frames * (dec_evals * he_gf + args.T * 3.1 / 8.0)
            roof["step"] = {"algorithmic_tflop_per_step": step_gf / 1e3, "achieved": step_gf / clip_ms,
                            "peak": peak_tf, "unit": "TFLOP/s", "frac": step_gf / clip_ms / peak_tf,
                            "backbone": "Swin-B+FPN 423.4 GFLOP/frame" if args.backbone == "swinb"
                            else "R-101+FPN 213.08 GFLOP/frame",
                            "backbone_share_of_flops": nfr * bb_gf / step_gf}
            ext = units.get(("extract", hp["infer_batch"])) or units.get(("features", hp["infer_batch"]))
            if ext:
                roof["extract_unit_ms"] = statistics.median(ext)
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
    e2e_eng_value = frames_eng * mult / (ms_eng / 1000.0)
    for rec in (fs, bs):
        if rec is not None:
            rec["efficiency_vs_video_sharded"] = rec["value"] / value       # value = N x the un-sharded per-GPU rate
            rec["speedup_vs_one_gpu"] = rec["value"] / (value / world)
    if rank != 0:
        if dist:
            torch.distributed.destroy_process_group()
        return
    cpu = parity = None
    if not args.no_cpu_baseline and world == 1 and args.backbone == "r101":
        fps, desc, _, pack = cpu_oracle_run(args, args.cpu_sample_frames, args.cpu_sample_global, 1)
        cpu = {"value": fps, "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port", "sample": desc}
        if not args.no_parity:
            with torch.no_grad():
                parity = parity_report(args, dev, pack)
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
            "e2e_engine": {"value": e2e_eng_value, "unit": "frames/s", "ms_per_step": ms_eng / args.steps,
                           "h2d_bytes_per_step": n_moved * img_bytes, "d2h_bytes_per_step": d2h_eng,
                           "loop": "mega_core/engine/inference.py:26-93 verbatim: cur + every ref ImageList .to(device) "
                                   "per call (:35-40; cur is a second upload of a frame that also arrives as a ref), "
                                   "model(images), torch.cuda.synchronize() per call (:70-73), outputs .to(cpu) (:75)"},
            "gpu_launches": launches, "clocks": clocks, "roofline": roof, "cpu_baseline": cpu, "parity": parity,
            "frame_sharded": fs, "batch_sharded": bs, "dead_code_elimination": dce}
    print(json.dumps(line))
    if dist:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)

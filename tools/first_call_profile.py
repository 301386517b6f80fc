import os, sys, time
sys.path.insert(0, os.getcwd())
import torch, bench
from diffusionvid_b200 import model as pm, synth
class A: pass
a = A(); a.frames=64; a.global_frames=24; a.height=600; a.width=1000; a.backbone="r101"
dev = torch.device("cuda", 0)
hp = dict(bench.HP_BASE, num_proposals=300, sample_step=4, device=str(dev))
m = pm.DiffusionDet(hp); m.load_state_dict(synth.make_state_dict(seed=1234, blocks=hp["blocks"]), strict=False); m.to(dev)
samples, _ = bench.make_clip_inputs(a, dev, pinned=False)
with torch.no_grad():
    for _ in range(3): bench.run_clip(m, samples, False)
    torch.cuda.synchronize()
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        t0 = time.perf_counter(); out = m(samples[0]); torch.cuda.synchronize(); t1 = time.perf_counter()
    print("first call ms", (t1 - t0) * 1e3)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
    print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=15, max_name_column_width=60))
    ev = sorted([e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA], key=lambda e: e.time_range.start)
    t00 = ev[0].time_range.start
    gaps = []
    for x, y in zip(ev[:-1], ev[1:]):
        g = y.time_range.start - x.time_range.end
        if g > 100: gaps.append((g, x.name[:50], y.name[:50], x.time_range.end - t00))
    print("GPU idle gaps > 100us:")
    for g in sorted(gaps, reverse=True)[:15]: print(g)

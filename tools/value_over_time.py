"""device-resident throughput in blocks of 5 clips (sync at block ends) - looks for drift over a long run"""
import sys, time
import torch
sys.path.insert(0, ".")
import bench
from diffusionvid_b200 import model as pm, synth

class A: pass
a = A(); a.frames=64; a.global_frames=24; a.height=600; a.width=1000; a.backbone="r101"
dev = torch.device("cuda:0")
hp = dict(bench.HP_BASE, num_proposals=300, sample_step=4, device=str(dev))
m = pm.DiffusionDet(hp); m.load_state_dict(synth.make_state_dict(seed=1234, blocks=hp["blocks"]), strict=False); m.to(dev)
samples, _ = bench.make_clip_inputs(a, dev, pinned=False)
with torch.no_grad():
    for _ in range(3):
        bench.run_clip(m, samples, False)
    torch.cuda.synchronize()
    for blk in range(8):
        t0 = time.perf_counter()
        for _ in range(5):
            bench.run_clip(m, samples, False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        st = torch.cuda.memory_stats()
        print("block %d: %.1f frames/s  reserved %.1f GB  allocs(cudaMalloc) %d" % (blk, 5 * 64 / dt, st["reserved_bytes.all.current"] / 1e9, st["num_device_alloc"]), flush=True)
    import gc
    for label, prep in (("gc on", gc.enable), ("gc off", gc.disable), ("gc on", gc.enable), ("gc off", gc.disable)):
        prep()
        ts = []
        for _ in range(12):
            t0 = time.perf_counter()
            bench.run_clip(m, samples, False)
            ts.append(time.perf_counter() - t0)      # host time to ENQUEUE a clip (the host runs ahead of the GPU)
        torch.cuda.synchronize()
        print(label, "host ms per clip:", " ".join("%.0f" % (1000 * t) for t in ts), flush=True)

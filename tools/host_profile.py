"""Where does the HOST time of a clip go?  cProfile of the enqueue path (device-resident and host-fed)."""
import cProfile, pstats, sys, io
import torch
sys.path.insert(0, ".")
import bench
from diffusionvid_b200 import model as pm, synth

class A: pass
a = A(); a.frames=64; a.global_frames=24; a.height=600; a.width=1000; a.backbone="r101"
dev = torch.device("cuda:0")
hp = dict(bench.HP_BASE, num_proposals=300, sample_step=4, device=str(dev))
m = pm.DiffusionDet(hp); m.load_state_dict(synth.make_state_dict(seed=1234, blocks=hp["blocks"]), strict=False); m.to(dev)
for pinned in (False, True):
    samples, _ = bench.make_clip_inputs(a, dev, pinned=pinned)
    m.host_results = pinned
    with torch.no_grad():
        for _ in range(3):
            bench.run_clip(m, samples, pinned)
        torch.cuda.synchronize()
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(3):
            bench.run_clip(m, samples, pinned)
        pr.disable()
        torch.cuda.synchronize()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22)
    print("==== pinned host frames" if pinned else "==== device-resident frames")
    print("\n".join(l[:150] for l in s.getvalue().splitlines()[:40]))

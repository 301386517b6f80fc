import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionvid_b200 import ops
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
B, N = 8, 300
feats = [torch.randn(B, h, w, 256, generator=g).half().to(dev) for h, w in ((76, 128), (38, 64), (19, 32))]
lv = ops.Levels(feats)
cx = torch.rand(B, N, 2, generator=g) * torch.tensor([1000., 600.])
wh = (torch.randn(B, N, 2, generator=g).clamp(-2, 2) / 4 + 0.5) * torch.tensor([1000., 600.])
boxes = torch.cat([cx - wh / 2, cx + wh / 2], -1).to(dev).contiguous()
params = torch.randn(B * N, 32768, generator=g).half().to(dev) * 0.05
ln = [torch.ones(64).to(dev), torch.zeros(64).to(dev), torch.ones(256).to(dev), torch.zeros(256).to(dev)]
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / reps
out = torch.empty(B * N, 49 * 256, device=dev, dtype=torch.float16)
print("ROI_WS", os.environ.get("DVID_ROI_WS", "1"), "roi_dynconv fused us %.1f" % t(lambda: ops.roi_dynconv(lv, boxes, N, params, *ln, out=out)))
print("roi_align us %.1f" % t(lambda: ops.roi_align(lv, boxes, N)))

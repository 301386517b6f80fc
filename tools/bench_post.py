import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionvid_b200 import ops
dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
B, N, C = 8, 300, 30
logits = (torch.randn(B, N, C, generator=g) * 2 - 2).to(dev)
xy = torch.rand(B, N, 2, generator=g) * 700
boxes = torch.cat([xy, xy + torch.rand(B, N, 2, generator=g) * 300 + 1], -1).to(dev)
cap = 900
eb = torch.empty(B, cap, 4, device=dev); es = torch.empty(B, cap, device=dev); el = torch.empty(B, cap, device=dev, dtype=torch.int32)
def t(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) * 1e3 / reps
print("topk_scores us", t(lambda: [ops.topk_scores(logits, boxes, N, eb, es, el, i * N) for i in range(3)]) / 3)
print("nms us", t(lambda: ops.nms(eb, es, el, thr=0.5, clip_wh=(1000., 600.))))
print("topk_mask us", t(lambda: ops.topk_mask(logits, 75, 25)))
# fused head tail (cls / reg towers + predictors + apply_deltas), M = 2400
M = B * N
mk = lambda n: (torch.randn(n, 256, generator=g) / 16).half().to(dev)
ln = lambda: ((1 + 0.1 * torch.randn(256, generator=g)).to(dev), (0.1 * torch.randn(256, generator=g)).to(dev))
fc = torch.randn(M, 256, generator=g).half().to(dev)
cls = (mk(256), ln()); reg = [(mk(256), ln()) for _ in range(3)]
lw = torch.zeros(32, 256).half(); lw[:C] = (torch.randn(C, 256, generator=g) / 16).half(); lw = lw.to(dev)
dw = torch.zeros(16, 256).half(); dw[:4] = (torch.randn(4, 256, generator=g) / 64).half(); dw = dw.to(dev)
lb = (0.1 * torch.randn(C, generator=g)).to(dev); db = (0.1 * torch.randn(4, generator=g)).to(dev)
bx = boxes.reshape(M, 4).contiguous()
print("head_tail us", t(lambda: ops.head_tail(fc, cls, reg, lw, lb, C, dw, db, bx)))
# farthest point sampling of the global memory (diffusion_det.py:841-896): 1800 -> 900 and 600 -> 150 candidates
for n, mm in ((1800, 900), (600, 150)):
    feats = torch.randn(n, 256, generator=g).to(dev)
    dist = ops.cdist(feats)
    idx = torch.empty((1, mm), device=dev, dtype=torch.int32)
    def run():
        temp = torch.full((1, n), 1e10, device=dev)
        ops.furthest_point_sampling(1, n, mm, dist, temp, idx)
    print("fps %d->%d us" % (n, mm), t(run, reps=5))

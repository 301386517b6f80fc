import sys
import torch
sys.path.insert(0, ".")
from diffusionvid_b200 import ops
g = torch.Generator().manual_seed(1)
frames, n = 8, 300
M = frames * n
feats = [torch.randn(frames, h, w, 256, generator=g).half().cuda() for h, w in ((76, 128), (38, 64), (19, 32))]
lv = ops.Levels(feats)
cx = torch.rand(frames, n, 4, generator=g)
ctr = cx[..., :2] * torch.tensor([1000., 600.])
wh = cx[..., 2:] * torch.tensor([1000., 600.]) * 0.8 + 4
boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).contiguous().cuda()
params = (torch.randn(M, 32768, generator=g) * 0.1).half().cuda()
ln = [t.cuda() for t in (torch.ones(64), torch.zeros(64), torch.ones(256), torch.zeros(256))]
out = torch.empty((M, 49 * 256), device="cuda", dtype=torch.float16)
for tc in (True, False):
    for _ in range(4):
        ops.roi_dynconv(lv, boxes, n, params, *ln, out=out, transposed=tc)
torch.cuda.synchronize()

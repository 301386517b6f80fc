#!/usr/bin/env python
"""DVID_TRACE=1 timeline of CTA 0 for a res4-style conv3 (1x1 256->1024 + residual) or other conv shapes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionvid_b200 import ops
dev = torch.device("cuda")
n, h, w, cin, cout, R = [int(x) for x in (sys.argv[1:7] if len(sys.argv) > 6 else (8, 38, 64, 256, 1024, 1))]
resid = int(sys.argv[7]) if len(sys.argv) > 7 else 1
x = torch.randn(n, h, w, cin).half().to(dev)
wt = (torch.randn(cout, R * R * cin) / (R * R * cin) ** 0.5).half().to(dev)
b = torch.randn(cout).to(dev)
rs = torch.randn(n, h, w, cout).half().to(dev) if resid else None
out = torch.empty(n, h, w, cout, device=dev, dtype=torch.float16)
ops.conv_streamk(bool(int(os.environ.get("DVID_STREAMK", "0"))))
for _ in range(3):
    ops.conv2d(x, wt, b, cout, R, R, 1, R // 2, True, resid=rs, out=out)
torch.cuda.synchronize()
os.environ["DVID_TRACE"] = "1"
ops.conv2d(x, wt, b, cout, R, R, 1, R // 2, True, resid=rs, out=out)
torch.cuda.synchronize()

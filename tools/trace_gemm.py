#!/usr/bin/env python
"""DVID_TRACE=1 python tools/trace_gemm.py : per-role event timeline of CTA 0 for the dynamic_layer GEMM."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionvid_b200 import ops
dev = torch.device("cuda")
m, k, n = 2400, 256, 32768
a = torch.randn(m, k).half().to(dev); w = (torch.randn(n, k) / 16).half().to(dev); b = torch.randn(n).to(dev)
out = torch.empty(m, n, device=dev, dtype=torch.float16)
os.environ.pop("DVID_TRACE", None)
for _ in range(3):
    ops.gemm(a, w, b, out=out)
torch.cuda.synchronize()
os.environ["DVID_TRACE"] = "1"
ops.gemm(a, w, b, out=out)
torch.cuda.synchronize()

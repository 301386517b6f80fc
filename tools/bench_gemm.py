#!/usr/bin/env python
"""Micro-benchmark of conv_gemm_kernel on the shapes that dominate the step (CUDA events, L2 flushed between reps)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from diffusionvid_b200 import ops

dev = torch.device("cuda")
g = torch.Generator(device="cpu").manual_seed(0)
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def conv_case(name, n, h, w, cin, cout, R, stride, resid=False, relu=True):
    x = torch.randn(n, h, w, cin, generator=g).half().to(dev)
    wt = (torch.randn(cout, R * R * cin, generator=g) / (R * R * cin) ** 0.5).half().to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    ho, wo = (h + 2 * (R // 2) - R) // stride + 1, (w + 2 * (R // 2) - R) // stride + 1
    rs = torch.randn(n, ho, wo, cout, generator=g).half().to(dev) if resid else None
    out = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.conv2d(x, wt, b, cout, R, R, stride, R // 2, relu, resid=rs, out=out))
    fl = 2.0 * n * ho * wo * cout * R * R * cin
    by = 2.0 * (x.numel() + wt.numel() + out.numel() * (2 if resid else 1))
    print("%-28s %8.1f us  %7.1f TFLOP/s  %6.2f TB/s(min bytes)" % (name, us, fl / us / 1e6, by / us / 1e6))


def gemm_case(name, m, k, n, relu=False):
    a = torch.randn(m, k, generator=g).half().to(dev)
    wt = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(dev)
    b = torch.randn(n, generator=g).to(dev)
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    us = timeit(lambda: ops.gemm(a, wt, b, relu=relu, out=out))
    print("%-28s %8.1f us  %7.1f TFLOP/s  %6.2f TB/s(min bytes)" % (name, us, 2.0 * m * n * k / us / 1e6, 2.0 * (m * k + n * k + m * n) / us / 1e6))


print("DVID_DBG =", os.environ.get("DVID_DBG", "0"), " DVID_STREAMK =", os.environ.get("DVID_STREAMK", "0"))
ops.conv_streamk(bool(int(os.environ.get("DVID_STREAMK", "0"))))
B = 8
conv_case("res2.conv3 64->256 +res", B, 152, 256, 64, 256, 1, 1, resid=True)
conv_case("res2.conv2 3x3 64->64", B, 152, 256, 64, 64, 3, 1)
conv_case("res3.conv1 512->128", B, 76, 128, 512, 128, 1, 1)
conv_case("res3.conv3 128->512 +res", B, 76, 128, 128, 512, 1, 1, resid=True)
conv_case("res3.conv2 3x3 128->128", B, 76, 128, 128, 128, 3, 1)
conv_case("res4.conv1 1024->256", B, 38, 64, 1024, 256, 1, 1)
conv_case("res4.conv2 3x3 256->256", B, 38, 64, 256, 256, 3, 1)
conv_case("res4.conv3 256->1024 +res", B, 38, 64, 256, 1024, 1, 1, resid=True)
conv_case("res5.conv2 3x3 512->512", B, 19, 32, 512, 512, 3, 1)
conv_case("fpn_output4 3x3 256->256", B, 38, 64, 256, 256, 3, 1, relu=False)
conv_case("fpn_output3 3x3 256->256", B, 76, 128, 256, 256, 3, 1, relu=False)
gemm_case("dynamic_layer 2400x256->32768", 2400, 256, 32768)
gemm_case("linear1 2400x256->2048", 2400, 256, 2048, relu=True)
gemm_case("qkv 2400x256->768", 2400, 256, 768)

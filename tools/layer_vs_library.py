#!/usr/bin/env python
"""Every distinct convolution shape of the R-101+FPN backbone (8 frames, 608x1024) and every GEMM shape of a head
evaluation (M = 2400 boxes), hand-written conv_gemm_kernel vs the library on the same B200, same call:

  dvid   ops.conv2d / ops.gemm (NHWC fp16, bias + ReLU / residual fused as in the model)
  lib    F.conv2d on channels_last fp16 tensors with cudnn.benchmark (+ separate add / ReLU kernels exactly where the
         reference has them: FrozenBN is folded into the weights for both arms), torch.nn.functional.linear for GEMMs

Both arms are timed as a CUDA graph of `reps` back-to-back launches (launch overhead excluded for both), median of 10
replays, L2 warm (the model runs these back to back on L2-resident activations).  Writes
gpurun_out/layer_vs_library.json -> profiles/r02_layer_vs_library.json."""
import json
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from diffusionvid_b200 import ops  # noqa: E402

dev = torch.device("cuda")
g = torch.Generator().manual_seed(0)
torch.backends.cudnn.benchmark = True


def graph_time(fn, reps=10, replays=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(replays):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); gr.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000.0 / reps)
    ts.sort()
    return ts[len(ts) // 2]


def conv_case(name, count, n, h, w, cin, cout, R, stride, resid=False, relu=True):
    pad = R // 2
    x = torch.randn(n, h, w, cin, generator=g).half().to(dev)
    wt = (torch.randn(cout, R, R, cin, generator=g) / (R * R * cin) ** 0.5).half().to(dev)
    b = torch.randn(cout, generator=g).to(dev)
    ho, wo = (h + 2 * pad - R) // stride + 1, (w + 2 * pad - R) // stride + 1
    rs = torch.randn(n, ho, wo, cout, generator=g).half().to(dev) if resid else None
    out = torch.empty(n, ho, wo, cout, device=dev, dtype=torch.float16)
    w2 = wt.view(cout, -1).contiguous()
    t_dvid = graph_time(lambda: ops.conv2d(x, w2, b, cout, R, R, stride, pad, relu, resid=rs, out=out))
    # library arm: NCHW logical / channels_last physical = the same NHWC bytes
    xl = x.permute(0, 3, 1, 2)
    wl = wt.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
    bl = b.half()
    rl = rs.permute(0, 3, 1, 2) if resid else None

    def lib():
        y = F.conv2d(xl, wl, bl, stride=stride, padding=pad)
        if rl is not None:
            y = y + rl
        if relu:
            y = F.relu_(y)
        return y
    t_lib = graph_time(lib)
    fl = 2.0 * n * ho * wo * cout * R * R * cin
    return dict(layer=name, count_per_backbone_pass=count, shape="%dx%dx%dx%d -> %d, %dx%d s%d%s" % (
        n, h, w, cin, cout, R, R, stride, " +res" if resid else ""), dvid_us=t_dvid, lib_us=t_lib,
        dvid_tflops=fl / t_dvid / 1e6, lib_tflops=fl / t_lib / 1e6, speedup=t_lib / t_dvid)


def gemm_case(name, count, m, k, n, relu=False):
    a = torch.randn(m, k, generator=g).half().to(dev)
    wt = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(dev)
    b = torch.randn(n, generator=g).to(dev)
    out = torch.empty(m, n, device=dev, dtype=torch.float16)
    t_dvid = graph_time(lambda: ops.gemm(a, wt, b, relu=relu, out=out))
    bh = b.half()

    def lib():
        y = F.linear(a, wt, bh)
        return F.relu_(y) if relu else y
    t_lib = graph_time(lib)
    fl = 2.0 * m * n * k
    return dict(layer=name, count_per_head_eval=count, shape="%dx%d -> %d" % (m, k, n), dvid_us=t_dvid, lib_us=t_lib,
                dvid_tflops=fl / t_dvid / 1e6, lib_tflops=fl / t_lib / 1e6, speedup=t_lib / t_dvid)


def gemm_partials_case(name, count, m, k, n, splits):
    """The call the model makes for the K-heavy / N=256 GEMMs: fp32 split-K partials (bias, residual and LayerNorm are
    applied by the row kernel that reduces them) vs the library's fp16 linear with bias."""
    a = torch.randn(m, k, generator=g).half().to(dev)
    wt = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(dev)
    b = torch.randn(n, generator=g).half().to(dev)
    out = torch.empty((max(1, splits), m, n), device=dev, dtype=torch.float32)
    t_dvid = graph_time(lambda: ops.gemm_partials(a, wt, splits, out=out))
    t_lib = graph_time(lambda: F.linear(a, wt, b))
    fl = 2.0 * m * n * k
    return dict(layer=name, count_per_head_eval=count, shape="%dx%d -> %d, split-K %d fp32 partials" % (m, k, n, splits),
                dvid_us=t_dvid, lib_us=t_lib, dvid_tflops=fl / t_dvid / 1e6, lib_tflops=fl / t_lib / 1e6,
                speedup=t_lib / t_dvid)


def main():
    B = 8
    rows = []
    # (name, count per backbone pass, n, h, w, cin, cout, R, stride, resid, relu)
    convs = [
        ("res2.0.conv1 / shortcut 64->64/256", 1, B, 152, 256, 64, 64, 1, 1, False, True),
        ("res2.0.shortcut 64->256", 1, B, 152, 256, 64, 256, 1, 1, False, False),
        ("res2.x.conv2 3x3 64->64", 3, B, 152, 256, 64, 64, 3, 1, False, True),
        ("res2.x.conv3 64->256 +res", 3, B, 152, 256, 64, 256, 1, 1, True, True),
        ("res2.x.conv1 256->64", 2, B, 152, 256, 256, 64, 1, 1, False, True),
        ("res3.0.conv1 256->128", 1, B, 152, 256, 256, 128, 1, 1, False, True),
        ("res3.0.conv2 3x3 s2 128->128", 1, B, 152, 256, 128, 128, 3, 2, False, True),
        ("res3.0.shortcut s2 256->512", 1, B, 152, 256, 256, 512, 1, 2, False, False),
        ("res3.x.conv3 128->512 +res", 4, B, 76, 128, 128, 512, 1, 1, True, True),
        ("res3.x.conv1 512->128", 3, B, 76, 128, 512, 128, 1, 1, False, True),
        ("res3.x.conv2 3x3 128->128", 3, B, 76, 128, 128, 128, 3, 1, False, True),
        ("res4.0.conv1 512->256", 1, B, 76, 128, 512, 256, 1, 1, False, True),
        ("res4.0.conv2 3x3 s2 256->256", 1, B, 76, 128, 256, 256, 3, 2, False, True),
        ("res4.0.shortcut s2 512->1024", 1, B, 76, 128, 512, 1024, 1, 2, False, False),
        ("res4.x.conv3 256->1024 +res", 23, B, 38, 64, 256, 1024, 1, 1, True, True),
        ("res4.x.conv1 1024->256", 22, B, 38, 64, 1024, 256, 1, 1, False, True),
        ("res4.x.conv2 3x3 256->256", 22, B, 38, 64, 256, 256, 3, 1, False, True),
        ("res5.0.conv1 1024->512", 1, B, 38, 64, 1024, 512, 1, 1, False, True),
        ("res5.0.conv2 3x3 s2 512->512", 1, B, 38, 64, 512, 512, 3, 2, False, True),
        ("res5.0.shortcut s2 1024->2048", 1, B, 38, 64, 1024, 2048, 1, 2, False, False),
        ("res5.x.conv3 512->2048 +res", 3, B, 19, 32, 512, 2048, 1, 1, True, True),
        ("res5.x.conv1 2048->512", 2, B, 19, 32, 2048, 512, 1, 1, False, True),
        ("res5.x.conv2 3x3 512->512", 2, B, 19, 32, 512, 512, 3, 1, False, True),
        ("fpn_lateral5 2048->256", 1, B, 19, 32, 2048, 256, 1, 1, False, False),
        ("fpn_lateral4 1024->256", 1, B, 38, 64, 1024, 256, 1, 1, False, False),
        ("fpn_lateral3 512->256", 1, B, 76, 128, 512, 256, 1, 1, False, False),
        ("fpn_output5 3x3 256->256", 1, B, 19, 32, 256, 256, 3, 1, False, False),
        ("fpn_output4 3x3 256->256", 1, B, 38, 64, 256, 256, 3, 1, False, False),
        ("fpn_output3 3x3 256->256", 1, B, 76, 128, 256, 256, 3, 1, False, False),
    ]
    for c in convs:
        r = conv_case(*c)
        rows.append(r)
        print("%-36s dvid %7.1f us %7.1f TF/s | lib %7.1f us %7.1f TF/s | x%.2f" % (
            r["layer"], r["dvid_us"], r["dvid_tflops"], r["lib_us"], r["lib_tflops"], r["speedup"]), flush=True)
    M = 2400
    gemms = [("self_attn.in_proj 256->768", 1, M, 256, 768, False), ("dynamic_layer 256->32768", 1, M, 256, 32768, False),
             ("out_layer 12544->256", 1, M, 12544, 256, False), ("linear1 256->2048 +ReLU", 1, M, 256, 2048, True),
             ("linear2 2048->256", 1, M, 2048, 256, False), ("out_proj / towers 256->256", 6, M, 256, 256, False),
             ("global_attention q 256->256 (cross)", 1, M, 256, 256, False)]
    grow = []
    for c in gemms:
        r = gemm_case(*c)
        grow.append(r)
        print("%-36s dvid %7.1f us %7.1f TF/s | lib %7.1f us %7.1f TF/s | x%.2f" % (
            r["layer"], r["dvid_us"], r["dvid_tflops"], r["lib_us"], r["lib_tflops"], r["speedup"]), flush=True)
    for c in [("out_layer 12544->256 (as called: split-K 7)", 1, M, 12544, 256, 7),
              ("linear2 2048->256 (as called: split-K 4)", 1, M, 2048, 256, 4),
              ("out_proj / cond 256->256 (as called: partials)", 3, M, 256, 256, 1)]:
        r = gemm_partials_case(*c)
        grow.append(r)
        print("%-36s dvid %7.1f us %7.1f TF/s | lib %7.1f us %7.1f TF/s | x%.2f" % (
            r["layer"], r["dvid_us"], r["dvid_tflops"], r["lib_us"], r["lib_tflops"], r["speedup"]), flush=True)
    tot_d = sum(r["dvid_us"] * r["count_per_backbone_pass"] for r in rows)
    tot_l = sum(r["lib_us"] * r["count_per_backbone_pass"] for r in rows)
    rec = {"gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "cudnn": torch.backends.cudnn.version(),
           "method": "CUDA graph of 10 back-to-back launches, median of 10 replays, warm L2; library arm = F.conv2d "
                     "channels_last fp16 + cudnn.benchmark (+ add / ReLU kernels) and F.linear",
           "backbone_convs": rows, "decoder_gemms": grow,
           "backbone_pass_sum_us": {"dvid": tot_d, "library": tot_l, "speedup": tot_l / tot_d},
           "slower_than_library": [r["layer"] for r in rows + grow if r["speedup"] < 1.0]}
    print("BACKBONE SUM dvid %.0f us  library %.0f us  x%.2f; slower than library: %s" % (
        tot_d, tot_l, tot_l / tot_d, rec["slower_than_library"]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "layer_vs_library.json"), "w") as f:
        json.dump(rec, f, indent=1)


if __name__ == "__main__":
    with torch.no_grad():
        main()

"""Plain (no residual) stride-1 convolutions of the R-101+FPN backbone at the headline shape (8 frames, 608x1024): time per launch of
the conv kernel; run it once per setting of DVID_CONV_PATCH / DVID_CONV_CTA2 (the switches are read once per process)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import ops  # noqa: E402

SHAPES = [("res2 3x3 64->64", 8, 152, 256, 64, 64, 3), ("res3 3x3 128->128", 8, 76, 128, 128, 128, 3),
          ("res4 3x3 256->256", 8, 38, 64, 256, 256, 3), ("res5 3x3 512->512", 8, 19, 32, 512, 512, 3),
          ("fpn_out3 256->256", 8, 76, 128, 256, 256, 3), ("fpn_out4 256->256", 8, 38, 64, 256, 256, 3),
          ("fpn_out5 256->256", 8, 19, 32, 256, 256, 3),
          ("res4 1x1 1024->256", 8, 38, 64, 1024, 256, 1), ("res4.0 1x1 512->256", 8, 76, 128, 512, 256, 1),
          ("res5 1x1 2048->512", 8, 19, 32, 2048, 512, 1), ("res5.0 1x1 1024->512", 8, 38, 64, 1024, 512, 1)]


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    res = {}
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    for name, n, h, w, cin, cout, R in SHAPES:
        x = torch.randn(n, h, w, cin, generator=g).half().to(dev)
        wt = (torch.randn(cout, R * R * cin, generator=g) / (R * R * cin) ** 0.5).half().to(dev)
        b = torch.zeros(cout, device=dev)
        out = torch.empty(n, h, w, cout, device=dev, dtype=torch.float16)
        for _ in range(3):
            ops.conv2d(x, wt, b, cout, R, R, 1, R // 2, True, out=out)
        # warm-L2 back-to-back launches (the pipeline's situation: the input was just written by the previous layer)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv2d(x, wt, b, cout, R, R, 1, R // 2, True, out=out)
        e1.record()
        torch.cuda.synchronize()
        warm = e0.elapsed_time(e1) * 1000 / 20
        ts = []
        for _ in range(10):
            flush.zero_()
            e0.record()
            ops.conv2d(x, wt, b, cout, R, R, 1, R // 2, True, out=out)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1000)
        ts.sort()
        fl = 2.0 * n * h * w * cout * R * R * cin
        res[name] = {"warm_us": round(warm, 2), "cold_us": round(ts[len(ts) // 2], 2),
                     "warm_tflops": round(fl / warm / 1e6, 1), "checksum": float(out.float().sum())}
    print("CONV_PATCH=%s CTA2=%s " % (os.environ.get("DVID_CONV_PATCH", "0"), os.environ.get("DVID_CONV_CTA2", "0")) + json.dumps(res))


if __name__ == "__main__":
    main()

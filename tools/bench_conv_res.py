"""1x1 convolutions with a same-resolution residual (bottleneck conv3) at the headline shape: time per launch, warm L2.
DVID_FORCE_BN selects the tile width (read once per process)."""
import json
import os
import sys

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import ops  # noqa: E402

SHAPES = [("res2 conv3 64->256", 8, 152, 256, 64, 256), ("res3 conv3 128->512", 8, 76, 128, 128, 512),
          ("res4 conv3 256->1024", 8, 38, 64, 256, 1024), ("res5 conv3 512->2048", 8, 19, 32, 512, 2048)]


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(0)
    res = {}
    for name, n, h, w, cin, cout in SHAPES:
        x = torch.randn(n, h, w, cin, generator=g).half().to(dev)
        wt = (torch.randn(cout, cin, generator=g) / cin ** 0.5).half().to(dev)
        b = torch.zeros(cout, device=dev)
        r = torch.randn(n, h, w, cout, generator=g).half().to(dev)
        out = torch.empty(n, h, w, cout, device=dev, dtype=torch.float16)
        for _ in range(3):
            ops.conv2d(x, wt, b, cout, 1, 1, 1, 0, True, resid=r, out=out)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.conv2d(x, wt, b, cout, 1, 1, 1, 0, True, resid=r, out=out)
        e1.record()
        torch.cuda.synchronize()
        res[name] = {"warm_us": round(e0.elapsed_time(e1) * 1000 / 20, 2), "checksum": float(out.float().sum())}
    print("FORCE_BN=%s " % os.environ.get("DVID_FORCE_BN", "0") + json.dumps(res))


if __name__ == "__main__":
    main()

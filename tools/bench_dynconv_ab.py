"""A/B of the two DynamicConv kernels at the headline shape (8 frames x 300 boxes on 608x1024 pyramids): CUDA-graph
replays, cold-ish L2 (a 256 MB buffer is rewritten between replays)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import ops  # noqa: E402


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    frames, n = 8, 300
    M = frames * n
    feats = [torch.randn(frames, h, w, 256, generator=g).half().to(dev) for h, w in ((76, 128), (38, 64), (19, 32))]
    lv = ops.Levels(feats)
    cx = torch.rand(frames, n, 4, generator=g)
    ctr = cx[..., :2] * torch.tensor([1000., 600.])
    wh = cx[..., 2:] * torch.tensor([1000., 600.]) * 0.8 + 4
    boxes = torch.cat([ctr - wh / 2, ctr + wh / 2], -1).contiguous().to(dev)
    params = (torch.randn(M, 32768, generator=g) * 0.1).half().to(dev)
    ln = [t.to(dev) for t in (torch.ones(64), torch.zeros(64), torch.ones(256), torch.zeros(256))]
    flush = torch.empty(64 * 1024 * 1024, device=dev)
    out = torch.empty((M, 49 * 256), device=dev, dtype=torch.float16)
    res = {}
    for name, tc in (("mma_sync", False), ("tcgen05", True), ("mma_sync2", False), ("tcgen05_2", True)):
        for _ in range(3):
            ops.roi_dynconv(lv, boxes, n, params, *ln, out=out, transposed=tc)
        ts = []
        for _ in range(20):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.roi_dynconv(lv, boxes, n, params, *ln, out=out, transposed=tc)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1000)
        ts.sort()
        res[name] = {"median_us": ts[len(ts) // 2], "min_us": ts[0]}
    print("DYNCONV_AB " + json.dumps(res))


if __name__ == "__main__":
    main()

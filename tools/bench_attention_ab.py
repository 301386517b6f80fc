"""A/B of the two attention kernels at the decoder's shapes (self: 8 frames x 300 boxes; cross: 2400 x 900)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import ops  # noqa: E402


def timeit(fn, n=20, reps=20):
    """device time per call: `reps` back-to-back calls captured in a CUDA graph, median over n replays"""
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1000 / reps)
    ts.sort()
    return ts[len(ts) // 2]


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    B, N = 8, 300
    qkv = torch.randn(B * N, 768, generator=g).half().to(dev)
    ctx = torch.empty(B * N, 256, dtype=torch.float16, device=dev)
    q = torch.randn(2400, 256, generator=g).half().to(dev)
    kv = torch.randn(900, 512, generator=g).half().to(dev)
    res = {}
    for tc in (False, True, False, True):
        name = "tcgen05" if tc else "mma_sync"
        res.setdefault(name, {})
        res[name].setdefault("self_us", []).append(timeit(lambda: ops.attention(
            qkv, qkv[:, 256:], qkv[:, 512:], ctx, B, 8, N, N, 768, 768, 768, 256, N * 768, N * 768, N * 768, N * 256, tc=tc)))
        res[name].setdefault("cross_us", []).append(timeit(lambda: ops.attention(
            q, kv, kv[:, 256:], ctx, 1, 8, 2400, 900, 256, 512, 512, 256, 0, 0, 0, 0, tc=tc)))
    print("ATTN_AB " + json.dumps(res))


if __name__ == "__main__":
    main()

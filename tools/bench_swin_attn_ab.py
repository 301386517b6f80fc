"""A/B of the two Swin window-attention kernels at the four Swin-B stage shapes of a 4-frame 608x1024 batch."""
import json
import sys

import torch

sys.path.insert(0, ".")
from diffusionvid_b200 import ops  # noqa: E402
from tools.bench_attention_ab import timeit  # noqa: E402


def main():
    dev = "cuda"
    g = torch.Generator().manual_seed(1)
    res = {}
    for name, (Ht, Wt, C, nh) in {"stage0": (152, 256, 128, 4), "stage1": (76, 128, 256, 8),
                                  "stage2": (38, 64, 512, 16), "stage3": (19, 32, 1024, 32)}.items():
        B = 4
        nw = B * ((Ht + 6) // 7) * ((Wt + 6) // 7)
        qkv = (0.7 * torch.randn(nw * 49, 3 * C, generator=g)).half().to(dev)
        bias = (0.5 * torch.randn(nh, 49, 49, generator=g)).to(dev)
        res[name] = {"windows": nw, "heads": nh}
        for tc in (False, True):
            res[name]["tcgen05_us" if tc else "mma_sync_us"] = timeit(
                lambda: ops.swin_window_attention(qkv, bias, B, Ht, Wt, C, nh, 3, tc=tc), n=10, reps=5)
    print("SWIN_ATTN_AB " + json.dumps(res))


if __name__ == "__main__":
    main()

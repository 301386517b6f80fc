import sys
import torch
sys.path.insert(0, ".")
from diffusionvid_b200 import ops
g = torch.Generator().manual_seed(1)
B, Ht, Wt, C, nh = 4, 38, 64, 512, 16          # Swin-B stage 2 (18 of the 24 blocks), 4 frames of 608x1024
nw = B * ((Ht + 6) // 7) * ((Wt + 6) // 7)
qkv = (0.7 * torch.randn(nw * 49, 3 * C, generator=g)).half().cuda()
bias = (0.5 * torch.randn(nh, 49, 49, generator=g)).cuda()
for tc in (False, True):
    for _ in range(3):
        ops.swin_window_attention(qkv, bias, B, Ht, Wt, C, nh, 3, tc=tc)
torch.cuda.synchronize()

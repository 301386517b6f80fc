import sys
import torch
sys.path.insert(0, ".")
from diffusionvid_b200 import ops
g = torch.Generator().manual_seed(1)
q = torch.randn(2400, 256, generator=g).half().cuda()
kv = torch.randn(900, 512, generator=g).half().cuda()
ctx = torch.empty(2400, 256, dtype=torch.float16, device="cuda")
for _ in range(3):
    ops.attention(q, kv, kv[:, 256:], ctx, 1, 8, 2400, 900, 256, 512, 512, 256, 0, 0, 0, 0, tc=True)
torch.cuda.synchronize()

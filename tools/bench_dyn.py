import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffusionvid_b200 import ops
dev = torch.device("cuda")
m, k, n = 2400, 256, 32768
a = torch.randn(m, k).half().to(dev); w = (torch.randn(n, k) / 16).half().to(dev); b = torch.randn(n).to(dev)
outs = [torch.empty(m, n, device=dev, dtype=torch.float16) for _ in range(4)]
res = {}
for relu in (False, True):
    for _ in range(5): ops.gemm(a, w, b, relu=relu, out=outs[0])
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for i in range(40): ops.gemm(a, w, b, relu=relu, out=outs[i & 3])
    e.record(); torch.cuda.synchronize()
    res[relu] = s.elapsed_time(e) * 1e3 / 40
print("DBG", os.environ.get("DVID_DBG", "0"), "BSTAT", os.environ.get("DVID_BSTAT", "1"), "dyn us  norelu %.1f  relu %.1f" % (res[False], res[True]))

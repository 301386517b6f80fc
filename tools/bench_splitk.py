"""split-K factor sweep for the decoder's K-heavy GEMMs (fp32 partials, as the model calls them) vs cuBLAS."""
import sys
import torch
import torch.nn.functional as F
sys.path.insert(0, ".")
from diffusionvid_b200 import ops
from tools.layer_vs_library import graph_time

g = torch.Generator().manual_seed(0)
dev = "cuda"
for name, m, k, n, splits in (("out_layer", 2400, 12544, 256, (6, 7)), ("linear2", 2400, 2048, 256, (2, 3, 4, 5, 6, 7)),
                              ("out_layer M=1200", 1200, 12544, 256, (7, 14)), ("linear2 M=1200", 1200, 2048, 256, (3, 4, 6, 7)), ("linear2 M=9600", 9600, 2048, 256, (1, 2, 3, 4)), ("out_layer M=9600", 9600, 12544, 256, (2, 4, 7))):
    a = torch.randn(m, k, generator=g).half().to(dev)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(dev)
    b = torch.randn(n, generator=g).half().to(dev)
    t_lib = graph_time(lambda: F.linear(a, w, b))
    res = []
    for s in splits:
        out = torch.empty((s, m, n), device=dev, dtype=torch.float32)
        res.append((s, graph_time(lambda: ops.gemm_partials(a, w, s, out=out))))
    print(name, "cuBLAS %.1f us |" % t_lib, " ".join("s%d: %.1f" % r for r in res), flush=True)

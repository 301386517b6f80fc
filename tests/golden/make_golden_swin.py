#!/usr/bin/env python
"""Golden Swin features from the REFERENCE'S OWN mega_core/modeling/backbone/swintransformer.py, loaded by path,
unmodified, with its third-party imports satisfied by stand-ins defined here:
  timm.models.layers.{DropPath,to_2tuple,trunc_normal_} -> identity module / tuple helper / torch trunc_normal_
  fvcore.nn.weight_init, detectron2 {ShapeSpec, Backbone, BACKBONE_REGISTRY, FPN, LastLevelMaxPool} -> nn.Module / inert
Runs SwinTransformer(embed 128 (Swin-B widths), depths 2-2-2-2, heads 4-8-16-32, window 7, out_indices 1-3) on a seeded 96x160 batch with
the synthetic state dict (strict load: pins the key names) and stores swin1..3.  Writes tests/golden/ref_swin_small.pt."""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DVID_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from diffusionvid_b200 import synth  # noqa: E402


def mod(name, **attrs):
    m = types.ModuleType(name)
    m.__path__ = []
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Reg:
    def register(self):
        return lambda f: f


mod("timm"); mod("timm.models")
mod("timm.models.layers", DropPath=torch.nn.Identity, to_2tuple=lambda x: (x, x) if not isinstance(x, tuple) else x,
    trunc_normal_=torch.nn.init.trunc_normal_)
mod("fvcore"); mod("fvcore.nn"); mod("fvcore.nn.weight_init")
sys.modules["fvcore.nn"].weight_init = sys.modules["fvcore.nn.weight_init"]
mod("detectron2"); mod("detectron2.layers", ShapeSpec=object); mod("detectron2.modeling")
mod("detectron2.modeling.backbone")
mod("detectron2.modeling.backbone.backbone", Backbone=torch.nn.Module)
mod("detectron2.modeling.backbone.build", BACKBONE_REGISTRY=_Reg())
mod("detectron2.modeling.backbone.fpn", FPN=object, LastLevelMaxPool=object)

spec = importlib.util.spec_from_file_location("ref_swin", os.path.join(REF, "mega_core/modeling/backbone/swintransformer.py"))
ref = importlib.util.module_from_spec(spec)
spec.loader.exec_module(ref)

CFG = dict(embed=128, depths=(2, 2, 2, 2), heads=(4, 8, 16, 32))
sd = synth.make_state_dict(seed=91, swin=CFG)
net = ref.SwinTransformer(embed_dim=CFG["embed"], depths=list(CFG["depths"]), num_heads=list(CFG["heads"]),
                          window_size=7, drop_path_rate=0.3, out_indices=(1, 2, 3))
body = {k[len("backbone.bottom_up."):]: v for k, v in sd.items() if k.startswith("backbone.bottom_up.")}
missing, unexpected = net.load_state_dict(body, strict=False)
assert not unexpected, unexpected
assert all(k.endswith("relative_position_index") for k in missing), missing     # buffers, not weights
net.eval()
g = torch.Generator().manual_seed(92)
x = torch.randn(2, 3, 96, 160, generator=g)
with torch.no_grad():
    out = net(x)
path = os.path.join(HERE, "ref_swin_small.pt")
torch.save(dict(meta=dict(cfg=CFG, weight_seed=91, input_seed=92, shape=(2, 3, 96, 160)),
                out={k: v.clone() for k, v in out.items()}), path)
print("wrote", path, os.path.getsize(path), {k: tuple(v.shape) for k, v in out.items()})

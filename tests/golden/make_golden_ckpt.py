#!/usr/bin/env python
"""Generate tests/golden/ckpt_key_mapping.json by RUNNING the reference's own state-dict loader on this container's CPU.

Loads /root/reference/mega_core/utils/model_serialization.py by path, unmodified.  Its two imports are satisfied with
stand-ins: ``mega_core.utils.imports`` (inert, never called) and ``mega_core.modeling.detector.diffusion_det.DiffusionDet``
(bound to this repo's module class so that the reference's ``isinstance(model, DiffusionDet)`` branch - the
DiffusionDet -> DiffusionVID head renaming - is taken for our model, whose parameter names are the reference's).

For each scenario a synthetic checkpoint is built whose tensors carry their own index as value; after the reference's
``load_state_dict(model, ckpt)`` the value found under every model key tells which loaded key it took (or that it was
left untouched).  Only key names and shapes are stored.

Run:  python tests/golden/make_golden_ckpt.py     (needs /root/reference)
"""
import importlib.util
import io
import json
import os
import sys
import types
from contextlib import redirect_stdout

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DVID_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from diffusionvid_b200 import model as pm  # noqa: E402

HP = dict(num_classes=30, num_proposals=20, sample_step=2, device="cpu", blocks=(1, 1, 1, 1))


def scenarios(model_keys):
    """name -> list of (loaded key, model key whose shape it has)."""
    out = {}
    out["ddp_prefix"] = [("module." + k, k) for k in model_keys]
    det = []
    for k in model_keys:
        if "global_attention" in k or ".c_mlp." in k:
            continue                                     # DiffusionDet has neither
        if "head_series_cond.0." in k:
            for i in (3, 4, 5):
                det.append((k.replace("head_series_cond.0.", "head_series.%d." % i), k))
        else:
            det.append((k, k))
    out["diffusiondet_six_heads"] = det
    out["local_heads"] = [(k.replace("head_series_cond", "head_series_local"), k) for k in model_keys]
    out["backbone_only_suffix"] = [(k[len("backbone.bottom_up."):], k) for k in model_keys
                                   if k.startswith("backbone.bottom_up.")]
    out["wrapped_ddp_partial"] = [("module." + k, k) for k in model_keys if "backbone" not in k]
    out["mixed_prefix_not_stripped"] = [(("module." + k) if i % 2 else k, k) for i, k in enumerate(model_keys)
                                        if "head" in k]
    return out


def main():
    for name in ("mega_core", "mega_core.utils", "mega_core.utils.imports", "mega_core.modeling",
                 "mega_core.modeling.detector", "mega_core.modeling.detector.diffusion_det"):
        mod = types.ModuleType(name)
        mod.__path__ = []
        sys.modules[name] = mod
    sys.modules["mega_core.utils.imports"].import_file = lambda *a, **k: None
    sys.modules["mega_core.modeling.detector.diffusion_det"].DiffusionDet = pm.DiffusionDet
    spec = importlib.util.spec_from_file_location(
        "mega_core.utils.model_serialization", os.path.join(REF, "mega_core/utils/model_serialization.py"))
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)

    import bench
    hp = dict(bench.HP_BASE, **HP)
    fixture = {"hp": {k: (list(v) if isinstance(v, tuple) else v) for k, v in HP.items()}, "scenarios": {}}
    for name, pairs in scenarios(list(pm.DiffusionDet(hp).state_dict().keys())).items():
        model = pm.DiffusionDet(hp)
        shapes = {k: tuple(v.shape) for k, v in model.state_dict().items()}
        with torch.no_grad():
            for v in model.state_dict().values():
                v.fill_(-1)
        loaded = {lk: torch.full(shapes[mk], float(i)) for i, (lk, mk) in enumerate(pairs)}
        with redirect_stdout(io.StringIO()):
            ref.load_state_dict(model, dict(loaded), flownet=None)       # tools/test_net.py:104 passes flownet=None
        mapping = {}
        for k, v in model.state_dict().items():
            val = float(v.flatten()[0]) if v.numel() else -1.0
            mapping[k] = None if val < 0 else pairs[int(val)][0]
        fixture["scenarios"][name] = {"loaded": [[lk, list(shapes[mk])] for lk, mk in pairs], "mapping": mapping}
        print(name, "loaded", len(pairs), "matched", sum(v is not None for v in mapping.values()), "of", len(mapping))
    with open(os.path.join(HERE, "ckpt_key_mapping.json"), "w") as f:
        json.dump(fixture, f, indent=0, sort_keys=True)


if __name__ == "__main__":
    main()

"""Checkpoint importer against key mappings produced by the reference's own loader
(tests/golden/make_golden_ckpt.py ran mega_core/utils/model_serialization.py:12-156 on this repo's module)."""
import json
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from diffusionvid_b200 import checkpoint as ck  # noqa: E402
from diffusionvid_b200 import model as pm  # noqa: E402

with open(os.path.join(ROOT, "tests", "golden", "ckpt_key_mapping.json")) as f:
    FIX = json.load(f)


def _model():
    hp = dict(bench.HP_BASE, **{k: (tuple(v) if isinstance(v, list) else v) for k, v in FIX["hp"].items()})
    m = pm.DiffusionDet(hp)
    with torch.no_grad():
        for v in m.state_dict().values():
            v.fill_(-1)
    return m


@pytest.mark.parametrize("name", sorted(FIX["scenarios"]))
def test_key_mapping_matches_reference_loader(name):
    sc = FIX["scenarios"][name]
    m = _model()
    loaded = {k: torch.full(tuple(shape), float(i)) for i, (k, shape) in enumerate(sc["loaded"])}
    missing = ck.load_state_dict(m, loaded)
    got = {}
    for k, v in m.state_dict().items():
        val = float(v.flatten()[0])
        got[k] = None if val < 0 else sc["loaded"][int(val)][0]
    assert got == sc["mapping"]
    assert sorted(missing) == sorted(k for k, v in sc["mapping"].items() if v is None)


def test_load_checkpoint_file_roundtrip(tmp_path):
    """A DDP-saved training checkpoint ({'model': module.*, 'optimizer': ...}) loads from disk and resets the packed
    weights so the next forward repacks (checkpoint.py:52-82,113-114)."""
    src = pm.DiffusionDet(dict(bench.HP_BASE, **{k: (tuple(v) if isinstance(v, list) else v)
                                                 for k, v in FIX["hp"].items()}))
    g = torch.Generator().manual_seed(3)
    with torch.no_grad():
        for v in src.state_dict().values():
            if v.is_floating_point():
                v.copy_(torch.randn(v.shape, generator=g))
    path = str(tmp_path / "model_final.pth")
    torch.save({"model": {"module." + k: v for k, v in src.state_dict().items()}, "optimizer": {}, "iteration": 7},
               path)
    dst = _model()
    dst._pk = object()
    assert ck.load_checkpoint(dst, path) == []
    assert dst._pk is None
    for (k, a), b in zip(src.state_dict().items(), dst.state_dict().values()):
        assert torch.equal(a, b), k

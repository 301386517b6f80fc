"""The product model behind the REFERENCE'S OWN caller code and containers, on the CPU of the build box.

The reference's package cannot be imported whole, so - exactly like tests/golden/make_golden.py - the
files that form the boundary are loaded UNMODIFIED by path under their own module names:

    mega_core/structures/{bounding_box,image_list,boxlist_ops}.py    BoxList / ImageList / to_image_list / cat_boxlist
    mega_core/engine/inference.py                                    compute_on_dataset (the loop of tools/test_net.py)

everything else they import resolves to inert placeholders.  The tests check that
  * the reference's ImageList objects are accepted by DiffusionDet.forward (engine/inference.py:35-40 hands them over),
  * the reference's `compute_on_dataset` drives the product model and collects its results unchanged,
  * the collected results pickle under the class path `mega_core.structures.bounding_box.BoxList` and unpickle as the
    reference's own class (what tools/test_prediction.py / vid_eval.py:14-25 load from predictions.pth),
  * the reference's real yaml files merge through this package's config reader into the right hot-path parameters.
These tests read /root/reference and are skipped where it does not exist (the GPU box); the product's compute is the
test-only CPU op shim (tests/cpu_ops_shim.py), as in tests/test_host_logic.py.
"""
import importlib.util
import io
import os
import pickle
import sys
import types

import pytest
import torch

from diffusionvid_b200 import compat, config, model as pm, structures, synth
from oracle import model as om
from tests import cpu_ops_shim
from tests.test_host_logic import SMALL

REF = os.environ.get("DVID_REFERENCE", "/root/reference")
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "mega_core")), reason="reference tree not present")

_ROOTS = ("mega_core", "apex", "seq_nms")


class _Inert:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        return _Inert()


class _InertModule(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Inert


class _Finder:
    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in _ROOTS:
            return importlib.util.spec_from_loader(name, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _InertModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _load(name, relpath):
    spec = importlib.util.spec_from_file_location(name, os.path.join(REF, relpath))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture
def reference():
    """The reference's boundary modules, loaded by path; restored to the alias modules afterwards."""
    compat.remove_mega_core_alias()
    finder = _Finder()
    sys.meta_path.insert(0, finder)
    try:
        ns = types.SimpleNamespace(
            bounding_box=_load("mega_core.structures.bounding_box", "mega_core/structures/bounding_box.py"),
            image_list=_load("mega_core.structures.image_list", "mega_core/structures/image_list.py"),
            boxlist_ops=_load("mega_core.structures.boxlist_ops", "mega_core/structures/boxlist_ops.py"),
            inference=_load("mega_core.engine.inference", "mega_core/engine/inference.py"))
        yield ns
    finally:
        sys.meta_path.remove(finder)
        for k in [k for k in sys.modules if k.split(".")[0] in _ROOTS]:
            del sys.modules[k]
        compat.install_mega_core_alias()


def _product(T=1):
    hp = dict(SMALL, sample_step=T)
    m = pm.DiffusionDet(hp)
    m.load_state_dict(synth.make_state_dict(seed=11, blocks=hp["blocks"]), strict=False)
    m.noise = om.NoiseSource(3, hp["num_proposals"])
    return m


def _loader(ImageList, frames, gidx, h, w):
    """(images, targets, image_ids) batches as VIDMEGADataset._get_test + BatchCollator produce them
    (data/datasets/vid_mega.py:164-250): every image an ImageList of the REFERENCE's class."""
    L = frames.shape[0]
    for s in synth.clip_samples(frames, gidx, h, w):
        images = dict(cur=ImageList(s["cur"], [(h, w)]), ref_l=[ImageList(t, [(h, w)]) for t in s["ref_l"]],
                      ref_g=[ImageList(t, [(h, w)]) for t in s["ref_g"]], frame_id=s["frame_id"], start_id=0,
                      end_id=L - 1, seg_len=L, frame_category=s["frame_category"], video_id=0)
        ids = list(range(s["frame_id"], min(L, s["frame_id"] + 8)))
        yield images, None, [ids]


def test_reference_engine_drives_the_product_model(reference, monkeypatch):
    monkeypatch.setattr(pm, "ops", cpu_ops_shim)
    h, w, L = 64, 96, 11
    frames = synth.make_clip(L, h, w, seed=5)
    gidx = [9, 2, 5]
    RefImageList, RefBoxList = reference.image_list.ImageList, reference.bounding_box.BoxList
    assert RefImageList is not structures.ImageList and RefBoxList is not structures.BoxList

    # the reference's own loop (engine/inference.py:22-93), "diffusion" branch
    res = reference.inference.compute_on_dataset(_product(), list(_loader(RefImageList, frames, gidx, h, w)),
                                                 torch.device("cpu"), False, "diffusion", timer=None)
    assert sorted(res) == list(range(L))

    # same clip through the product's own containers: identical detections
    m = _product()
    direct = []
    for images, _, _ in _loader(structures.ImageList, frames, gidx, h, w):
        direct += m(images)
    assert len(direct) == L
    for i, d in enumerate(direct):
        assert torch.equal(res[i].bbox, d.bbox) and torch.equal(res[i].get_field("scores"), d.get_field("scores"))
        assert torch.equal(res[i].get_field("labels"), d.get_field("labels")) and res[i].size == (w, h)

    # predictions.pth (engine/inference.py:165-168): the stream names the reference's class and nothing of this package
    preds = [res[i] for i in range(L)]
    buf = io.BytesIO()
    torch.save(preds, buf)
    raw = buf.getvalue()
    assert b"mega_core.structures.bounding_box" in raw and b"diffusionvid_b200" not in raw
    back = torch.load(io.BytesIO(raw), weights_only=False)
    for b, p in zip(back, preds):
        assert type(b) is RefBoxList                                   # the reference's own class, its own methods
        assert torch.equal(b.bbox, p.bbox) and b.size == p.size and b.mode == "xyxy"
        assert b.fields() == ["scores", "labels"] and len(b) == len(p)
        assert torch.equal(b.get_field("labels"), p.get_field("labels"))
        assert torch.equal(b.clip_to_image(remove_empty=False).bbox, p.bbox)     # already clipped the legacy way
    # ... and the reference's cat_boxlist accepts what was loaded
    assert len(reference.boxlist_ops.cat_boxlist(back[:2])) == len(back[0]) + len(back[1])


def test_reference_to_image_list_objects_pass_through(reference):
    il = reference.image_list.to_image_list([torch.rand(3, 30, 50), torch.rand(3, 28, 60)], 32)
    out = structures.to_image_list(il)
    assert out is il and out.tensors.shape == (2, 3, 32, 64)


def test_alias_modules_serve_hosts_without_the_reference():
    """No reference checkout on the path: `mega_core.*` names resolve to this package and predictions round-trip."""
    mods = compat.install_mega_core_alias()
    assert "mega_core.structures.bounding_box" in mods
    from mega_core.structures.bounding_box import BoxList as AliasBoxList      # noqa: E402
    from mega_core.modeling.detector import build_detection_model             # noqa: E402,F401
    from mega_core.structures.image_list import to_image_list                 # noqa: E402
    assert issubclass(AliasBoxList, structures.BoxList) and to_image_list is structures.to_image_list
    b = structures.BoxList(torch.tensor([[1., 2., 3., 4.]]), (10, 20))
    b.add_field("scores", torch.tensor([0.5]))
    raw = pickle.dumps([b])
    assert b"mega_core.structures.bounding_box" in raw and b"diffusionvid_b200" not in raw
    back = pickle.loads(raw)[0]
    assert type(back) is AliasBoxList and torch.equal(back.bbox, b.bbox) and back.get_field("scores").item() == 0.5


@pytest.mark.parametrize("model_yaml,expect", [
    ("vid_R_101_DiffusionVID.yaml", dict(num_proposals=300, sample_step=1, num_heads=3, num_heads_local=1,
                                         blocks=(3, 4, 23, 3), mem_size=900, global_enable=True, infer_batch=8,
                                         all_frame_interval=8, key_frame_location=0, swin=None)),
    ("vid_R_101_DiffusionDET.yaml", dict(num_heads_local=0, blocks=(3, 4, 23, 3), swin=None)),
    ("vid_Swin_B_DiffusionVID.yaml", dict(num_proposals=300, infer_batch=4, all_frame_interval=4, global_enable=True)),
])
def test_reference_yaml_files_merge_into_hot_path_params(model_yaml, expect):
    """tools/test_net.py:77-83: BASE_RCNN_{N}gpu.yaml <- add_diffusiondet_config <- model yaml <- opts."""
    cfg = config.get_default_cfg()
    cfg.merge_from_file(os.path.join(REF, "configs", "BASE_RCNN_1gpu.yaml"))
    cfg.merge_from_file(os.path.join(REF, "configs", model_yaml))
    cfg.merge_from_list(["DTYPE", "float16", "MODEL.DiffusionDet.SAMPLE_STEP", str(expect.get("sample_step", 1))])
    cfg.freeze()
    hp = config.hot_path_params(cfg)
    for k, v in expect.items():
        if k == "swin":
            assert hp.get("swin") is None
        else:
            assert hp[k] == v, (k, hp[k], v)
    assert hp["num_classes"] == 30 and hp["hidden"] == 256 and hp["topk"] == (75, 25)
    if "Swin" in model_yaml:
        sw = hp["swin"]
        assert sw["embed"] == 128 and tuple(sw["depths"]) == (2, 2, 18, 2) and tuple(sw["heads"]) == (4, 8, 16, 32)

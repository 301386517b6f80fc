"""`mega_core` look-alike modules for hosts that do not have the reference's package on the path.

SURVEY.md 8b: the drop-in boundary is `mega_core.modeling.detector.build_detection_model` plus the containers
`mega_core.structures.{bounding_box.BoxList, image_list.ImageList/to_image_list, boxlist_ops.cat_boxlist}` and the native
module `mega_core._C`; `predictions.pth` is a pickle that names `mega_core.structures.bounding_box.BoxList`
(mega_core/engine/inference.py:165-168, read back by tools/test_prediction.py and vid_eval.py:14-25).

`install_mega_core_alias()` registers thin alias modules under those names in `sys.modules` - ONLY when no real
`mega_core` package can be found (a checkout of the reference always wins; then BoxLists pickle as the reference's own
class, see structures._pickle_class).  With the alias in place

    torch.load("predictions.pth")                       works without the reference installed, and
    torch.save(list_of_boxlists, "predictions.pth")     writes the reference's class path,

so the artefact travels in both directions between this package and the reference's tools.
"""
import importlib.util
import sys
import types

_ALIAS_FLAG = "__dvid_b200_alias__"


def _real_mega_core_present():
    m = sys.modules.get("mega_core")
    if m is not None:
        return not getattr(m, _ALIAS_FLAG, False)
    try:
        return importlib.util.find_spec("mega_core") is not None
    except (ImportError, ValueError):
        return False


def _module(name, package=False, **attrs):
    m = types.ModuleType(name)
    m.__dict__[_ALIAS_FLAG] = True
    if package:
        m.__path__ = []
    m.__dict__.update(attrs)
    sys.modules[name] = m
    parent, _, leaf = name.rpartition(".")
    if parent and parent in sys.modules:
        setattr(sys.modules[parent], leaf, m)
    return m


def install_mega_core_alias(force=False):
    """Returns {module name: module} of the alias modules now registered ({} when the real package is present and
    `force` is False).  Idempotent."""
    if _real_mega_core_present() and not force:
        return {}
    if getattr(sys.modules.get("mega_core"), _ALIAS_FLAG, False):
        return {k: v for k, v in sys.modules.items() if k.split(".")[0] == "mega_core"}
    from . import structures, _C_shim
    from .config import add_diffusiondet_config, get_default_cfg

    # a distinct class object whose __module__/__qualname__ are the reference's: pickle stores classes by that path
    BoxList = type("BoxList", (structures.BoxList,), {"__module__": "mega_core.structures.bounding_box",
                                                     "__doc__": structures.BoxList.__doc__})

    def build_detection_model(cfg):
        from . import build_detection_model as _b
        return _b(cfg)

    def _diffusion_det(*a, **k):
        from .model import DiffusionDet
        return DiffusionDet(*a, **k)

    _module("mega_core", package=True)
    _module("mega_core.structures", package=True)
    _module("mega_core.structures.bounding_box", BoxList=BoxList, FLIP_LEFT_RIGHT=0, FLIP_TOP_BOTTOM=1)
    _module("mega_core.structures.image_list", ImageList=structures.ImageList, to_image_list=structures.to_image_list)
    _module("mega_core.structures.boxlist_ops", cat_boxlist=structures.cat_boxlist)
    _module("mega_core.modeling", package=True)
    _module("mega_core.modeling.detector", package=True, build_detection_model=build_detection_model)
    _module("mega_core.modeling.detector.detectors", build_detection_model=build_detection_model)
    _module("mega_core.modeling.detector.diffusion_det", DiffusionDet=_diffusion_det,
            add_diffusiondet_config=add_diffusiondet_config)
    _module("mega_core.config", package=True, cfg=get_default_cfg())
    sys.modules["mega_core._C"] = _C_shim
    sys.modules["mega_core"]._C = _C_shim
    return {k: v for k, v in sys.modules.items() if k.split(".")[0] == "mega_core"}


def remove_mega_core_alias():
    """Drops the alias modules again (tests; or before putting a real checkout on the path)."""
    if not getattr(sys.modules.get("mega_core"), _ALIAS_FLAG, False):
        return
    for k in [k for k in sys.modules if k.split(".")[0] == "mega_core"]:
        del sys.modules[k]

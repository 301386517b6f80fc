"""diffusionvid_b200 — B200-native (sm_100a) implementation of the DiffusionVID inference hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute is hand-written CUDA behind the
C ABI declared in include/dvid_b200.h (libdvid_b200.so, loaded by diffusionvid_b200._lib). There is no CPU fallback:
using any operator without the built extension raises.

Public surface (mirrors the reference's mega_core boundary):
  build_detection_model(cfg) -> DiffusionDet     mega_core/modeling/detector/detectors.py:11-22
  DiffusionDet.forward(images, targets=None)     mega_core/modeling/detector/diffusion_det.py:306-336
  BoxList / ImageList / to_image_list            mega_core/structures/{bounding_box,image_list}.py
  add_diffusiondet_config / get_default_cfg      mega_core/modeling/detector/diffusion_det.py:74-179
  _C                                             mega_core/_C (csrc/vision.cpp:10-27), see diffusionvid_b200/_C_shim.py
"""
__version__ = "0.2.0"

from .config import CfgNode, add_diffusiondet_config, get_default_cfg  # noqa: E402,F401
from .structures import BoxList, ImageList, cat_boxlist, to_image_list  # noqa: E402,F401

_DETECTION_META_ARCHITECTURES = {}

# Without a checkout of the reference on the path, `mega_core.structures.bounding_box.BoxList` & co. resolve to this
# package (diffusionvid_b200/compat.py), so predictions.pth files carry - and load under - the reference's class path.
from . import compat as _compat  # noqa: E402
_compat.install_mega_core_alias()


def build_detection_model(cfg):
    """Registry lookup by cfg.MODEL.META_ARCHITECTURE, like mega_core.modeling.detector.build_detection_model."""
    from .model import DiffusionDet
    _DETECTION_META_ARCHITECTURES.setdefault("DiffusionDet", DiffusionDet)
    name = cfg.MODEL.META_ARCHITECTURE if hasattr(cfg, "MODEL") else "DiffusionDet"
    return _DETECTION_META_ARCHITECTURES[name](cfg)

"""Result / input containers with the reference's public API, so callers such as mega_core/engine/inference.py and the
VID evaluator keep working unchanged:

  BoxList   - mega_core/structures/bounding_box.py:9-255 (bbox fp32 Nx4, size=(W,H), mode, extra_fields; the legacy
              TO_REMOVE=1 pixel convention in clip_to_image/area/convert)
  ImageList - mega_core/structures/image_list.py:7-27, to_image_list :29-72 (zero-pad to a common, divisible size)
  cat_boxlist - mega_core/structures/boxlist_ops.py:103-133

Only the members the DiffusionVID inference path and its callers touch are provided.

Pickle contract (SURVEY.md 8b "output artefact"): `predictions.pth` is `torch.save(list[BoxList])` and its consumers
(`tools/test_prediction.py`, `vid_eval.py:14-25`) unpickle `mega_core.structures.bounding_box.BoxList`.  BoxLists made
here therefore pickle under THAT class path (`BoxList.__reduce_ex__`): when the reference's package is importable its own
class is named in the stream (the reference's class has no `__setstate__`, so the state dict - bbox / size / mode /
extra_fields - lands in its `__dict__`), otherwise `diffusionvid_b200.compat` registers alias modules under the
`mega_core.*` names that resolve to the classes of this file.
"""
import copyreg
import importlib
import math
import sys

import torch

_REF_BOXLIST_MODULE = "mega_core.structures.bounding_box"


def _pickle_class():
    """The class object that pickled BoxLists name: `mega_core.structures.bounding_box.BoxList` (the reference's own
    class when that package is importable, else this package's alias of it), falling back to this module's class."""
    mod = sys.modules.get(_REF_BOXLIST_MODULE)
    if mod is None:
        try:
            mod = importlib.import_module(_REF_BOXLIST_MODULE)
        except Exception:
            from . import compat
            mod = compat.install_mega_core_alias().get(_REF_BOXLIST_MODULE)
    cls = getattr(mod, "BoxList", None) if mod is not None else None
    return cls if isinstance(cls, type) else BoxList


class BoxList(object):
    def __init__(self, bbox, image_size, mode="xyxy"):
        device = bbox.device if isinstance(bbox, torch.Tensor) else torch.device("cpu")
        bbox = torch.as_tensor(bbox, dtype=torch.float32, device=device)
        if bbox.ndimension() != 2:
            raise ValueError("bbox should have 2 dimensions, got {}".format(bbox.ndimension()))
        if bbox.size(-1) != 4:
            raise ValueError("last dimension of bbox should have a size of 4, got {}".format(bbox.size(-1)))
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        self._bbox = bbox
        self._fields = {}
        self._pending = None
        self.size = image_size   # (image_width, image_height)
        self.mode = mode

    # ---- deferred results
    # The model's `host_results` mode returns BoxLists whose tensors are still in flight (an asynchronous device->host
    # copy of the key batch).  `resolver()` waits for that copy and returns (bbox, {field: tensor}); it runs on the first
    # access to the boxes, a field or the length, so a caller that only stores the result - like the reference's loop,
    # `output = [o.to(cpu_device) for o in output]; results_dict.update(...)` (mega_core/engine/inference.py:75-78) - never
    # blocks on the GPU, and a caller that looks at it immediately gets the same values as before.
    @classmethod
    def deferred(cls, resolver, image_size, mode="xyxy", on_host=True):
        """`on_host`: the resolved tensors live in host memory (then `.to("cpu")` has nothing to do and does not
        resolve); False for device-resident results whose per-frame detection count is still on the device."""
        self = cls.__new__(cls)
        self._bbox = None
        self._fields = {}
        self._pending = resolver
        self._pending_on_host = bool(on_host)
        self.size = image_size
        self.mode = mode
        return self

    def _materialize(self):
        if self._pending is not None:
            resolver, self._pending = self._pending, None
            bbox, fields = resolver()
            self._bbox = bbox
            for k, v in fields.items():
                self._fields.setdefault(k, v)

    @property
    def is_pending(self):
        return self._pending is not None

    @property
    def bbox(self):
        self._materialize()
        return self._bbox

    @bbox.setter
    def bbox(self, value):
        self._materialize()
        self._bbox = value

    @property
    def extra_fields(self):
        self._materialize()
        return self._fields

    def __getstate__(self):      # pickled like the reference's BoxList: bbox / size / mode / extra_fields
        self._materialize()
        return {"bbox": self._bbox, "size": self.size, "mode": self.mode, "extra_fields": self._fields}

    def __reduce_ex__(self, protocol):
        # class reference = mega_core.structures.bounding_box.BoxList (see the module docstring), state = its attributes
        # (copyreg._reconstructor is what protocol < 2 pickles of plain objects use; unlike __newobj__ it may name a class
        # other than type(self))
        return copyreg._reconstructor, (_pickle_class(), object, None), self.__getstate__()

    def __setstate__(self, state):
        self._bbox = state["bbox"]
        self._fields = state.get("extra_fields", {})
        self._pending = None
        self.size = state["size"]
        self.mode = state["mode"]

    # ---- fields
    def add_field(self, field, field_data):
        self.extra_fields[field] = field_data

    def get_field(self, field):
        return self.extra_fields[field]

    def has_field(self, field):
        return field in self.extra_fields

    def fields(self):
        return list(self.extra_fields.keys())

    def _copy_extra_fields(self, other):
        for k, v in other.extra_fields.items():
            self.extra_fields[k] = v

    # ---- geometry
    def _split_into_xyxy(self):
        if self.mode == "xyxy":
            return self.bbox.split(1, dim=-1)
        xmin, ymin, w, h = self.bbox.split(1, dim=-1)
        return xmin, ymin, xmin + (w - 1).clamp(min=0), ymin + (h - 1).clamp(min=0)

    def convert(self, mode):
        if mode not in ("xyxy", "xywh"):
            raise ValueError("mode should be 'xyxy' or 'xywh'")
        if mode == self.mode:
            return self
        xmin, ymin, xmax, ymax = self._split_into_xyxy()
        if mode == "xyxy":
            out = BoxList(torch.cat((xmin, ymin, xmax, ymax), dim=-1), self.size, mode=mode)
        else:
            out = BoxList(torch.cat((xmin, ymin, xmax - xmin + 1, ymax - ymin + 1), dim=-1), self.size, mode=mode)
        out._copy_extra_fields(self)
        return out

    def resize(self, size, *args, **kwargs):
        rw, rh = (float(s) / float(o) for s, o in zip(size, self.size))
        xmin, ymin, xmax, ymax = self._split_into_xyxy()
        out = BoxList(torch.cat((xmin * rw, ymin * rh, xmax * rw, ymax * rh), dim=-1), size, mode="xyxy")
        for k, v in self.extra_fields.items():
            if not isinstance(v, torch.Tensor) and hasattr(v, "resize"):
                v = v.resize(size, *args, **kwargs)
            out.add_field(k, v)
        return out.convert(self.mode)

    def clip_to_image(self, remove_empty=True):
        w, h = self.size
        self.bbox[:, 0].clamp_(min=0, max=w - 1)
        self.bbox[:, 1].clamp_(min=0, max=h - 1)
        self.bbox[:, 2].clamp_(min=0, max=w - 1)
        self.bbox[:, 3].clamp_(min=0, max=h - 1)
        if remove_empty:
            b = self.bbox
            return self[(b[:, 3] > b[:, 1]) & (b[:, 2] > b[:, 0])]
        return self

    def area(self):
        b = self.bbox
        if self.mode == "xyxy":
            return (b[:, 2] - b[:, 0] + 1) * (b[:, 3] - b[:, 1] + 1)
        return b[:, 2] * b[:, 3]

    # ---- container protocol
    def to(self, device):
        if self._pending is not None and getattr(self, "_pending_on_host", True) and torch.device(device).type == "cpu":
            return self          # deferred results land in host memory: nothing to move, nothing to wait for
        out = BoxList(self.bbox.to(device), self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v.to(device) if hasattr(v, "to") else v)
        return out

    def __getitem__(self, item):
        out = BoxList(self.bbox[item], self.size, self.mode)
        for k, v in self.extra_fields.items():
            out.add_field(k, v[item])
        return out

    def __len__(self):
        return self.bbox.shape[0]

    def copy_with_fields(self, fields, skip_missing=False):
        out = BoxList(self.bbox, self.size, self.mode)
        if not isinstance(fields, (list, tuple)):
            fields = [fields]
        for f in fields:
            if self.has_field(f):
                out.add_field(f, self.get_field(f))
            elif not skip_missing:
                raise KeyError("Field '{}' not found in {}".format(f, self))
        return out

    def __repr__(self):
        return "BoxList(num_boxes={}, image_width={}, image_height={}, mode={})".format(
            len(self), self.size[0], self.size[1], self.mode)


def cat_boxlist(bboxes):
    """Concatenate BoxLists that share size, mode and fields."""
    assert isinstance(bboxes, (list, tuple)) and all(isinstance(b, BoxList) for b in bboxes)
    size, mode, fields = bboxes[0].size, bboxes[0].mode, set(bboxes[0].fields())
    assert all(b.size == size and b.mode == mode and set(b.fields()) == fields for b in bboxes)
    out = BoxList(torch.cat([b.bbox for b in bboxes], dim=0), size, mode)
    for f in fields:
        out.add_field(f, torch.cat([b.get_field(f) for b in bboxes], dim=0))
    return out


class ImageList(object):
    """A batch of images padded to one size, plus each image's own (h, w)."""

    def __init__(self, tensors, image_sizes):
        self.tensors = tensors
        self.image_sizes = image_sizes

    def to(self, *args, **kwargs):
        return ImageList(self.tensors.to(*args, **kwargs), self.image_sizes)


def to_image_list(tensors, size_divisible=0):
    if isinstance(tensors, torch.Tensor) and size_divisible > 0:
        tensors = [tensors]
    if isinstance(tensors, ImageList) or (hasattr(tensors, "tensors") and hasattr(tensors, "image_sizes")):
        # any ImageList look-alike passes through untouched - in particular the reference's own class
        # (mega_core/structures/image_list.py:7-27), which is what engine/inference.py:35-40 hands to the model
        return tensors
    if isinstance(tensors, torch.Tensor):
        if tensors.dim() == 3:
            tensors = tensors[None]
        assert tensors.dim() == 4
        return ImageList(tensors, [t.shape[-2:] for t in tensors])
    if isinstance(tensors, (tuple, list)):
        max_size = [max(s) for s in zip(*[img.shape for img in tensors])]
        if size_divisible > 0:
            max_size[1] = int(math.ceil(max_size[1] / size_divisible) * size_divisible)
            max_size[2] = int(math.ceil(max_size[2] / size_divisible) * size_divisible)
        batched = tensors[0].new_zeros((len(tensors),) + tuple(max_size))
        for img, pad in zip(tensors, batched):
            pad[: img.shape[0], : img.shape[1], : img.shape[2]].copy_(img)
        return ImageList(batched, [im.shape[-2:] for im in tensors])
    raise TypeError("Unsupported type for to_image_list: {}".format(type(tensors)))

"""Checkpoint importer: load a reference-format ``.pth`` into ``diffusionvid_b200.model.DiffusionDet``.

Mirrors the behaviour of the reference's loader for this model, without its dependencies:

  mega_core/utils/checkpoint.py:52-82,113-114     ``checkpoint["model"]`` is the state dict (optimizer/scheduler ignored)
  mega_core/utils/model_serialization.py:77-87    a ``module.`` prefix is stripped when EVERY key carries it (DDP saves)
  mega_core/utils/model_serialization.py:89-138   DiffusionDet -> DiffusionVID renames: ``head_series.k`` with
                                                  k >= #head_series of the model becomes ``head_series_cond.(k - #)``
                                                  (only as many as the model has conditional heads), and
                                                  ``head_series_local`` becomes ``head_series_cond``
  mega_core/utils/model_serialization.py:12-75    each model key takes the loaded key that is its LONGEST string suffix
                                                  (plain ``endswith``, not dot-aligned); unmatched model keys keep their
                                                  current value
  mega_core/utils/model_serialization.py:140-156  the merged dict is then loaded strictly

The module's parameter names are the reference's, so nothing here touches layouts: BN folding, NHWC repacking and the
fp16 kernel layouts happen in ``DiffusionDet._pack`` the first time the model runs after a load.
"""
import re
from collections import OrderedDict

import torch

_SERIES, _COND, _LOCAL = "head_series", "head_series_cond", "head_series_local"


def strip_prefix_if_present(state_dict, prefix="module."):
    """Drop ``prefix`` only when all keys have it (model_serialization.py:77-87)."""
    if not state_dict or not all(k.startswith(prefix) for k in state_dict):
        return state_dict
    return OrderedDict((k[len(prefix):], v) for k, v in state_dict.items())


def _count_modules(model_keys, kind):
    idx = [int(k.split(kind + ".")[1][0]) for k in model_keys if kind + "." in k]
    return max(idx) + 1 if idx else 0


def rename_diffusiondet_heads(model_keys, state_dict):
    """DiffusionDet checkpoints number all six heads ``head_series.0..5``; DiffusionVID keeps the first ``n_series``
    there and calls the following ``n_cond`` ones ``head_series_cond.*`` (model_serialization.py:89-138).  Heads beyond
    ``n_series + n_cond`` keep their name (and later match nothing)."""
    n_series = _count_modules(model_keys, _SERIES)
    n_cond = _count_modules(model_keys, _COND)
    moved = ["%s.%d" % (_SERIES, i) for i in range(n_series, n_series + n_cond)]
    out = OrderedDict()
    for key, value in state_dict.items():
        if any(m in key for m in moved):
            # the reference rewrites the FIRST run of digits in the key (one character) - the head index
            m = re.search(r"\d+", key)
            k = key[:m.start()] + str(int(m.group()) - n_series) + key[m.start() + 1:]
            out[k.replace(_SERIES, _COND)] = value
        elif _LOCAL in key:
            out[key.replace(_LOCAL, _COND)] = value
        else:
            out[key] = value
    return out


def match_keys(model_keys, loaded_keys):
    """model key -> loaded key that is its longest suffix, or None (model_serialization.py:12-50)."""
    loaded = set(loaded_keys)
    mapping = {}
    for key in model_keys:
        hit = None
        for i in range(len(key)):            # longest suffix first
            if key[i:] in loaded:
                hit = key[i:]
                break
        mapping[key] = hit
    return mapping


def adapt_state_dict(model, loaded_state_dict):
    """Returns (merged state dict ready for a strict load, list of model keys left untouched)."""
    model_sd = model.state_dict()
    loaded = strip_prefix_if_present(loaded_state_dict, "module.")
    loaded = rename_diffusiondet_heads(list(model_sd.keys()), loaded)
    mapping = match_keys(sorted(model_sd.keys()), loaded.keys())
    missing = []
    for key, src in mapping.items():
        if src is None:
            missing.append(key)
            continue
        v = loaded[src]
        model_sd[key] = v if isinstance(v, torch.Tensor) else torch.as_tensor(v)
    return model_sd, missing


def load_state_dict(model, loaded_state_dict):
    """``mega_core.utils.model_serialization.load_state_dict`` for this model.  Returns the untouched model keys."""
    merged, missing = adapt_state_dict(model, loaded_state_dict)
    model.load_state_dict(merged)
    left = [k for k in missing if k.startswith("head.")]
    if left:
        # the model's default initialisation is random (synthetic weights): a head that silently keeps it produces
        # garbage detections, so say so loudly (the reference only logs "keys are not updated")
        import warnings
        warnings.warn("diffusionvid_b200.checkpoint: %d head parameters were not found in the checkpoint and keep "
                      "their random initialisation, e.g. %s" % (len(left), ", ".join(left[:3])), RuntimeWarning)
    return missing


def load_checkpoint(model, f, map_location="cpu"):
    """``DetectronCheckpointer.load`` for inference: ``f`` is a path to a ``.pth`` or an already loaded object; the
    weights live under ``"model"`` (checkpoint.py:113-114) - a bare state dict is accepted too."""
    ckpt = torch.load(f, map_location=map_location, weights_only=False) if isinstance(f, (str, bytes)) else f
    sd = ckpt["model"] if isinstance(ckpt, dict) and "model" in ckpt and not isinstance(ckpt["model"], torch.Tensor) \
        else ckpt
    return load_state_dict(model, sd)

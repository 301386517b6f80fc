#!/usr/bin/env python
"""Generate tests/golden/*.pt by RUNNING THE REFERENCE'S OWN CODE on the CPU of this container.

The reference cannot be imported as a package (torch._six, apex, detectron2, timm, yacs, nvidia-smi at
import).  This script therefore loads exactly three reference source files by path, unmodified, under their own
module names

    /root/reference/mega_core/modeling/roi_heads/box_head/box_head.py   (DynamicHead, RCNNHead, RCNNHead_cond,
                                                                          DynamicConv, SinusoidalPositionEmbeddings)
    /root/reference/mega_core/modeling/roi_heads/box_head/loss.py       (box_cxcywh_to_xyxy / box_xyxy_to_cxcywh)
    /root/reference/mega_core/modeling/detector/diffusion_det.py        (DiffusionDet: _forward_test, model_predictions,
                                                                          inference, update_erase_memory)
    + mega_core/structures/{bounding_box,image_list,boxlist_ops}.py     (BoxList, clip_to_image, to_image_list)

and satisfies every other import with stand-ins defined HERE (none of them is oracle code):

    detectron2.modeling.poolers.ROIPooler  -> level assignment per SURVEY.md A2 around the REAL torchvision.ops.roi_align
    detectron2.layers.batched_nms          -> the REAL torchvision.ops.batched_nms on boxes.float()
    detectron2.structures.Boxes            -> holder of `.tensor`
    detectron2.modeling.build_backbone     -> a frozen torch module computing R-x + FPN from the synthetic weights with
                                              plain F.conv2d (detectron2's ResNet/FPN source is not available; the
                                              backbone stays "parity unpinned")
    mega_core.layers.fps                   -> numpy farthest-point sampling with the tie rule of csrc/cuda/fps.cu:25-142
                                              (the CUDA kernel cannot run here; FPS stays pinned only by restatement)
    timm / apex / fvcore / yacs / other mega_core modules -> inert placeholders (never executed on this path)

Two CPU-only patches are applied while the reference runs (both hard-code CUDA): `.to('cuda')` maps to
the CPU and torch.cuda.{Int,Float}Tensor map to their CPU types.  Randomness: torch.randn / torch.randn_like inside
diffusion_det.py are replaced by a scripted provider that hands out oracle.NoiseSource tensors in the reference's own
draw order (box_init per split, img, then per step and frame: randn_like -> eps rows, randn -> replenish rows), which
is the explicit-noise contract of SURVEY.md 8c(1).

Run:  python tests/golden/make_golden.py      (needs /root/reference; writes tests/golden/ref_*.pt, a few 100 KB)
The fixtures are committed; tests/test_golden_reference.py compares the oracle (fp32) with them.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F
import torchvision

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("DVID_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)

from diffusionvid_b200 import synth  # noqa: E402  (synthetic weights / clips: inputs only)
from oracle.model import NoiseSource  # noqa: E402  (seeded noise tensors: inputs only)


# ------------------------------------------------------------------------------------------------ placeholders
class _Inert:
    """Callable/class placeholder for names that are imported but never executed on the inference path."""

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
    """Any not-yet-registered module under these roots resolves to an inert placeholder package."""
    ROOTS = ("mega_core", "detectron2", "timm", "apex", "fvcore", "yacs")

    def find_spec(self, name, path=None, target=None):
        if name.split(".")[0] in self.ROOTS:
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


# ------------------------------------------------------------------------------------------------ third-party stand-ins
class Boxes:
    def __init__(self, tensor):
        self.tensor = tensor

    def __len__(self):
        return self.tensor.shape[0]


class ROIPooler(torch.nn.Module):
    """detectron2.modeling.poolers.ROIPooler as configured by DynamicHead._init_box_pooler (box_head.py:250-271):
    ROIAlignV2 = torchvision roi_align(aligned=True); canonical_box_size 224, canonical_level 4."""

    def __init__(self, output_size, scales, sampling_ratio, pooler_type, canonical_box_size=224, canonical_level=4):
        super().__init__()
        assert pooler_type == "ROIAlignV2"
        self.output_size, self.scales, self.sampling_ratio = output_size, scales, sampling_ratio
        self.min_level = int(round(-np.log2(scales[0])))
        self.max_level = int(round(-np.log2(scales[-1])))
        self.canonical_box_size, self.canonical_level = canonical_box_size, canonical_level

    def forward(self, x, box_lists):
        rois = torch.cat([torch.cat([torch.full((len(b), 1), i, dtype=b.tensor.dtype), b.tensor], dim=1)
                          for i, b in enumerate(box_lists)], dim=0)
        t = rois[:, 1:]
        sizes = torch.sqrt((t[:, 2] - t[:, 0]) * (t[:, 3] - t[:, 1]))
        lvl = torch.floor(self.canonical_level + torch.log2(sizes / self.canonical_box_size + 1e-8))
        lvl = torch.clamp(lvl, min=self.min_level, max=self.max_level).to(torch.int64) - self.min_level
        out = torch.zeros((rois.shape[0], x[0].shape[1], self.output_size, self.output_size), dtype=x[0].dtype)
        for l, (feat, scale) in enumerate(zip(x, self.scales)):
            inds = torch.nonzero(lvl == l).squeeze(1)
            if inds.numel():
                out[inds] = torchvision.ops.roi_align(feat, rois[inds].to(feat.dtype), self.output_size, scale,
                                                      self.sampling_ratio, True)
        return out


def batched_nms(boxes, scores, idxs, thr):
    return torchvision.ops.batched_nms(boxes.float(), scores, idxs, thr)


class _Shape:
    def __init__(self, channels, stride):
        self.channels, self.stride = channels, stride


class SyntheticBackbone(torch.nn.Module):
    """R-x + FPN forward from the synthetic state dict (backbone.* keys) with plain torch ops, fp32."""

    size_divisibility = 32

    def __init__(self, sd):
        super().__init__()
        self.sd = {k: v for k, v in sd.items() if k.startswith("backbone.")}

    def output_shape(self):
        return {"p3": _Shape(256, 8), "p4": _Shape(256, 16), "p5": _Shape(256, 32), "p6": _Shape(256, 64)}

    def _cbn(self, x, name, stride, pad):
        sd = self.sd
        y = F.conv2d(x, sd[name + ".weight"], None, stride=stride, padding=pad)
        scale = sd[name + ".norm.weight"] * (sd[name + ".norm.running_var"] + 1e-5).rsqrt()
        bias = sd[name + ".norm.bias"] - sd[name + ".norm.running_mean"] * scale
        return y * scale.view(1, -1, 1, 1) + bias.view(1, -1, 1, 1)

    def forward(self, x):
        p = "backbone.bottom_up."
        x = F.max_pool2d(F.relu(self._cbn(x, p + "stem.conv1", 2, 3)), 3, 2, 1)
        outs = {}
        for si in range(4):
            bi = 0
            while (p + "res%d.%d.conv1.weight" % (si + 2, bi)) in self.sd:
                b = p + "res%d.%d." % (si + 2, bi)
                st = 2 if (bi == 0 and si > 0) else 1
                sc = self._cbn(x, b + "shortcut", st, 0) if (b + "shortcut.weight") in self.sd else x
                y = F.relu(self._cbn(x, b + "conv1", 1, 0))
                y = F.relu(self._cbn(y, b + "conv2", st, 1))
                x = F.relu(self._cbn(y, b + "conv3", 1, 0) + sc)
                bi += 1
            outs[si + 2] = x
        sd = self.sd
        lat = lambda l, t: F.conv2d(t, sd["backbone.fpn_lateral%d.weight" % l], sd["backbone.fpn_lateral%d.bias" % l])
        out = lambda l, t: F.conv2d(t, sd["backbone.fpn_output%d.weight" % l], sd["backbone.fpn_output%d.bias" % l],
                                    padding=1)
        prev = lat(5, outs[5]); p5 = out(5, prev)
        prev = lat(4, outs[4]) + F.interpolate(prev, scale_factor=2.0, mode="nearest"); p4 = out(4, prev)
        prev = lat(3, outs[3]) + F.interpolate(prev, scale_factor=2.0, mode="nearest"); p3 = out(3, prev)
        return {"p3": p3, "p4": p4, "p5": p5, "p6": F.max_pool2d(p5, kernel_size=1, stride=2)}


def fps_numpy(b, n, m, dist, temp, idx):
    """mega_core._C.furthest_point_sampling semantics (csrc/fps.h:15-36, csrc/cuda/fps.cu:25-142) on CPU tensors:
    start at index 0, running min of the picked rows, arg-max with ties resolved to the lowest index."""
    d = dist.reshape(b, n, n).numpy()
    for bi in range(b):
        t = np.full(n, 1e10, dtype=np.float32)
        old = 0
        idx[bi, 0] = 0
        for j in range(1, m):
            t = np.minimum(t, d[bi, old])
            old = int(np.argmax(t))          # first maximum = lowest index
            idx[bi, j] = old
    return 1


class CN(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError:
            raise AttributeError(k)

    def __setattr__(self, k, v):
        self[k] = v


def make_cfg(N, T, mem_size, num_heads=3, num_heads_local=1, global_enable=True):
    c = CN()
    c.MODEL = CN(DEVICE="cpu", PIXEL_MEAN=[123.675, 116.280, 103.530], PIXEL_STD=[58.395, 57.120, 57.375])
    c.MODEL.DiffusionDet = CN(NUM_CLASSES=30, NUM_PROPOSALS=N, NHEADS=8, DROPOUT=0.0, DIM_FEEDFORWARD=2048,
                              ACTIVATION="relu", HIDDEN_DIM=256, NUM_CLS=1, NUM_REG=3, NUM_HEADS=num_heads,
                              NUM_HEADS_LOCAL=num_heads_local, NUM_DYNAMIC=2, DIM_DYNAMIC=64, CLASS_WEIGHT=2.0,
                              GIOU_WEIGHT=2.0, L1_WEIGHT=5.0, DEEP_SUPERVISION=True, NO_OBJECT_WEIGHT=0.1,
                              USE_FOCAL=True, USE_FED_LOSS=False, ALPHA=0.25, GAMMA=2.0, PRIOR_PROB=0.01, OTA_K=5,
                              SNR_SCALE=2.0, SAMPLE_STEP=T, USE_NMS=True)
    c.MODEL.ROI_HEADS = CN(IN_FEATURES=["p3", "p4", "p5"])
    c.MODEL.ROI_BOX_HEAD = CN(POOLER_RESOLUTION=7, POOLER_SAMPLING_RATIO=2, POOLER_TYPE="ROIAlignV2")
    c.MODEL.VID = CN(RPN=CN(REF_POST_NMS_TOP_N=75), ROI_BOX_HEAD=CN(ATTENTION=CN(ENABLE=False, STAGE=1)))
    c.MODEL.VID.MEGA = CN(ALL_FRAME_INTERVAL=8, KEY_FRAME_LOCATION=0, MEMORY_MANAGEMENT_METRIC="distance",
                          MEMORY_MANAGEMENT_SIZE_TEST=mem_size, RATIO=0.2,
                          GLOBAL=CN(ENABLE=global_enable, RES_STAGE=1, SIZE=4))
    c.INPUT = CN(INFER_BATCH=8)
    return c


class ScriptedNoise:
    """Feeds NoiseSource tensors to the reference's torch.randn / torch.randn_like calls in its own draw order."""

    def __init__(self, noise, N, T):
        self.noise, self.N, self.T = noise, N, T
        self.video = 0
        self.key = None

    def begin_call(self, fid, n_init):
        self.fid = fid
        self.n_init = n_init        # extraction splits of this call: one box_init draw each (diffusion_det.py:449)
        self.split = 0
        self.stage = "init"
        self.step = 0
        self.frame = 0

    def randn(self, *shape, **kw):
        shape = tuple(shape[0]) if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)) else shape
        if len(shape) == 3 and self.split < self.n_init:
            out = self.noise.get("init", self.video, self.fid, self.split, shape[0])
            self.split += 1
            return out
        if len(shape) == 3:                      # img = torch.randn(shape)
            self.stage = "loop"
            self.batch = shape[0]
            return self.noise.get("img", self.video, self.fid, 0, shape[0])
        # replenish: torch.randn(N - num_remain[i], 4)
        rows = int(shape[0])
        out = self.noise.get("fill", self.video, self.fid, self.step, self.batch)[self.frame, :rows]
        self.frame += 1
        if self.frame == self.batch:
            self.frame = 0
            self.step += 1
        return out

    def randn_like(self, t):
        return self.noise.get("eps", self.video, self.fid, self.step, self.batch)[self.frame, :t.shape[0]]


def build_reference(sd, cfg):
    sys.meta_path.insert(0, _Finder())
    d2p = _InertModule("detectron2.modeling.poolers"); d2p.ROIPooler = ROIPooler
    d2s = _InertModule("detectron2.structures"); d2s.Boxes = Boxes
    d2l = _InertModule("detectron2.layers"); d2l.batched_nms = batched_nms
    d2m = _InertModule("detectron2.modeling"); d2m.__path__ = []
    d2m.build_backbone = lambda cfg_: SyntheticBackbone(sd)
    for m in (d2p, d2s, d2l, d2m):
        sys.modules[m.__name__] = m
    layers = _InertModule("mega_core.layers"); layers.__path__ = []
    layers.fps = fps_numpy
    sys.modules["mega_core.layers"] = layers
    yc = _InertModule("yacs.config"); yc.CfgNode = CN
    sys.modules["yacs.config"] = yc
    # real reference sources, unmodified
    _load("mega_core.structures.bounding_box", "mega_core/structures/bounding_box.py")
    _load("mega_core.structures.image_list", "mega_core/structures/image_list.py")
    _load("mega_core.structures.boxlist_ops", "mega_core/structures/boxlist_ops.py")
    loss = _load("mega_core.modeling.roi_heads.box_head.loss", "mega_core/modeling/roi_heads/box_head/loss.py")
    loss.SetCriterionDynamicK = _Inert       # training-only classes (their constructors read training cfg keys)
    loss.HungarianMatcherDynamicK = _Inert
    bh = _load("mega_core.modeling.roi_heads.box_head.box_head", "mega_core/modeling/roi_heads/box_head/box_head.py")
    dd = _load("mega_core.modeling.detector.diffusion_det", "mega_core/modeling/detector/diffusion_det.py")
    model = dd.DiffusionDet(cfg)
    head_sd = {k[len("head."):]: v for k, v in sd.items() if k.startswith("head.")}
    missing, unexpected = model.head.load_state_dict(head_sd, strict=False)
    assert not missing and not unexpected, (missing, unexpected)     # pins the state-dict key names
    model.eval()
    return model, dd, bh


def run():
    torch.manual_seed(0)
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    N, mem = 50, 60
    h, w, L = 96, 160, 11
    blocks = (1, 1, 1, 1)
    sd = synth.make_state_dict(seed=77, blocks=blocks)
    frames = synth.make_clip(L, h, w, seed=78)
    gidx = [9, 2, 5]

    # CPU-only patches for the two hard-coded CUDA spots (diffusion_det.py:590-591, :892-893)
    orig_to = torch.Tensor.to

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **k)

    torch.Tensor.to = to_cpu
    torch.cuda.IntTensor = torch.IntTensor
    torch.cuda.FloatTensor = torch.FloatTensor

    out = {"meta": dict(N=N, mem_size=mem, h=h, w=w, L=L, blocks=blocks, weight_seed=77, clip_seed=78, global_idx=gidx,
                        noise_seed=79, reference="sdroh1027/DiffusionVID@8375542", torch=torch.__version__,
                        torchvision=torchvision.__version__)}
    model = dd = None
    for T in (1, 4):
        cfg = make_cfg(N, T, mem)
        if model is None:
            model, dd, bh = build_reference(sd, cfg)
        else:
            model = dd.DiffusionDet(cfg)
            model.head.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("head.")})
            model.eval()
        noise = NoiseSource(79, N)
        sn = ScriptedNoise(noise, N, T)
        real_randn, real_randn_like = torch.randn, torch.randn_like

        class _TorchProxy:
            """`torch` as seen from diffusion_det.py: randn / randn_like scripted, everything else real."""

            def __getattr__(self, name):
                if name == "randn":
                    return sn.randn
                if name == "randn_like":
                    return sn.randn_like
                return getattr(torch, name)

        dd.torch = _TorchProxy()
        il = sys.modules["mega_core.structures.image_list"]
        samples = synth.clip_samples(frames, gidx, h, w)
        per_frame = []
        with torch.no_grad():
            for s in samples:
                queued = len(model.local_img_queue) if s["frame_category"] != 0 else 0
                is_key = s["frame_id"] % 8 == 0
                n_new = queued + len(s["ref_l"]) + len(s["ref_g"])
                sn.begin_call(s["frame_id"], (n_new + 7) // 8 if is_key else 0)
                mk = lambda t: il.ImageList(t, [(h, w)])
                infos = dict(cur=mk(s["cur"]), ref_l=[mk(t) for t in s["ref_l"]], ref_g=[mk(t) for t in s["ref_g"]],
                             frame_id=s["frame_id"], start_id=0, end_id=L - 1, seg_len=L,
                             last_queue_id=s["last_queue_id"], frame_category=s["frame_category"])
                res = model._forward_test(infos["cur"], infos)
                for bl in res:
                    sc = bl.get_field("scores"); lb = bl.get_field("labels")
                    order = torch.sort(sc, descending=True, stable=True)[1]
                    per_frame.append(dict(boxes=bl.bbox[order].clone(), scores=sc[order].clone(),
                                          labels=lb[order].clone(), size=bl.size))
        assert len(per_frame) == L
        out["clip_T%d" % T] = per_frame
        out["mem_T%d" % T] = [m_.clone() for m_ in model.head.proposal_feats_global]
        dd.torch = torch

    # head-level vectors: DynamicHead.forward in extraction mode (3 base heads + top-k) on backbone features
    cfg = make_cfg(N, 4, mem)
    model = dd.DiffusionDet(cfg)
    model.head.load_state_dict({k[5:]: v for k, v in sd.items() if k.startswith("head.")})
    model.eval()
    noise = NoiseSource(79, N)
    with torch.no_grad():
        imgs = model.normalizer(frames[:2])
        feats = model.backbone(imgs)
        f = [feats[p] for p in ("p3", "p4", "p5")]
        whwh = torch.tensor([w, h, w, h], dtype=torch.float32)[None].expand(2, -1)
        x = noise.get("init", 0, 0, 0, 2)
        t = torch.full((2,), 999, dtype=torch.long)
        (lg, bx, obj), k1, k2 = model.model_predictions(f, whwh, x, t, None, clip_x_start=True, box_extract=1)
        temb = model.head.time_mlp(torch.tensor([999, 749, 0]))
    out["head"] = dict(p3=f[0].clone(), p4=f[1].clone(), p5=f[2].clone(), logits=lg.clone(), boxes=bx.clone(),
                       obj=obj.clone(), k1=k1.clone(), k2=k2.clone(), time_emb=temb.clone())
    out["schedule"] = dict(alphas_cumprod=model.alphas_cumprod.clone(),
                           sqrt_recip=model.sqrt_recip_alphas_cumprod.clone(),
                           sqrt_recipm1=model.sqrt_recipm1_alphas_cumprod.clone())
    torch.Tensor.to = orig_to
    path = os.path.join(HERE, "ref_diffusionvid_small.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    run()

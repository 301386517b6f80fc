"""CPU restatement of the DiffusionVID inference path (DiffusionDet + DynamicHead + R-101-FPN), functional over a
reference-keyed state dict.  TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Parity pin: the reference has no tests or golden vectors for this code, so the oracle is pinned against OUTPUTS OF THE
REFERENCE ITSELF run in this container: tests/golden/make_golden.py executes the unmodified reference sources
diffusion_det.py / box_head.py / loss.py / structures/* on the CPU (third-party imports satisfied by torchvision's
real roi_align / batched_nms and inert placeholders) and tests/test_golden_reference.py checks this file against the
committed results (DynamicHead chain to 1e-4, whole T=1 / T=4 clips to 1e-4 px / 1e-5 score).  Two pieces remain
restatement-only ("parity unpinned"): the detectron2 R-101+FPN backbone (source absent) and the tie rule
of the CUDA farthest-point-sampling kernel (csrc/cuda/fps.cu cannot run without a GPU build of the old extension).
Sub-operators are additionally pinned against the installed torchvision / torch in tests/test_oracle_ops.py.

Reference files restated (sdroh1027/DiffusionVID @ 8375542):
  mega_core/modeling/detector/diffusion_det.py:43-61,222-267 (schedule), :377-646 (_forward_test), :649-677
  (model_predictions), :754-817 (inference), :841-896 (global memory);
  mega_core/modeling/roi_heads/box_head/box_head.py:273-435 (DynamicHead.forward), :495-590 (RCNNHead),
  :605-664 (RCNNHead_cond), :687-711 (DynamicConv), :729-741 (time embedding);
  mega_core/modeling/roi_heads/box_head/loss.py:201-212 (box format helpers).
detectron2's ResNet/FPN/ROIPooler/batched_nms are restated from SURVEY.md Appendix A (source not in the tree).

Two arithmetic modes:
  Quant(False): plain fp32 everywhere (the "fp32 oracle").
  Quant(True):  fp32 arithmetic with values rounded to fp16 at exactly the points where the sm_100a kernels store
                fp16 (GEMM operands, feature maps, ROI features, generated dynamic weights); accumulation stays fp32.
                This is the checker for the fp16 product path ("fp16 oracle").
Randomness is an explicit input (NoiseSource) - SURVEY.md 8c contract (1).
"""
import math
import zlib
from collections import deque

import torch
import torch.nn.functional as F

from . import ops


class Quant:
    def __init__(self, enabled):
        self.enabled = enabled
        self._wcache = {}

    def a(self, x):
        return x.half().float() if self.enabled else x

    def w(self, t):
        if not self.enabled:
            return t
        key = id(t)
        if key not in self._wcache:
            self._wcache[key] = (t, t.half().float())
        return self._wcache[key][1]


class NoiseSource:
    """Seeded CPU Gaussian noise with a fixed draw order; the same class feeds the oracle and the product."""

    def __init__(self, seed, num_proposals):
        self.seed = seed
        self.n = num_proposals
        self._cache = {}

    def _draw(self, key, frames):
        if key not in self._cache:
            g = torch.Generator(device="cpu").manual_seed(zlib.crc32(repr((self.seed,) + key).encode()) % (2 ** 31))
            self._cache[key] = torch.randn(8 * 64, self.n, 4, generator=g)
        return self._cache[key][:frames].clone()

    # kind: "init" (extraction box_init, diffusion_det.py:449), "img" (:542), "eps" (:587), "fill" (:595)
    def get(self, kind, video, key_frame, index, frames):
        return self._draw((kind, int(video), int(key_frame), int(index)), frames)


# ------------------------------------------------------------------------------------------------ helpers
def box_cxcywh_to_xyxy(x):
    """box_head/loss.py:201-205."""
    cx, cy, w, h = x.unbind(-1)
    return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], dim=-1)


def box_xyxy_to_cxcywh(x):
    """box_head/loss.py:208-212."""
    x0, y0, x1, y1 = x.unbind(-1)
    return torch.stack([(x0 + x1) / 2, (y0 + y1) / 2, x1 - x0, y1 - y0], dim=-1)


def cosine_alphas_cumprod(timesteps=1000, s=0.008):
    """diffusion_det.py:50-61,226-228: float64 schedule, alphas_cumprod cast to fp32."""
    steps = timesteps + 1
    x = torch.linspace(0, timesteps, steps, dtype=torch.float64)
    ac = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    ac = ac / ac[0]
    betas = torch.clip(1 - (ac[1:] / ac[:-1]), 0, 0.999)
    return torch.cumprod(1. - betas, dim=0).to(torch.float32)


def ln(x, sd, name, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


class Ctx:
    """state dict + arithmetic policy."""

    def __init__(self, sd, quant):
        self.sd = sd
        self.q = quant
        self._folded = {}

    def folded(self, name):
        """FrozenBatchNorm2d(eps=1e-5) folded into the conv: w' = w*scale, b' = beta - mean*scale (the state dict is
        constant for the life of the oracle, so the folded pair is computed once per conv)."""
        if name not in self._folded:
            sd = self.sd
            scale = sd[name + ".norm.weight"] * torch.rsqrt(sd[name + ".norm.running_var"] + 1e-5)
            shift = sd[name + ".norm.bias"] - sd[name + ".norm.running_mean"] * scale
            wf = sd[name + ".weight"] * scale[:, None, None, None]
            if self.q.enabled:
                wf = wf.half().float()
            self._folded[name] = (wf, shift)
        return self._folded[name]

    def lin(self, x, name, bias=True, out16=False, round_in=True):
        w = self.q.w(self.sd[name + ".weight"])
        b = self.sd.get(name + ".bias") if bias else None
        y = F.linear(self.q.a(x) if round_in else x, w, b)
        return self.q.a(y) if out16 else y

    def lin_rows(self, x, name, r0, r1, out16=False):
        """slice of a packed in_proj (rows r0:r1 of in_proj_weight / in_proj_bias)."""
        w = self.q.w(self.sd[name + "_weight"])[r0:r1]
        b = self.sd[name + "_bias"][r0:r1]
        y = F.linear(self.q.a(x), w, b)
        return self.q.a(y) if out16 else y


# ------------------------------------------------------------------------------------------------ backbone
def _conv_bn(c, x, name, stride, pad, relu, resid=None):
    """conv (no bias) + FrozenBatchNorm2d(eps=1e-5) folded: w' = w*scale, b' = beta - mean*scale; optional residual;
    ReLU; output rounded to fp16 in emulation mode (the kernels store NHWC fp16)."""
    wf, shift = c.folded(name)
    y = F.conv2d(x, wf, shift, stride=stride, padding=pad)
    if resid is not None:
        y = y + resid
    if relu:
        y = F.relu(y)
    return c.q.a(y)


def _conv_bias(c, x, name, pad, resid=None):
    y = F.conv2d(x, c.q.w(c.sd[name + ".weight"]), c.sd[name + ".bias"], padding=pad)
    if resid is not None:
        y = y + resid
    return c.q.a(y)


R101_BLOCKS = (3, 4, 23, 3)


def _count_blocks(sd):
    out = []
    for si in range(4):
        n = 0
        while ("backbone.bottom_up.res%d.%d.conv1.weight" % (si + 2, n)) in sd:
            n += 1
        out.append(n)
    return tuple(out)


def resnet_fpn(c, x, blocks=None):
    """detectron2 build_resnet_fpn_backbone (R-101, STRIDE_IN_1X1 False, FPN on res3..res5) - SURVEY.md A1.
    x: normalised (B,3,H,W) -> [p3,p4,p5]."""
    p = "backbone.bottom_up."
    if blocks is None:
        blocks = _count_blocks(c.sd)
    x = c.q.a(x)
    x = _conv_bn(c, x, p + "stem.conv1", 2, 3, True)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    outs = {}
    for si, nb in enumerate(blocks):
        stage = "res%d" % (si + 2)
        for bi in range(nb):
            b = "%s%s.%d." % (p, stage, bi)
            stride = 2 if (bi == 0 and si > 0) else 1
            if (b + "shortcut.weight") in c.sd:
                sc = _conv_bn(c, x, b + "shortcut", stride, 0, False)
            else:
                sc = x
            y = _conv_bn(c, x, b + "conv1", 1, 0, True)
            y = _conv_bn(c, y, b + "conv2", stride, 1, True)
            x = _conv_bn(c, y, b + "conv3", 1, 0, True, resid=sc)
        outs[stage] = x
    prev = _conv_bias(c, outs["res5"], "backbone.fpn_lateral5", 0)
    p5 = _conv_bias(c, prev, "backbone.fpn_output5", 1)
    prev = _conv_bias(c, outs["res4"], "backbone.fpn_lateral4", 0,
                      resid=F.interpolate(prev, scale_factor=2.0, mode="nearest"))
    p4 = _conv_bias(c, prev, "backbone.fpn_output4", 1)
    prev = _conv_bias(c, outs["res3"], "backbone.fpn_lateral3", 0,
                      resid=F.interpolate(prev, scale_factor=2.0, mode="nearest"))
    p3 = _conv_bias(c, prev, "backbone.fpn_output3", 1)
    return [p3, p4, p5]


# ------------------------------------------------------------------------------------------------ head
SCALE_CLAMP = math.log(100000.0 / 16)


def apply_deltas(deltas, boxes):
    """box_head.py:550-590, weights (2,2,1,1), k = 1."""
    w = boxes[:, 2] - boxes[:, 0]
    h = boxes[:, 3] - boxes[:, 1]
    cx = boxes[:, 0] + 0.5 * w
    cy = boxes[:, 1] + 0.5 * h
    dx = deltas[:, 0] / 2.0
    dy = deltas[:, 1] / 2.0
    dw = torch.clamp(deltas[:, 2] / 1.0, max=SCALE_CLAMP)
    dh = torch.clamp(deltas[:, 3] / 1.0, max=SCALE_CLAMP)
    pcx = dx * w + cx
    pcy = dy * h + cy
    pw = torch.exp(dw) * w
    ph = torch.exp(dh) * h
    return torch.stack([pcx - 0.5 * pw, pcy - 0.5 * ph, pcx + 0.5 * pw, pcy + 0.5 * ph], dim=1)


def attention(c, xq, xkv, name, nheads):
    """nn.MultiheadAttention forward with fp16 storage of q/k/v/ctx in emulation mode.
    xq (L,Bt,E), xkv (S,Bt,E) -> (L,Bt,E)."""
    L, Bt, E = xq.shape
    S = xkv.shape[0]
    hd = E // nheads
    q = c.lin_rows(xq.reshape(L * Bt, E), name + ".in_proj", 0, E, out16=True).view(L, Bt, nheads, hd)
    k = c.lin_rows(xkv.reshape(S * Bt, E), name + ".in_proj", E, 2 * E, out16=True).view(S, Bt, nheads, hd)
    v = c.lin_rows(xkv.reshape(S * Bt, E), name + ".in_proj", 2 * E, 3 * E, out16=True).view(S, Bt, nheads, hd)
    att = torch.einsum("lbhd,sbhd->bhls", q, k) * (1.0 / math.sqrt(hd))
    att = torch.softmax(att, dim=-1)
    ctx = c.q.a(torch.einsum("bhls,sbhd->lbhd", att, v).reshape(L * Bt, E))
    return c.lin(ctx, name + ".out_proj").view(L, Bt, E)


def time_embedding(c, t, dim=256):
    """box_head.py:218-223,729-741: sinusoid -> Linear -> GELU -> Linear.  t (B,) -> (B, 4*dim)."""
    half = dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32, device=t.device) * -(math.log(10000) / (half - 1)))
    e = t.float()[:, None] * freq[None, :]
    e = torch.cat((e.sin(), e.cos()), dim=-1)
    e = c.lin(e, "head.time_mlp.1", round_in=False)
    e = F.gelu(e)
    return c.lin(e, "head.time_mlp.3", round_in=False)


def dynamic_conv(c, pre, pro, roi):
    """DynamicConv.forward, box_head.py:687-711.  pro (M,256), roi (M,49,256) -> (M,256) (before the residual)."""
    sd = c.sd
    M = pro.shape[0]
    d = pro.shape[1]
    params = c.lin(pro, pre + "dynamic_layer", out16=True)
    dd = params.shape[1] // (2 * d)
    p1 = params[:, :d * dd].view(M, d, dd)
    p2 = params[:, d * dd:].view(M, dd, d)
    f = torch.bmm(roi, p1)
    f = c.q.a(F.relu(ln(f, sd, pre + "norm1")))
    f = torch.bmm(f, p2)
    f = c.q.a(F.relu(ln(f, sd, pre + "norm2")))
    f = c.lin(f.flatten(1), pre + "out_layer")
    return F.relu(ln(f, sd, pre + "norm3"))


def rcnn_head(c, pre, feats, boxes, pro, time_emb, cfg, cond=None):
    """RCNNHead.forward (box_head.py:495-548) / RCNNHead_cond.forward (:605-664).
    feats [p3,p4,p5] (B,C,H,W); boxes (B,N,4); pro (M,256) or None; time_emb (B,1024); cond (M,256) or None.
    Returns logits (B,N,C), boxes (B,N,4), obj (M,256)."""
    sd = c.sd
    B, N = boxes.shape[:2]
    d = cfg["hidden"]
    roi = c.q.a(ops.roi_pooler(feats, boxes))                       # (M,d,7,7)
    roi = roi.view(B * N, d, -1).permute(0, 2, 1).contiguous()      # (M,49,d): position-major, box_head.py:512
    if pro is None:
        pro = roi.mean(1)
    x = pro.view(B, N, d).permute(1, 0, 2)                          # (N,B,d): self-attention is per frame
    x2 = attention(c, x, x, pre + "self_attn", cfg["nheads"])
    pro = ln((x + x2).permute(1, 0, 2).reshape(B * N, d), sd, pre + "norm1")
    pro2 = dynamic_conv(c, pre + "inst_interact.", pro, roi)
    obj = ln(pro + pro2, sd, pre + "norm2")
    h = c.q.a(F.relu(c.lin(obj, pre + "linear1")))
    obj = ln(obj + c.lin(h, pre + "linear2"), sd, pre + "norm3")
    temb = F.silu(time_emb)
    if cond is None:
        ss = c.lin(temb, pre + "block_time_mlp.1", round_in=False).repeat_interleave(N, dim=0)
        scale, shift = ss.chunk(2, dim=1)
    else:
        scale = c.lin(temb, pre + "block_time_mlp.1", round_in=False).repeat_interleave(N, dim=0)
        shift = c.lin(F.silu(cond), pre + "c_mlp.1")
    fc = obj * (scale + 1) + shift
    cls = fc
    for i in range(cfg["num_cls"]):
        cls = F.relu(ln(c.lin(cls, pre + "cls_module.%d" % (3 * i), bias=False), sd, pre + "cls_module.%d" % (3 * i + 1)))
    reg = fc
    for i in range(cfg["num_reg"]):
        reg = F.relu(ln(c.lin(reg, pre + "reg_module.%d" % (3 * i), bias=False), sd, pre + "reg_module.%d" % (3 * i + 1)))
    logits = c.lin(cls, pre + "class_logits")
    deltas = c.lin(reg, pre + "bboxes_delta")
    pred = apply_deltas(deltas, boxes.reshape(-1, 4))
    return logits.view(B, N, -1), pred.view(B, N, 4), obj


def global_attention(c, obj, mem, cfg):
    """box_head.py:349,366-371,382-394 (adaptive_norm=True): attn_ = MHA(query=obj, key=value=mem); no residual."""
    out = attention(c, obj[:, None, :], mem[:, None, :], "head.global_attention.0.0", cfg["nheads"])
    return out[:, 0, :]


def head_base_stages(c, feats, boxes, time_emb, cfg):
    """head_series[0..num_heads) on the same boxes chain (box_head.py:294-299)."""
    pro = None
    logits = None
    for i in range(cfg["num_heads"]):
        logits, boxes, pro = rcnn_head(c, "head.head_series.%d." % i, feats, boxes, pro, time_emb, cfg)
    return logits, boxes, pro


def select_topk_feats(logits, obj, B, N, ks):
    """box_head.py:304-317: per-frame top-k by max logit -> bool mask -> rows in box-index order."""
    mx = logits.max(dim=-1)[0]                                       # (B,N)
    order = torch.sort(mx, dim=-1, descending=True, stable=True)[1]  # ties: lower box index first (canonical)
    outs = []
    objf = obj.view(B, N, -1)
    for k in ks:
        mask = torch.zeros_like(mx, dtype=torch.bool)
        mask.scatter_(1, order[:, :k], True)
        outs.append(objf[mask])
    return outs


def update_erase_memory(feats_new, feats_mem, target):
    """diffusion_det.py:841-896 with mem_management_type='greedy'."""
    merged = feats_new if feats_mem is None else torch.cat([feats_mem, feats_new], dim=0)
    if merged.shape[0] <= target:
        return merged
    dist = ops.cdist_l2(merged.contiguous())
    idx = ops.fps(dist.cpu().numpy(), target)
    return merged[torch.from_numpy(idx).long().to(merged.device)]


def topk_scores(logits_frame, boxes_frame, n):
    """diffusion_det.py:772-784 for one frame: sigmoid, flatten (box-major), top-n.  Canonical order: score desc then
    flat index asc (reference uses sorted=False, whose order is unspecified; SURVEY.md 8c contract 3)."""
    C = logits_frame.shape[-1]
    scores = torch.sigmoid(logits_frame).flatten()
    order = torch.sort(scores, descending=True, stable=True)[1][:n]
    return boxes_frame[order // C], scores[order], (order % C) + 1


def finalize_frame(boxes, scores, labels, size_wh, use_nms=True):
    """diffusion_det.py:616-627 / :792-812: batched_nms(0.5) then BoxList.clip_to_image (legacy TO_REMOVE=1)."""
    if use_nms:
        keep = ops.batched_nms(boxes, scores, labels, 0.5)
        boxes, scores, labels = boxes[keep], scores[keep], labels[keep]
    W, H = size_wh
    boxes = boxes.clone()
    boxes[:, 0].clamp_(min=0, max=W - 1)
    boxes[:, 1].clamp_(min=0, max=H - 1)
    boxes[:, 2].clamp_(min=0, max=W - 1)
    boxes[:, 3].clamp_(min=0, max=H - 1)
    return {"boxes": boxes, "scores": scores, "labels": labels}


DEFAULT_CFG = dict(num_proposals=300, num_classes=30, hidden=256, nheads=8, num_heads=3, num_heads_local=1,
                   sample_step=4, num_cls=1, num_reg=3, snr_scale=2.0, infer_batch=8, all_frame_interval=8,
                   key_frame_location=0, global_enable=True, mem_size=900, mem_size2=150, topk=(75, 25),
                   pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375), use_nms=True)


class OracleDiffusionVID:
    """The clip state machine of DiffusionDet._forward_test (diffusion_det.py:377-646) over plain tensors.

    forward(sample) with sample = dict(cur=(1,3,Hp,Wp) padded image in [0,1], image_size=(h,w) unpadded,
    ref_l=[(1,3,Hp,Wp)...], ref_g=[...], frame_id, start_id, end_id, seg_len, frame_category, video_id)
    returns [] on non-key frames, else one dict(boxes,scores,labels) per frame of the key batch.
    """

    def __init__(self, sd, cfg=None, fp16=False, noise=None):
        self.cfg = dict(DEFAULT_CFG)
        if cfg:
            self.cfg.update(cfg)
        self.c = Ctx(sd, Quant(fp16))
        self.noise = noise
        self.alphas_cumprod = cosine_alphas_cumprod()
        self.trace = {}

    # -- diffusion_det.py:655-677
    def _x_to_boxes(self, x, whwh):
        s = self.cfg["snr_scale"]
        xb = torch.clamp(x, min=-s, max=s)
        xb = ((xb / s) + 1) / 2
        return box_cxcywh_to_xyxy(xb) * whwh[:, None, :]

    def _x_start(self, coord, whwh):
        s = self.cfg["snr_scale"]
        xs = coord / whwh[0][None, None, :]
        xs = box_xyxy_to_cxcywh(xs)
        xs = (xs * 2 - 1.) * s
        return torch.clamp(xs, min=-s, max=s)

    def backbone(self, imgs):
        mean = (torch.tensor(self.cfg["pixel_mean"]).view(1, 3, 1, 1) / 255.).to(imgs.device)
        std = (torch.tensor(self.cfg["pixel_std"]).view(1, 3, 1, 1) / 255.).to(imgs.device)
        if "backbone.bottom_up.patch_embed.proj.weight" in self.c.sd:      # vid_Swin_B_DiffusionVID.yaml
            from . import swin
            return swin.swin_fpn(self.c, (imgs - mean) / std)
        return resnet_fpn(self.c, (imgs - mean) / std)

    def forward(self, s):
        cfg = self.cfg
        c = self.c
        N = cfg["num_proposals"]
        ib = cfg["infer_batch"]
        if s["frame_category"] == 0:
            self.local_img_queue = []
            self.mem = [None, None]
            self.feats = deque(maxlen=cfg["all_frame_interval"])
            self.cache = deque(maxlen=cfg["all_frame_interval"])
            self.video = s.get("video_id", 0)
        fid = s["frame_id"]
        if fid % ib != 0:
            self.local_img_queue += list(s["ref_l"])
            return []
        ref_l = self.local_img_queue + list(s["ref_l"])
        self.local_img_queue = []
        h, w = s["image_size"]
        dev = s["cur"].device        # CPU everywhere except the library-kernel timing bar (tests/test_gpu_library_bar.py)
        whwh1 = torch.tensor([w, h, w, h], dtype=torch.float32, device=dev)

        if ref_l or s["ref_g"]:
            imgs = torch.cat(ref_l + list(s["ref_g"]))
            len_l = len(ref_l)
            logits_all, boxes_all, obj_all, k1_all, k2_all, feats_all = [], [], [], [], [], []
            for bi, split in enumerate(imgs.split(ib)):
                f = self.backbone(split)
                B = split.shape[0]
                whwh = whwh1[None].expand(B, -1)
                box_init = self.noise.get("init", self.video, fid, bi, B).to(dev)
                t = torch.full((B,), 999, dtype=torch.long, device=dev)
                temb = time_embedding(c, t)
                lg, bx, obj = head_base_stages(c, f, self._x_to_boxes(box_init, whwh), temb, cfg)
                k1, k2 = select_topk_feats(lg, obj, B, N, [min(k, N) for k in cfg["topk"]])
                logits_all.append(lg); boxes_all.append(bx); obj_all.append(obj.view(B, N, -1))
                k1_all.append(k1); k2_all.append(k2); feats_all.append(f)
            logits_t = torch.cat(logits_all); boxes_t = torch.cat(boxes_all); obj_t = torch.cat(obj_all)
            feats_t = [torch.cat([f[l] for f in feats_all]) for l in range(3)]
            k1_t = torch.cat(k1_all).view(-1, min(cfg["topk"][0], N), cfg["hidden"])
            k2_t = torch.cat(k2_all).view(-1, min(cfg["topk"][1], N), cfg["hidden"])
            if s["ref_g"] and cfg["global_enable"]:
                self.mem = [update_erase_memory(k1_t[len_l:].reshape(-1, cfg["hidden"]), self.mem[0], cfg["mem_size"]),
                            update_erase_memory(k2_t[len_l:].reshape(-1, cfg["hidden"]), self.mem[1], cfg["mem_size2"])]
            if s["frame_category"] == 0:
                kl = cfg["key_frame_location"]
                fd = fid - s["start_id"]
                fill = [0] * (kl - fd) + list(range(len_l)) + \
                       [len_l - 1] * (cfg["all_frame_interval"] - ((kl - fd) + len_l))
            else:
                fill = range(len_l)
            for i in fill:
                self.feats.append([feats_t[l][i:i + 1] for l in range(3)])
                self.cache.append((logits_t[i:i + 1], boxes_t[i:i + 1], obj_t[i]))

        batch = min(ib, s["end_id"] - fid + 1)
        r0 = cfg["key_frame_location"]
        feats_cur = [torch.cat([self.feats[i][l] for i in range(r0, r0 + batch)]) for l in range(3)]
        cached = (torch.cat([self.cache[i][0] for i in range(r0, r0 + batch)]),
                  torch.cat([self.cache[i][1] for i in range(r0, r0 + batch)]),
                  torch.cat([self.cache[i][2] for i in range(r0, r0 + batch)]))
        whwh = whwh1[None].expand(batch, -1)
        T = cfg["sample_step"]
        times = torch.linspace(-1, 999, steps=T + 1)
        times = list(reversed(times.int().tolist()))
        pairs = list(zip(times[:-1], times[1:]))
        img = self.noise.get("img", self.video, fid, 0, batch).to(dev)
        ens = []
        logits = coord = None
        for si, (time, time_next) in enumerate(pairs):
            t = torch.full((batch,), time, dtype=torch.long, device=dev)
            temb = time_embedding(c, t)
            if T > 1:
                lg, bx, obj = head_base_stages(c, feats_cur, self._x_to_boxes(img, whwh), temb, cfg)
            else:
                lg, bx, obj = cached
            if cfg["global_enable"] and cfg["num_heads_local"] > 0:
                attn_ = global_attention(c, obj, self.mem[0], cfg)
                for hi in range(cfg["num_heads_local"]):
                    lg, bx, obj = rcnn_head(c, "head.head_series_cond.%d." % hi, feats_cur, bx, obj, temb, cfg,
                                            cond=attn_)
            logits, coord = lg, bx
            self.trace[("logits", fid, si)] = logits
            self.trace[("coord", fid, si)] = coord
            x_start = self._x_start(coord, whwh)
            ac = self.alphas_cumprod[time]
            pred_noise = (torch.sqrt(1. / ac) * img - x_start) / torch.sqrt(1. / ac - 1)
            keep = torch.sigmoid(logits).max(dim=-1)[0] > 0.5                      # :559-565
            if time_next < 0:
                break
            a = self.alphas_cumprod[time].to(torch.float64)
            an = self.alphas_cumprod[time_next].to(torch.float64)
            sigma = (1.0 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()).to(torch.float32)
            cc = (1 - an - ((1 - a / an) * (1 - an) / (1 - a))).sqrt().to(torch.float32)
            san = self.alphas_cumprod[time_next].sqrt()
            eps = self.noise.get("eps", self.video, fid, si, batch).to(dev)
            fillz = self.noise.get("fill", self.video, fid, si, batch).to(dev)
            new = []
            for i in range(batch):
                kx = x_start[i, keep[i]]
                kn = pred_noise[i, keep[i]]
                nk = kx.shape[0]
                upd = kx * san + cc * kn + sigma * eps[i, :nk]
                new.append(torch.cat((upd, fillz[i, :N - nk]), dim=0))
            img = torch.stack(new)
            self.trace[("img", fid, si)] = img
            if T > 1:
                ens.append([topk_scores(logits[i], coord[i], N) for i in range(batch)])
        results = []
        for i in range(batch):
            if T > 1:
                bx = torch.cat([e[i][0] for e in ens]); sc = torch.cat([e[i][1] for e in ens])
                lb = torch.cat([e[i][2] for e in ens])
            else:
                bx, sc, lb = topk_scores(logits[i], coord[i], N)
            results.append(finalize_frame(bx, sc, lb, (w, h), cfg["use_nms"]))
        return results

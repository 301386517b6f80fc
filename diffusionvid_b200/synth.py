"""Deterministic synthetic weights and clips for the DiffusionVID hot path (no checkpoints or datasets are reachable).

State-dict keys are the reference's: detectron2 ResNet/FPN names under `backbone.`, DynamicHead names
under `head.` (mega_core/modeling/roi_heads/box_head/box_head.py:155-237,438-493,593-603,666-685).
Init follows DynamicHead._reset_parameters (box_head.py:239-248: xavier_uniform on every dim>1 parameter) except that
`class_logits.bias` is shifted from -log(99) so that roughly a third of the boxes score above the 0.5 renewal
threshold (diffusion_det.py:561); otherwise the DDIM update branch would never be exercised.
"""
import math

import torch

R101_BLOCKS = (3, 4, 23, 3)
R50_BLOCKS = (3, 4, 6, 3)


def _xavier(g, *shape):
    fan_out, fan_in = shape[0], shape[1]
    rf = 1
    for s in shape[2:]:
        rf *= s
    a = math.sqrt(6.0 / ((fan_in + fan_out) * rf))
    return (torch.rand(*shape, generator=g) * 2 - 1) * a


def _linear(sd, g, name, out_f, in_f, bias=True, bias_scale=0.02):
    sd[name + ".weight"] = _xavier(g, out_f, in_f)
    if bias:
        sd[name + ".bias"] = torch.randn(out_f, generator=g) * bias_scale


def _ln(sd, g, name, d):
    sd[name + ".weight"] = 1.0 + 0.1 * torch.randn(d, generator=g)
    sd[name + ".bias"] = 0.05 * torch.randn(d, generator=g)


def _conv_bn(sd, g, name, cout, cin, k, gamma_lo=0.5, gamma_hi=1.5):
    fan_in = cin * k * k
    sd[name + ".weight"] = torch.randn(cout, cin, k, k, generator=g) * math.sqrt(2.0 / fan_in)
    sd[name + ".norm.weight"] = gamma_lo + (gamma_hi - gamma_lo) * torch.rand(cout, generator=g)
    sd[name + ".norm.bias"] = 0.1 * torch.randn(cout, generator=g)
    sd[name + ".norm.running_mean"] = 0.1 * torch.randn(cout, generator=g)
    sd[name + ".norm.running_var"] = 0.5 + torch.rand(cout, generator=g)


def backbone_state_dict(g, blocks=R101_BLOCKS, fpn_ch=256):
    sd = {}
    p = "backbone.bottom_up."
    _conv_bn(sd, g, p + "stem.conv1", 64, 3, 7)
    cin = 64
    for si, nb in enumerate(blocks):
        mid = 64 * 2 ** si
        cout = 4 * mid
        for bi in range(nb):
            b = "%sres%d.%d." % (p, si + 2, bi)
            if bi == 0:
                _conv_bn(sd, g, b + "shortcut", cout, cin, 1, 0.7, 1.0)
            _conv_bn(sd, g, b + "conv1", mid, cin, 1)
            _conv_bn(sd, g, b + "conv2", mid, mid, 3)
            _conv_bn(sd, g, b + "conv3", cout, mid, 1, 0.1, 0.3)   # small residual-branch gain keeps fp16 in range
            cin = cout
    for lvl, c in ((3, 512), (4, 1024), (5, 2048)):
        for kind, k, ci in (("lateral", 1, c), ("output", 3, fpn_ch)):
            name = "backbone.fpn_%s%d" % (kind, lvl)
            sd[name + ".weight"] = _xavier(g, fpn_ch, ci, k, k)
            sd[name + ".bias"] = 0.02 * torch.randn(fpn_ch, generator=g)
    return sd


def swin_state_dict(g, embed=128, depths=(2, 2, 18, 2), heads=(4, 8, 16, 32), ws=7, fpn_ch=256):
    """Swin backbone + FPN keys of build_swintransformer_fpn_backbone (mega_core/modeling/backbone/swintransformer.py:
    464-751): trunc-normal(.02)-like Linear weights, small biases, LayerNorm near identity; FPN on swin1..3."""
    sd = {}
    p = "backbone.bottom_up."
    sd[p + "patch_embed.proj.weight"] = torch.randn(embed, 3, 4, 4, generator=g) * math.sqrt(1.0 / 48)
    sd[p + "patch_embed.proj.bias"] = 0.02 * torch.randn(embed, generator=g)
    _ln(sd, g, p + "patch_embed.norm", embed)
    for i, (d, nh) in enumerate(zip(depths, heads)):
        C = embed * 2 ** i
        for b in range(d):
            pre = "%slayers.%d.blocks.%d." % (p, i, b)
            _ln(sd, g, pre + "norm1", C)
            sd[pre + "attn.qkv.weight"] = torch.randn(3 * C, C, generator=g) * math.sqrt(1.0 / C)
            sd[pre + "attn.qkv.bias"] = 0.02 * torch.randn(3 * C, generator=g)
            sd[pre + "attn.proj.weight"] = torch.randn(C, C, generator=g) * (0.5 * math.sqrt(1.0 / C))
            sd[pre + "attn.proj.bias"] = 0.02 * torch.randn(C, generator=g)
            sd[pre + "attn.relative_position_bias_table"] = 0.5 * torch.randn((2 * ws - 1) ** 2, nh, generator=g)
            _ln(sd, g, pre + "norm2", C)
            sd[pre + "mlp.fc1.weight"] = torch.randn(4 * C, C, generator=g) * math.sqrt(1.0 / C)
            sd[pre + "mlp.fc1.bias"] = 0.02 * torch.randn(4 * C, generator=g)
            sd[pre + "mlp.fc2.weight"] = torch.randn(C, 4 * C, generator=g) * (0.5 * math.sqrt(1.0 / (4 * C)))
            sd[pre + "mlp.fc2.bias"] = 0.02 * torch.randn(C, generator=g)
        if i < len(depths) - 1:
            pre = "%slayers.%d.downsample." % (p, i)
            _ln(sd, g, pre + "norm", 4 * C)
            sd[pre + "reduction.weight"] = torch.randn(2 * C, 4 * C, generator=g) * math.sqrt(1.0 / (4 * C))
        if i >= 1:
            _ln(sd, g, "%snorm%d" % (p, i), C)
    for lvl, i in ((3, 1), (4, 2), (5, 3)):
        c = embed * 2 ** i
        for kind, k, ci in (("lateral", 1, c), ("output", 3, fpn_ch)):
            name = "backbone.fpn_%s%d" % (kind, lvl)
            sd[name + ".weight"] = _xavier(g, fpn_ch, ci, k, k)
            sd[name + ".bias"] = 0.02 * torch.randn(fpn_ch, generator=g)
    return sd


def _rcnn_head(sd, g, pre, d, dd, ff, ncls, num_cls, num_reg, cond, cls_bias, pool=7):
    sd[pre + "self_attn.in_proj_weight"] = _xavier(g, 3 * d, d)
    sd[pre + "self_attn.in_proj_bias"] = 0.02 * torch.randn(3 * d, generator=g)
    _linear(sd, g, pre + "self_attn.out_proj", d, d)
    _linear(sd, g, pre + "inst_interact.dynamic_layer", 2 * d * dd, d)
    _ln(sd, g, pre + "inst_interact.norm1", dd)
    _ln(sd, g, pre + "inst_interact.norm2", d)
    _linear(sd, g, pre + "inst_interact.out_layer", d, d * pool * pool)
    _ln(sd, g, pre + "inst_interact.norm3", d)
    _linear(sd, g, pre + "linear1", ff, d)
    _linear(sd, g, pre + "linear2", d, ff)
    for n in ("norm1", "norm2", "norm3"):
        _ln(sd, g, pre + n, d)
    _linear(sd, g, pre + "block_time_mlp.1", d if cond else 2 * d, 4 * d)
    if cond:
        _linear(sd, g, pre + "c_mlp.1", d, d)
    for i in range(num_cls):
        _linear(sd, g, pre + "cls_module.%d" % (3 * i), d, d, bias=False)
        _ln(sd, g, pre + "cls_module.%d" % (3 * i + 1), d)
    for i in range(num_reg):
        _linear(sd, g, pre + "reg_module.%d" % (3 * i), d, d, bias=False)
        _ln(sd, g, pre + "reg_module.%d" % (3 * i + 1), d)
    _linear(sd, g, pre + "class_logits", ncls, d)
    sd[pre + "class_logits.bias"] = torch.full((ncls,), cls_bias)
    _linear(sd, g, pre + "bboxes_delta", 4, d)
    sd[pre + "bboxes_delta.weight"] *= 0.25     # keep refined boxes near their inputs, as a trained head does


def head_state_dict(g, num_heads=3, num_heads_local=1, d=256, dd=64, ff=2048, ncls=30, num_cls=1, num_reg=3,
                    global_enable=True, cls_bias=-1.9):
    sd = {}
    _linear(sd, g, "head.time_mlp.1", 4 * d, d)
    _linear(sd, g, "head.time_mlp.3", 4 * d, 4 * d)
    for i in range(num_heads):
        _rcnn_head(sd, g, "head.head_series.%d." % i, d, dd, ff, ncls, num_cls, num_reg, False, cls_bias)
    for i in range(num_heads_local):
        _rcnn_head(sd, g, "head.head_series_cond.%d." % i, d, dd, ff, ncls, num_cls, num_reg, True, cls_bias)
    if global_enable:
        pre = "head.global_attention.0.0."
        sd[pre + "in_proj_weight"] = _xavier(g, 3 * d, d)
        sd[pre + "in_proj_bias"] = 0.02 * torch.randn(3 * d, generator=g)
        _linear(sd, g, pre + "out_proj", d, d)
    return sd


def make_state_dict(seed=1234, blocks=R101_BLOCKS, swin=None, **head_kw):
    """swin: None for the R-x+FPN backbone, else dict(embed=, depths=, heads=) for Swin + FPN."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    sd = swin_state_dict(g, **swin) if swin else backbone_state_dict(g, blocks)
    sd.update(head_state_dict(g, **head_kw))
    return sd


def make_clip(num_frames, h=600, w=1000, seed=1234, pad_to=32):
    """Seeded smooth synthetic frames in [0,1] (low-pass noise drifting over time), zero-padded to a multiple of
    `pad_to` like BatchCollator/to_image_list (mega_core/data/collate_batch.py:17-41).  Returns (F,3,Hp,Wp)."""
    g = torch.Generator(device="cpu").manual_seed(seed)
    hp = (h + pad_to - 1) // pad_to * pad_to
    wp = (w + pad_to - 1) // pad_to * pad_to
    lo_h, lo_w = max(2, h // 40), max(2, w // 40)
    base = torch.rand(1, 3, lo_h, lo_w, generator=g)
    drift = torch.rand(num_frames, 3, lo_h, lo_w, generator=g)
    t = torch.linspace(0, 1, num_frames).view(-1, 1, 1, 1)
    lo = 0.7 * base + 0.3 * ((1 - t) * drift[:1] + t * drift)
    up = torch.nn.functional.interpolate(lo, size=(h, w), mode="bilinear", align_corners=False)
    fine = torch.rand(num_frames, 3, max(2, h // 6), max(2, w // 6), generator=g)
    up = (0.8 * up + 0.2 * torch.nn.functional.interpolate(fine, size=(h, w), mode="bilinear",
                                                            align_corners=False)).clamp(0, 1)
    out = torch.zeros(num_frames, 3, hp, wp)
    out[:, :, :h, :w] = up
    return out


def clip_samples(frames, global_idx, h, w, infer_batch=8, max_offset=7, video_id=0):
    """The per-frame sample dicts VIDMEGADataset._get_test builds (mega_core/data/datasets/vid_mega.py:164-250) for
    one video: frame 0 carries ref_l = frames [0..max_offset] and the global frames; frame f>0 carries the single
    frame min(f+max_offset, L-1).  `frames` (L,3,Hp,Wp); `global_idx` explicit (SURVEY.md 8c contract 2)."""
    L = frames.shape[0]
    out = []
    for f in range(L):
        if f == 0:
            ref_l = [frames[i][None] for i in range(0, min(max_offset, L - 1) + 1)]
            ref_g = [frames[i][None] for i in global_idx]
        else:
            ref_l = [frames[min(f + max_offset, L - 1)][None]]
            ref_g = []
        out.append(dict(cur=frames[f][None], image_size=(h, w), ref_l=ref_l, ref_g=ref_g, frame_id=f, start_id=0,
                        end_id=L - 1, seg_len=L, last_queue_id=min(f + max_offset, L - 1),
                        frame_category=0 if f == 0 else 1, video_id=video_id))
    return out

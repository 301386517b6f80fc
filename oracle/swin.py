"""CPU restatement of the Swin Transformer backbone + FPN of vid_Swin_B_DiffusionVID.yaml.  TEST INFRASTRUCTURE ONLY.

Restates mega_core/modeling/backbone/swintransformer.py: PatchEmbed :422-461, WindowAttention :98-176,
SwinTransformerBlock :179-276 (LN -> zero pad to x7 -> cyclic shift -> window partition -> W-MSA with relative position
bias and shift mask -> reverse -> un-shift -> crop -> residual -> LN -> MLP(GELU) -> residual), PatchMerging :279-317,
BasicLayer :320-419 (SW-MSA mask built on the padded grid), SwinTransformer.forward :621-648 (per-output LayerNorm),
build_swintransformer_fpn_backbone :735-751 (detectron2 FPN on swin1..3, restated in oracle.model / SURVEY.md A1).
Pinned against the reference file itself by tests/golden/make_golden_swin.py + tests/test_golden_reference.py.

Functional over a state dict with the reference's key names under `backbone.bottom_up.`; `q` is oracle.model.Quant:
in fp16-emulation mode values are rounded to fp16 exactly where the sm_100a path stores fp16 (LayerNorm outputs that
feed GEMMs, GEMM outputs, attention output, output feature maps); the residual stream stays fp32.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from .model import Ctx, Quant, _conv_bias  # noqa: F401


def relative_position_index(ws):
    """swintransformer.py:124-135."""
    coords = torch.stack(torch.meshgrid([torch.arange(ws), torch.arange(ws)], indexing="ij"))
    cf = torch.flatten(coords, 1)
    rel = (cf[:, :, None] - cf[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def shift_mask(H, W, ws, shift):
    """BasicLayer.forward :387-406: (nW, ws*ws, ws*ws) with 0 / -100."""
    Hp = int(np.ceil(H / ws)) * ws
    Wp = int(np.ceil(W / ws)) * ws
    img = torch.zeros((1, Hp, Wp, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    mw = window_partition(img, ws).view(-1, ws * ws)
    am = mw.unsqueeze(1) - mw.unsqueeze(2)
    return am.masked_fill(am != 0, -100.0).masked_fill(am == 0, 0.0)


def window_partition(x, ws):
    B, H, W, C = x.shape
    x = x.view(B, H // ws, ws, W // ws, ws, C)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(-1, ws, ws, C)


def window_reverse(win, ws, H, W):
    B = int(win.shape[0] / (H * W / ws / ws))
    x = win.view(B, H // ws, W // ws, ws, ws, -1)
    return x.permute(0, 1, 3, 2, 4, 5).contiguous().view(B, H, W, -1)


def _ln(x, sd, name):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], 1e-5)


def swin_block(c, pre, x, H, W, heads, ws, shift, mask):
    """x (B, H*W, C) fp32 residual stream -> same."""
    sd = c.sd
    B, L, C = x.shape
    hd = C // heads
    xn = c.q.a(_ln(x, sd, pre + "norm1")).view(B, H, W, C)
    pad_r = (ws - W % ws) % ws
    pad_b = (ws - H % ws) % ws
    xn = F.pad(xn, (0, 0, 0, pad_r, 0, pad_b))
    Hp, Wp = H + pad_b, W + pad_r
    if shift > 0:
        xn = torch.roll(xn, shifts=(-shift, -shift), dims=(1, 2))
    xw = window_partition(xn, ws).view(-1, ws * ws, C)
    Bw, N = xw.shape[:2]
    qkv = c.lin(xw.reshape(-1, C), pre + "attn.qkv", out16=True).view(Bw, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q @ k.transpose(-2, -1)
    table = sd[pre + "attn.relative_position_bias_table"]
    bias = table[relative_position_index(ws).view(-1).to(table.device)].view(N, N, -1).permute(2, 0, 1).contiguous()
    attn = attn + bias.unsqueeze(0)
    if shift > 0:
        nW = mask.shape[0]
        attn = (attn.view(Bw // nW, nW, heads, N, N) + mask.unsqueeze(1).unsqueeze(0)).view(-1, heads, N, N)
    attn = torch.softmax(attn, dim=-1)
    o = c.q.a((attn @ v).transpose(1, 2).reshape(Bw * N, C))
    o = c.lin(o, pre + "attn.proj", out16=True).view(-1, ws, ws, C)
    xs = window_reverse(o, ws, Hp, Wp)
    if shift > 0:
        xs = torch.roll(xs, shifts=(shift, shift), dims=(1, 2))
    xs = xs[:, :H, :W, :].contiguous().view(B, H * W, C)
    x = x + xs
    y = c.q.a(_ln(x, sd, pre + "norm2"))
    y = c.q.a(F.gelu(c.lin(y.reshape(-1, C), pre + "mlp.fc1")))
    y = c.lin(y, pre + "mlp.fc2", out16=True).view(B, L, C)
    return x + y


def patch_merging(c, pre, x, H, W):
    sd = c.sd
    B, L, C = x.shape
    x = x.view(B, H, W, C)
    if H % 2 == 1 or W % 2 == 1:
        x = F.pad(x, (0, 0, 0, W % 2, 0, H % 2))
    x = torch.cat([x[:, 0::2, 0::2, :], x[:, 1::2, 0::2, :], x[:, 0::2, 1::2, :], x[:, 1::2, 1::2, :]], -1)
    x = x.view(B, -1, 4 * C)
    x = c.q.a(_ln(x, sd, pre + "norm"))
    return c.lin(x.reshape(-1, 4 * C), pre + "reduction", bias=False, out16=True).view(B, -1, 2 * C)


def swin_config(sd, p="backbone.bottom_up."):
    """(embed_dim, depths, heads, window) recovered from the state dict (heads from the bias table's last dim)."""
    embed = sd[p + "patch_embed.proj.weight"].shape[0]
    depths, heads = [], []
    s = 0
    while (p + "layers.%d.blocks.0.norm1.weight" % s) in sd:
        d = 0
        while (p + "layers.%d.blocks.%d.norm1.weight" % (s, d)) in sd:
            d += 1
        depths.append(d)
        t = sd[p + "layers.%d.blocks.0.attn.relative_position_bias_table" % s]
        heads.append(t.shape[1])
        ws = (int(math.isqrt(t.shape[0])) + 1) // 2
        s += 1
    return embed, depths, heads, ws


def swin_body(c, x, out_indices=(1, 2, 3), p="backbone.bottom_up."):
    """SwinTransformer.forward: normalised (B,3,H,W) -> {'swin<i>': (B,C_i,H_i,W_i)}."""
    sd = c.sd
    embed, depths, heads, ws = swin_config(sd, p)
    _, _, H, W = x.shape
    if W % 4:
        x = F.pad(x, (0, 4 - W % 4))
    if H % 4:
        x = F.pad(x, (0, 0, 0, 4 - H % 4))
    x = F.conv2d(c.q.a(x), c.q.w(sd[p + "patch_embed.proj.weight"]), sd[p + "patch_embed.proj.bias"], stride=4)
    x = c.q.a(x)
    Wh, Ww = x.shape[2], x.shape[3]
    x = x.flatten(2).transpose(1, 2)
    x = _ln(x, sd, p + "patch_embed.norm")
    outs = {}
    for i, (depth, nh) in enumerate(zip(depths, heads)):
        mask = shift_mask(Wh, Ww, ws, ws // 2).to(x.device)
        for b in range(depth):
            x = swin_block(c, "%slayers.%d.blocks.%d." % (p, i, b), x, Wh, Ww, nh, ws, 0 if b % 2 == 0 else ws // 2,
                           mask)
        if i in out_indices:
            o = c.q.a(_ln(x, sd, "%snorm%d" % (p, i)))
            outs["swin%d" % i] = o.view(-1, Wh, Ww, o.shape[-1]).permute(0, 3, 1, 2).contiguous()
        if i < len(depths) - 1:
            x = patch_merging(c, "%slayers.%d.downsample." % (p, i), x, Wh, Ww)
            Wh, Ww = (Wh + 1) // 2, (Ww + 1) // 2
    return outs


def swin_fpn(c, x):
    """build_swintransformer_fpn_backbone: FPN(in_features swin1..3, 256 ch, sum fuse) -> [p3, p4, p5]."""
    o = swin_body(c, x)
    prev = _conv_bias(c, o["swin3"], "backbone.fpn_lateral5", 0)
    p5 = _conv_bias(c, prev, "backbone.fpn_output5", 1)
    prev = _conv_bias(c, o["swin2"], "backbone.fpn_lateral4", 0,
                      resid=F.interpolate(prev, scale_factor=2.0, mode="nearest"))
    p4 = _conv_bias(c, prev, "backbone.fpn_output4", 1)
    prev = _conv_bias(c, o["swin1"], "backbone.fpn_lateral3", 0,
                      resid=F.interpolate(prev, scale_factor=2.0, mode="nearest"))
    p3 = _conv_bias(c, prev, "backbone.fpn_output3", 1)
    return [p3, p4, p5]

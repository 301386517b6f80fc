"""TEST-ONLY stand-in for diffusionvid_b200.ops on machines without a GPU.

Re-implements the semantics of every op wrapper with plain PyTorch CPU code (fp16 storage emulated with .half()), so
the *host-side* logic of the product - weight packing, the clip state machine, the DDIM loop, the multi-rank frame
sharding - can be exercised against the oracle under `pytest -m "not gpu"`.  It is installed by monkeypatching in
tests only; the product never imports it and keeps failing loudly without CUDA (tests/test_host_logic.py checks that).
"""
import math

import torch
import torch.nn.functional as F

from oracle import model as om
from oracle import ops as oo

H = torch.float16
F32 = torch.float32
LAUNCHES = 0


def require_device(dev):
    return None


def _nhwc_to_nchw(x):
    return x.float().permute(0, 3, 1, 2)


def conv2d(x, w, bias, cout, R, S, stride, pad, relu, resid=None, resid_shift=0, out=None):
    n, h, wd, cin = x.shape
    wt = w.float().view(cout, R, S, cin).permute(0, 3, 1, 2)
    y = F.conv2d(_nhwc_to_nchw(x), wt, bias, stride=stride, padding=pad)
    if resid is not None:
        r = _nhwc_to_nchw(resid)
        if resid_shift:
            r = F.interpolate(r, scale_factor=2.0 ** resid_shift, mode="nearest")[:, :, :y.shape[2], :y.shape[3]]
        y = y + r
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().half()


def stem_conv(x_haloed, w, bias, n, H_, W_, cout, relu=True, out=None):
    wk = w.float().view(cout, 7, 8, 8)[:, :, :7, :3].permute(0, 3, 1, 2)
    img = _nhwc_to_nchw(x_haloed)[:, :3, 3:-3, 3:-3]
    y = F.conv2d(img, wk, bias, stride=2, padding=3)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous().half()


def gemm(a, w, bias=None, relu=False, resid=None, out=None):
    y = F.linear(a.float(), w.float(), bias)
    if resid is not None:
        y = y + resid.float()
    if relu:
        y = F.relu(y)
    return y.half()


def gemm_partials(a, w, splits=1, out=None):
    return F.linear(a.float(), w.float())[None].contiguous(), 1


def preprocess(img, mean, std, halo=3):
    m = torch.tensor(mean, dtype=F32).view(1, 3, 1, 1)
    s = torch.tensor(std, dtype=F32).view(1, 3, 1, 1)
    if img.dtype == torch.uint8:          # clip-loader mode: ToTensor (u8 / 255) fused into the kernel
        img = img.to(F32).div(255)
    x = ((img - m) / s).half()
    n, _, h, w = img.shape
    out = torch.zeros((n, h + 2 * halo, w + 2 * halo, 8), dtype=H)
    out[:, halo:halo + h, halo:halo + w, :3] = x.permute(0, 2, 3, 1)
    return out


def maxpool3x3s2(x):
    return F.max_pool2d(_nhwc_to_nchw(x), 3, 2, 1).permute(0, 2, 3, 1).contiguous().half()


def attention(q, k, v, out, batch, heads, lq, lk, q_rs, k_rs, v_rs, o_rs, q_bs, k_bs, v_bs, o_bs, tc=None):
    def view(t, L, rs, bs):
        base = t.reshape(-1) if t.is_contiguous() else None
        rows = []
        flat = torch.as_strided(t, (batch, L, heads * 32), (bs, rs, 1), t.storage_offset()) if base is None else \
            torch.as_strided(t, (batch, L, heads * 32), (bs, rs, 1), t.storage_offset())
        return flat.float().view(batch, L, heads, 32)
    qq, kk, vv = view(q, lq, q_rs, q_bs), view(k, lk, k_rs, k_bs), view(v, lk, v_rs, v_bs)
    att = torch.softmax(torch.einsum("blhd,bshd->bhls", qq, kk) / math.sqrt(32), dim=-1)
    ctx = torch.einsum("bhls,bshd->blhd", att, vv).reshape(batch, lq, heads * 32).half()
    o = torch.as_strided(out, (batch, lq, heads * 32), (o_bs, o_rs, 1), out.storage_offset())
    o.copy_(ctx)
    return out


class Levels:
    def __init__(self, feats, scales=(1 / 8., 1 / 16., 1 / 32.)):
        self.feats = feats
        self.scales = scales

    def nchw(self):
        return [_nhwc_to_nchw(f) for f in self.feats]


def roi_align(levels, boxes, boxes_per_frame, want_roi=True, want_mean=True):
    B = boxes.numel() // 4 // boxes_per_frame
    r = oo.roi_pooler(levels.nchw(), boxes.view(B, boxes_per_frame, 4), scales=levels.scales)
    roi = r.view(r.shape[0], 256, 49).permute(0, 2, 1).contiguous().half()
    mean = roi.float().mean(1)
    return (roi if want_roi else None), (mean if want_mean else None), (mean.half() if want_mean else None)


def dynconv_permutation(d=256, dd=64):
    i = torch.arange(d)[None, :]
    j = torch.arange(dd)[:, None]
    p1 = (i * dd + j).reshape(-1)
    p2 = d * dd + (torch.arange(dd)[None, :] * d + torch.arange(d)[:, None]).reshape(-1)
    return torch.cat([p1, p2])


def roi_dynconv(levels, boxes, boxes_per_frame, params, g1, b1, g2, b2, roi_in=None, out=None, transposed=False):
    if roi_in is None:
        roi_in, _, _ = roi_align(levels, boxes, boxes_per_frame, True, False)
    M = params.shape[0]
    if transposed:
        p1 = params[:, :16384].float().view(M, 64, 256).transpose(1, 2)
        p2 = params[:, 16384:].float().view(M, 256, 64).transpose(1, 2)
    else:
        p1 = params[:, :16384].float().view(M, 256, 64)
        p2 = params[:, 16384:].float().view(M, 64, 256)
    f = F.relu(F.layer_norm(torch.bmm(roi_in.float(), p1), (64,), g1, b1)).half().float()
    f = F.relu(F.layer_norm(torch.bmm(f, p2), (256,), g2, b2)).half().reshape(M, 49 * 256)
    if out is not None:
        out.copy_(f)
        return out
    return f


def gemm_row(a, w, bias=None, resid=None, ln=None, act=0, out_f32=None, out_f16=None):
    y = F.linear(a.float(), w.float(), bias)
    if resid is not None:
        y = y + resid
    if ln is not None:
        y = F.layer_norm(y, (256,), ln[0], ln[1])
    if out_f32 is not None:
        out_f32.copy_(y)
    if out_f16 is not None:
        out_f16.copy_((F.relu(y) if act == 1 else F.silu(y) if act == 2 else y).half())


def row_post(M, partials=None, splits=1, in_f16=None, bias=None, ln1=None, relu1=False, resid=None, ln2=None, act2=0,
             act2_f16_only=False, out_f32=None, out_f16=None, mod_scale=None, mod_shift=None, rows_per_group=1,
             scale_stride=0, shift_stride=0, shift_per_row=False, out_mod_f16=None):
    v = partials.view(-1, M, 256)[:splits].sum(0) if partials is not None else in_f16.float()
    if bias is not None:
        v = v + bias
    if ln1 is not None:
        v = F.layer_norm(v, (256,), ln1[0], ln1[1])
    if relu1:
        v = F.relu(v)
    if resid is not None:
        v = v + resid
    if ln2 is not None:
        v = F.layer_norm(v, (256,), ln2[0], ln2[1])
    y = F.relu(v) if act2 == 1 else (F.silu(v) if act2 == 2 else v)
    if not act2_f16_only:
        v = y
    if out_f32 is not None:
        out_f32.copy_(v)
    if out_f16 is not None:
        out_f16.copy_(y.half())
    if out_mod_f16 is not None:
        grp = torch.arange(M) // rows_per_group
        sc = mod_scale.reshape(-1)[(grp * scale_stride)[:, None] + torch.arange(256)[None]] if scale_stride else \
            mod_scale.reshape(-1)[:256][None].expand(M, 256)
        if shift_per_row:
            sh = mod_shift
        else:
            flat = torch.as_strided(mod_shift, (int(grp.max()) + 1, 256), (shift_stride, 1), mod_shift.storage_offset())
            sh = flat[grp]
        out_mod_f16.copy_((v * (sc + 1) + sh).half())


def small_linear(a, w, bias, act_in=0, act_out=0):
    x = F.silu(a) if act_in == 1 else a
    y = F.linear(x, w.float(), bias)
    return F.gelu(y) if act_out == 1 else y


def time_sinusoid(t, freq):
    e = t[:, None] * freq[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def head_tail(fc16, cls, reg, logit_w, logit_b, C, delta_w, delta_b, boxes_in):
    def tower(x, w, ln):
        y = F.linear(x.float(), w.float())
        return F.relu(F.layer_norm(y, (256,), ln[0], ln[1], 1e-5)).half()
    c = tower(fc16, cls[0], cls[1])
    logits = F.linear(c.float(), logit_w.float()[:C], logit_b)
    r = fc16
    for w, ln in reg:
        r = tower(r, w, ln)
    deltas = F.linear(r.float(), delta_w.float()[:4], delta_b)
    return logits, om.apply_deltas(deltas, boxes_in)


def head_final(logit_part, cls_bias, C, delta_part, delta_bias, boxes_in, logits_out=None, boxes_out=None):
    return logit_part[:, :C] + cls_bias, om.apply_deltas(delta_part[:, :4] + delta_bias, boxes_in)


def noise_to_boxes(x, scale, W, Hh):
    xb = ((torch.clamp(x, -scale, scale) / scale) + 1) / 2
    return om.box_cxcywh_to_xyxy(xb) * torch.tensor([W, Hh, W, Hh])


def ddim_step(logits, coord, x_t, eps, fill, scale, W, Hh, sqrt_recip_a, sqrt_recipm1_a, sqrt_a_next, c_coef, sigma):
    frames, N, C = logits.shape
    whwh = torch.tensor([W, Hh, W, Hh])
    xs = torch.clamp((om.box_xyxy_to_cxcywh(coord / whwh) * 2 - 1.) * scale, -scale, scale)
    pn = (torch.tensor(sqrt_recip_a) * x_t - xs) / torch.tensor(sqrt_recipm1_a)
    keep = torch.sigmoid(logits).max(-1)[0] > 0.5
    new = []
    for i in range(frames):
        nk = int(keep[i].sum())
        upd = xs[i, keep[i]] * sqrt_a_next + c_coef * pn[i, keep[i]] + sigma * eps[i, :nk]
        new.append(torch.cat((upd, fill[i, :N - nk]), 0))
    xn = torch.stack(new)
    return xn, noise_to_boxes(xn, scale, W, Hh), keep.sum(-1).int()


def topk_scores(logits, boxes, k, out_boxes, out_scores, out_labels, slot0):
    for i in range(logits.shape[0]):
        b, s, l = om.topk_scores(logits[i], boxes[i], k)
        out_boxes[i, slot0:slot0 + k] = b
        out_scores[i, slot0:slot0 + k] = s
        out_labels[i, slot0:slot0 + k] = l.int()


def topk_mask(logits, k1, k2):
    mx = logits.max(-1)[0]
    order = torch.sort(mx, dim=-1, descending=True, stable=True)[1]
    m1 = torch.zeros_like(mx, dtype=torch.uint8).scatter_(1, order[:, :k1], 1)
    m2 = torch.zeros_like(mx, dtype=torch.uint8).scatter_(1, order[:, :k2], 1)
    return m1, m2


def gather_masked_rows(src, mask, k):
    frames, N = mask.shape
    return src.view(frames, N, 256)[mask.bool()].contiguous()


def nms(boxes, scores, labels=None, counts=None, n=None, thr=0.5, plus_one=False, ge=False, ascending_out=False,
        clip_wh=None, want_compact=True):
    frames, cap = scores.shape
    keep = torch.full((frames, cap), -1, dtype=torch.int64)
    count = torch.zeros((frames,), dtype=torch.int32)
    ob = torch.zeros((frames, cap, 4)); os_ = torch.zeros((frames, cap)); ol = torch.zeros((frames, cap), dtype=torch.int32)
    for i in range(frames):
        c = cap if counts is None else int(counts[i])
        c = c if n is None else min(c, n)
        k = oo.batched_nms(boxes[i, :c], scores[i, :c], labels[i, :c], thr) if labels is not None else \
            oo.nms(boxes[i, :c], scores[i, :c], thr)
        count[i] = k.numel()
        keep[i, :k.numel()] = k
        b = boxes[i][k].clone()
        if clip_wh is not None:
            b[:, 0].clamp_(0, clip_wh[0] - 1); b[:, 1].clamp_(0, clip_wh[1] - 1)
            b[:, 2].clamp_(0, clip_wh[0] - 1); b[:, 3].clamp_(0, clip_wh[1] - 1)
        ob[i, :k.numel()] = b
        os_[i, :k.numel()] = scores[i][k]
        if labels is not None:
            ol[i, :k.numel()] = labels[i][k]
    return dict(keep=keep, count=count, boxes=ob, scores=os_, labels=ol)


def cdist(x):
    return oo.cdist_l2(x)


def furthest_point_sampling(b, n, m, dist, temp, idx):
    idx[0].copy_(torch.from_numpy(oo.fps(dist.numpy(), m)))
    return 1

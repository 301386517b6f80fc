"""Torch-tensor front ends of the C ABI (include/dvid_b200.h).  PyTorch is used for device memory and streams only;
every function launches hand-written sm_100a kernels from libdvid_b200.so on the current CUDA stream and raises
DvidError when the library is missing or a call fails (no fallback)."""
import ctypes

import torch

from . import _lib
from ._lib import check, cur_stream, ptr

LAUNCHES = 0   # number of kernel launches issued through this module (bench.py reports it as gpu_launches)
PROFILE = None  # when a dict: family -> [list of (start_event, end_event), flops, bytes]; filled by _prof (bench.py)


def _cnt(n=1):
    global LAUNCHES
    LAUNCHES += n


class _prof:
    """Context manager: when ops.PROFILE is a dict, bracket one launch with CUDA events on the launching stream and
    account its algorithmic FLOPs / bytes under `family` (bench.py's live per-kernel roofline measurement)."""

    def __init__(self, family, flops=0.0, nbytes=0.0, tag=None):
        self.family, self.flops, self.nbytes = family, flops, nbytes
        self.tag = tag

    def __enter__(self):
        if PROFILE is not None:
            self.s = torch.cuda.Event(enable_timing=True)
            self.e = torch.cuda.Event(enable_timing=True)
            self.s.record()
        return self

    def __exit__(self, *exc):
        if PROFILE is not None:
            self.e.record()
            rec = PROFILE.setdefault(self.family, [[], 0.0, 0.0])
            rec[0].append((self.s, self.e))
            rec[1] += self.flops
            rec[2] += self.nbytes
            if self.tag is not None:      # per-shape breakdown (tools/profile_shapes.py)
                rec = PROFILE.setdefault(self.family + ":" + self.tag, [[], 0.0, 0.0])
                rec[0].append((self.s, self.e))
                rec[1] += self.flops
                rec[2] += self.nbytes
        return False


def _chk(t, dtype, name):
    if t is None:
        return
    if not t.is_cuda or t.dtype != dtype or not t.is_contiguous():
        raise _lib.DvidError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} "
                             f"cuda={t.is_cuda} contiguous={t.is_contiguous()}")


H = torch.float16
F32 = torch.float32


def require_device(dev):
    """The product path is CUDA-only: raise unless `dev` is a CUDA device and the native library loads."""
    if dev.type != "cuda" or not torch.cuda.is_available():
        raise _lib.DvidError("diffusionvid_b200 runs on CUDA (sm_100a) only; there is no CPU path")
    _lib.lib()


# ------------------------------------------------------------------------------------------------ dense
def conv2d(x, w, bias, cout, R, S, stride, pad, relu, resid=None, resid_shift=0, out=None):
    """x NHWC fp16, w [cout][R*S*cin] fp16 -> NHWC fp16."""
    _chk(x, H, "x"); _chk(w, H, "w"); _chk(bias, F32, "bias"); _chk(resid, H, "resid")
    n, h, wd, cin = x.shape
    ho = (h + 2 * pad - R) // stride + 1
    wo = (wd + 2 * pad - S) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), device=x.device, dtype=H)
    with _prof("conv_gemm", 2.0 * n * ho * wo * cout * R * S * cin,
               2.0 * (x.numel() + w.numel() + n * ho * wo * cout * (2 if resid is not None else 1)),
               tag="conv %dx%d s%d %d->%d @%dx%dx%d%s" % (R, S, stride, cin, cout, n, ho, wo, "+res" if resid is not None else "")):
        check(_lib.lib().dvid_conv2d_nhwc_f16(ptr(x), ptr(w), ptr(bias), ptr(resid), ptr(out), n, h, wd, cin, cout, R,
                                              S, stride, pad, resid_shift, int(relu), cur_stream()),
              "dvid_conv2d_nhwc_f16")
    _cnt()
    return out


def stem_conv(x_haloed, w, bias, n, H_, W_, cout, relu=True, out=None):
    _chk(x_haloed, H, "x"); _chk(w, H, "w"); _chk(bias, F32, "bias")
    if out is None:
        out = torch.empty((n, H_ // 2, W_ // 2, cout), device=x_haloed.device, dtype=H)
    with _prof("conv_gemm", 2.0 * n * (H_ // 2) * (W_ // 2) * cout * 147, 2.0 * (x_haloed.numel() + out.numel())):
        check(_lib.lib().dvid_stem_conv_f16(ptr(x_haloed), ptr(w), ptr(bias), ptr(out), n, H_, W_, cout, int(relu),
                                            cur_stream()), "dvid_stem_conv_f16")
    _cnt()
    return out


def gemm(a, w, bias=None, relu=False, resid=None, out=None):
    """fp16 out: act(a @ w^T + bias + resid); relu: False/0 none, True/1 ReLU, 2 GELU(erf)."""
    _chk(a, H, "a"); _chk(w, H, "w"); _chk(bias, F32, "bias"); _chk(resid, H, "resid")
    m, k = a.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((m, n), device=a.device, dtype=H)
    with _prof("conv_gemm", 2.0 * m * n * k, 2.0 * (m * k + n * k + m * n), tag="gemm %dx%d->%d" % (m, k, n)):
        check(_lib.lib().dvid_gemm_f16(ptr(a), ptr(w), ptr(bias), ptr(resid), ptr(out), None, m, n, k, int(relu), 1,
                                       None, cur_stream()), "dvid_gemm_f16")
    _cnt()
    return out


def gemm_partials(a, w, splits=1, out=None):
    """fp32 split-K partial sums [splits_used][m][n] (no bias); returns (partials, splits_used)."""
    _chk(a, H, "a"); _chk(w, H, "w")
    m, k = a.shape
    n = w.shape[0]
    if out is None:
        out = torch.empty((max(1, splits), m, n), device=a.device, dtype=F32)
    used = ctypes.c_int(0)
    with _prof("conv_gemm", 2.0 * m * n * k, 2.0 * (m * k + n * k) + 4.0 * m * n * max(1, splits),
               tag="gemm(f32 partials x%d) %dx%d->%d" % (splits, m, k, n)):
        check(_lib.lib().dvid_gemm_f16(ptr(a), ptr(w), None, None, None, ptr(out), m, n, k, 0, splits,
                                       ctypes.byref(used), cur_stream()), "dvid_gemm_f16(partials)")
    _cnt()
    return out, used.value


def conv_streamk(enable):
    """Stream-K scheduling of the conv/GEMM kernel for layers whose tile count leaves >= 1/4 of a wave idle
    (include/dvid_b200.h, dvid_conv_streamk).  Process-wide switch; the first enable allocates the workspace and must
    happen outside a stream capture.  Keep it off while convolutions run concurrently on several streams."""
    check(_lib.lib().dvid_conv_streamk(1 if enable else 0), "dvid_conv_streamk")


# ------------------------------------------------------------------------------------------------ image side
def preprocess(img, mean, std, halo=3):
    """img [n,3,H,W] fp32 in [0,1], or uint8 as decoded (the reference's ToTensor, transforms.py:295-297, is then
    evaluated inside the kernel) -> [n, H+2*halo, W+2*halo, 8] fp16 normalised, zero halo."""
    u8 = img.dtype == torch.uint8
    _chk(img, torch.uint8 if u8 else F32, "img")
    n, _, h, w = img.shape
    out = torch.empty((n, h + 2 * halo, w + 2 * halo, 8), device=img.device, dtype=H)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    fn = _lib.lib().dvid_preprocess_u8 if u8 else _lib.lib().dvid_preprocess
    check(fn(ptr(img), ptr(out), n, h, w, halo, h + 2 * halo, w + 2 * halo, m, s, cur_stream()), "dvid_preprocess")
    _cnt()
    return out


def resize_frames_u8(frames_hwc, oh, ow, pad_to=32):
    """Decoded frames [n, Hin, Win, 3] uint8 (PIL order) -> [n, 3, Hp, Wp] uint8, the Pillow-bilinear resize of the
    reference's test transform (transforms.py:31-67) in the top-left oh x ow corner, zero padding up to a multiple of
    `pad_to` (to_image_list).  Bytes equal Pillow's (dvid_resize_bilinear_u8)."""
    _chk(frames_hwc, torch.uint8, "frames_hwc")
    n, hin, win, c = frames_hwc.shape
    if c != 3:
        raise _lib.DvidError("frames_hwc: expected [n, H, W, 3]")
    hp = (oh + pad_to - 1) // pad_to * pad_to if pad_to > 0 else oh
    wp = (ow + pad_to - 1) // pad_to * pad_to if pad_to > 0 else ow
    need = _lib.lib().dvid_resize_workspace_bytes(n, hin, win, oh, ow)
    if need < 0:
        raise _lib.DvidError("dvid_resize_workspace_bytes: bad shape")
    ws = torch.empty((need,), device=frames_hwc.device, dtype=torch.uint8)
    out = torch.empty((n, 3, hp, wp), device=frames_hwc.device, dtype=torch.uint8)
    check(_lib.lib().dvid_resize_bilinear_u8(ptr(frames_hwc), n, hin, win, oh, ow, ptr(out), hp, wp, ptr(ws), need,
                                             cur_stream()), "dvid_resize_bilinear_u8")
    _cnt(4)
    return out


def maxpool3x3s2(x):
    _chk(x, H, "x")
    n, h, w, c = x.shape
    out = torch.empty((n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c), device=x.device, dtype=H)
    check(_lib.lib().dvid_maxpool3x3s2_nhwc_f16(ptr(x), ptr(out), n, h, w, c, cur_stream()), "dvid_maxpool")
    _cnt()
    return out


# ------------------------------------------------------------------------------------------------ decoder
# default: the mma.sync flash kernel.  The tcgen05 kernel (attention_tc.cu) is within 23 % (self, 300 keys: 12.1 vs 9.8 us)
# / 9 % (cross, 900 keys: 27.6 vs 25.3 us) of it in isolation and costs 1.2 % of the step inside the captured decode unit
# (profiles/r02_attention_ab.json): with head dim 32 both contractions are tiny (K = 32, N = 32) and the chain
# S -> TMEM -> softmax -> smem -> P.V -> TMEM of every key chunk is longer than keeping S and P in registers.
ATTENTION_TC = bool(int(__import__("os").environ.get("DVID_ATTN_TC", "0")))


def attention(q, k, v, out, batch, heads, lq, lk, q_rs, k_rs, v_rs, o_rs, q_bs, k_bs, v_bs, o_bs, tc=None):
    """q/k/v/out: fp16 tensors (possibly views into a packed buffer); strides in elements.  tc: tcgen05 kernel
    (dvid_attention_hd32_tc) or the mma.sync flash kernel (dvid_attention_hd32); None = module default ATTENTION_TC."""
    tc = ATTENTION_TC if tc is None else tc
    for t, nme in ((q, "q"), (k, "k"), (v, "v"), (out, "out")):
        if not t.is_cuda or t.dtype != H:
            raise _lib.DvidError(f"attention: {nme} must be CUDA fp16")
    with _prof("attention", 4.0 * batch * heads * lq * lk * 32, 2.0 * batch * heads * 32 * (2 * lq + 2 * lk)):
        fn = _lib.lib().dvid_attention_hd32_tc if tc else _lib.lib().dvid_attention_hd32
        check(fn(ptr(q), ptr(k), ptr(v), ptr(out), batch, heads, lq, lk, q_rs, k_rs, v_rs,
                 o_rs, q_bs, k_bs, v_bs, o_bs, cur_stream()), "dvid_attention_hd32_tc" if tc else "dvid_attention_hd32")
    _cnt()
    return out


class Levels:
    """The three FPN maps (NHWC fp16, [frames, h, w, 256]) as the host-side pointer arrays the C ABI takes."""

    def __init__(self, feats, scales=(1 / 8., 1 / 16., 1 / 32.)):
        for f in feats:
            _chk(f, H, "feat")
        self.feats = feats
        self.ptrs = (ctypes.c_void_p * 3)(*[f.data_ptr() for f in feats])
        self.hs = (ctypes.c_int * 3)(*[f.shape[1] for f in feats])
        self.ws = (ctypes.c_int * 3)(*[f.shape[2] for f in feats])
        self.scales = (ctypes.c_float * 3)(*scales)


def roi_align(levels, boxes, boxes_per_frame, want_roi=True, want_mean=True):
    _chk(boxes, F32, "boxes")
    m = boxes.numel() // 4
    dev = boxes.device
    roi = torch.empty((m, 49, 256), device=dev, dtype=H) if want_roi else None
    mean32 = torch.empty((m, 256), device=dev, dtype=F32) if want_mean else None
    mean16 = torch.empty((m, 256), device=dev, dtype=H) if want_mean else None
    with _prof("roi_align", 2.0 * m * 49 * 16 * 256, 2.0 * m * 49 * 256 * (16 + 1)):
        check(_lib.lib().dvid_roi_align(levels.ptrs, levels.hs, levels.ws, levels.scales, ptr(boxes), m,
                                        boxes_per_frame, ptr(roi), ptr(mean32), ptr(mean16), cur_stream()),
              "dvid_roi_align")
    _cnt()
    return roi, mean32, mean16


def dynconv_permutation(d=256, dd=64):
    """Row permutation of the dynamic_layer weight / bias that makes the GEMM emit the per-box weights in the layout
    of dvid_roi_dynconv_tc: new row j*d+i <- old row i*dd+j (P1^T), new row d*dd+i*dd+j <- old row d*dd+j*d+i (P2^T)."""
    i = torch.arange(d)[None, :]
    j = torch.arange(dd)[:, None]
    p1 = (i * dd + j).reshape(-1)                       # [j][i]
    p2 = d * dd + (torch.arange(dd)[None, :] * d + torch.arange(d)[:, None]).reshape(-1)     # [i][j]
    return torch.cat([p1, p2])


def roi_dynconv(levels, boxes, boxes_per_frame, params, g1, b1, g2, b2, roi_in=None, out=None, transposed=False):
    """transposed=False: params in the reference's layout, mma.sync kernel (dvid_roi_dynconv); transposed=True: params
    from a dynamic_layer whose rows were permuted with dynconv_permutation(), tcgen05 kernel (dvid_roi_dynconv_tc)."""
    _chk(boxes, F32, "boxes"); _chk(params, H, "params"); _chk(roi_in, H, "roi_in")
    for t in (g1, b1, g2, b2):
        _chk(t, F32, "ln")
    m = params.shape[0]
    if out is None:
        out = torch.empty((m, 49 * 256), device=params.device, dtype=H)
    lv = levels
    # algorithmic: two 49x256x64 bmm per box; bytes: generated weights in, 49x256 activations out (+ROI tile in)
    with _prof("roi_dynconv", 4.0 * m * 49 * 256 * 64, 2.0 * m * (32768 + 49 * 256 * 2)):
        fn = _lib.lib().dvid_roi_dynconv_tc if transposed else _lib.lib().dvid_roi_dynconv
        check(fn(lv.ptrs if lv else None, lv.hs if lv else None, lv.ws if lv else None,
                 lv.scales if lv else None, ptr(boxes), m, boxes_per_frame, ptr(roi_in),
                 ptr(params), ptr(g1), ptr(b1), ptr(g2), ptr(b2), ptr(out), cur_stream()),
              "dvid_roi_dynconv_tc" if transposed else "dvid_roi_dynconv")
    _cnt()
    return out


def gemm_row(a, w, bias=None, resid=None, ln=None, act=0, out_f32=None, out_f16=None):
    """Fused 256->256 Linear + bias [+ residual] [+ LayerNorm]: out_f32 = y, out_f16 = act(y) (dvid_gemm256_row)."""
    _chk(a, H, "a"); _chk(w, H, "w"); _chk(bias, F32, "bias"); _chk(resid, F32, "resid")
    _chk(out_f32, F32, "out_f32"); _chk(out_f16, H, "out_f16")
    m, k = a.shape
    if k != 256 or tuple(w.shape) != (256, 256):
        raise _lib.DvidError("gemm_row: a [m][256], w [256][256]")
    g, b = ln if ln is not None else (None, None)
    with _prof("conv_gemm", 2.0 * m * 256 * 256, 2.0 * (m * 256 + 65536) + 6.0 * m * 256,
               tag="gemm+row %dx256->256" % m):
        check(_lib.lib().dvid_gemm256_row(ptr(a), ptr(w), ptr(bias), ptr(resid), ptr(g), ptr(b), int(act), ptr(out_f32),
                                          ptr(out_f16), m, cur_stream()), "dvid_gemm256_row")
    _cnt()


def row_post(M, partials=None, splits=1, in_f16=None, bias=None, ln1=None, relu1=False, resid=None, ln2=None, act2=0,
             act2_f16_only=False, out_f32=None, out_f16=None, mod_scale=None, mod_shift=None, rows_per_group=1,
             scale_stride=0, shift_stride=0, shift_per_row=False, out_mod_f16=None):
    _chk(partials, F32, "partials"); _chk(in_f16, H, "in_f16"); _chk(bias, F32, "bias"); _chk(resid, F32, "resid")
    _chk(out_f32, F32, "out_f32"); _chk(out_f16, H, "out_f16"); _chk(out_mod_f16, H, "out_mod_f16")
    ln1g, ln1b = ln1 if ln1 is not None else (None, None)
    ln2g, ln2b = ln2 if ln2 is not None else (None, None)
    check(_lib.lib().dvid_row_post(ptr(partials), splits, M * 256, ptr(in_f16), ptr(bias), ptr(ln1g), ptr(ln1b),
                                   int(relu1), ptr(resid), ptr(ln2g), ptr(ln2b), act2, int(act2_f16_only),
                                   ptr(out_f32), ptr(out_f16), ptr(mod_scale), ptr(mod_shift), rows_per_group,
                                   scale_stride, shift_stride, int(shift_per_row), ptr(out_mod_f16), M, cur_stream()),
          "dvid_row_post")
    _cnt()


def small_linear(a, w, bias, act_in=0, act_out=0):
    _chk(a, F32, "a"); _chk(w, H, "w"); _chk(bias, F32, "bias")
    m, k = a.shape
    n = w.shape[0]
    out = torch.empty((m, n), device=a.device, dtype=F32)
    check(_lib.lib().dvid_small_linear(ptr(a), ptr(w), ptr(bias), ptr(out), m, n, k, act_in, act_out, cur_stream()),
          "dvid_small_linear")
    _cnt()
    return out


def time_sinusoid(t, freq):
    _chk(t, F32, "t"); _chk(freq, F32, "freq")
    out = torch.empty((t.numel(), 256), device=t.device, dtype=F32)
    check(_lib.lib().dvid_time_sinusoid(ptr(t), ptr(freq), ptr(out), t.numel(), cur_stream()), "dvid_time_sinusoid")
    _cnt()
    return out


def head_final(logit_part, cls_bias, C, delta_part, delta_bias, boxes_in, logits_out=None, boxes_out=None):
    _chk(logit_part, F32, "logit_part"); _chk(delta_part, F32, "delta_part"); _chk(boxes_in, F32, "boxes_in")
    M = logit_part.shape[0]
    dev = logit_part.device
    if logits_out is None:
        logits_out = torch.empty((M, C), device=dev, dtype=F32)
    if boxes_out is None:
        boxes_out = torch.empty((M, 4), device=dev, dtype=F32)
    check(_lib.lib().dvid_head_final(ptr(logit_part), logit_part.shape[1], ptr(cls_bias), C, ptr(delta_part),
                                     delta_part.shape[1], ptr(delta_bias), ptr(boxes_in), ptr(logits_out),
                                     ptr(boxes_out), M, cur_stream()), "dvid_head_final")
    _cnt()
    return logits_out, boxes_out


def head_tail(fc16, cls, reg, logit_w, logit_b, C, delta_w, delta_b, boxes_in):
    """Fused cls / reg towers + predictors + apply_deltas (dvid_head_tail).  cls = (w, (g, b)); reg = 3 x (w, (g, b));
    logit_w [32][256], delta_w [16][256] fp16 zero-padded.  Returns (logits [M,C], boxes [M,4]) fp32."""
    _chk(fc16, H, "fc16"); _chk(boxes_in, F32, "boxes_in"); _chk(logit_w, H, "logit_w"); _chk(delta_w, H, "delta_w")
    M = fc16.shape[0]
    dev = fc16.device
    logits = torch.empty((M, C), device=dev, dtype=F32)
    boxes = torch.empty((M, 4), device=dev, dtype=F32)
    (cw, (cg, cb)) = cls
    with _prof("head_tail", 2.0 * M * 256 * (4 * 256 + 34), 2.0 * M * 256 + 4.0 * M * (C + 8)):
        check(_lib.lib().dvid_head_tail(ptr(fc16), ptr(cw), ptr(cg), ptr(cb), ptr(logit_w), ptr(logit_b), C,
                                        ptr(reg[0][0]), ptr(reg[1][0]), ptr(reg[2][0]),
                                        ptr(reg[0][1][0]), ptr(reg[0][1][1]), ptr(reg[1][1][0]), ptr(reg[1][1][1]),
                                        ptr(reg[2][1][0]), ptr(reg[2][1][1]), ptr(delta_w), ptr(delta_b),
                                        ptr(boxes_in), ptr(logits), ptr(boxes), M, cur_stream()), "dvid_head_tail")
    _cnt()
    return logits, boxes


# ------------------------------------------------------------------------------------------------ diffusion loop
def noise_to_boxes(x, scale, W, Hh):
    _chk(x, F32, "x")
    out = torch.empty_like(x)
    check(_lib.lib().dvid_noise_to_boxes(ptr(x), ptr(out), x.numel() // 4, scale, W, Hh, cur_stream()),
          "dvid_noise_to_boxes")
    _cnt()
    return out


def ddim_step(logits, coord, x_t, eps, fill, scale, W, Hh, sqrt_recip_a, sqrt_recipm1_a, sqrt_a_next, c_coef, sigma):
    for t in (logits, coord, x_t, eps, fill):
        _chk(t, F32, "ddim input")
    frames, N, C = logits.shape
    x_next = torch.empty((frames, N, 4), device=logits.device, dtype=F32)
    boxes_next = torch.empty((frames, N, 4), device=logits.device, dtype=F32)
    kept = torch.empty((frames,), device=logits.device, dtype=torch.int32)
    check(_lib.lib().dvid_ddim_step(ptr(logits), C, ptr(coord), ptr(x_t), ptr(eps), ptr(fill), ptr(x_next),
                                    ptr(boxes_next), ptr(kept), frames, N, scale, W, Hh, sqrt_recip_a, sqrt_recipm1_a,
                                    sqrt_a_next, c_coef, sigma, cur_stream()), "dvid_ddim_step")
    _cnt()
    return x_next, boxes_next, kept


def topk_scores(logits, boxes, k, out_boxes, out_scores, out_labels, slot0):
    _chk(logits, F32, "logits"); _chk(boxes, F32, "boxes")
    frames, N, C = logits.shape
    cap = out_scores.shape[1]
    check(_lib.lib().dvid_topk_scores(ptr(logits), ptr(boxes), frames, N, C, k, ptr(out_boxes), ptr(out_scores),
                                      ptr(out_labels), cap, slot0, cur_stream()), "dvid_topk_scores")
    _cnt()


def topk_mask(logits, k1, k2):
    _chk(logits, F32, "logits")
    frames, N, C = logits.shape
    m1 = torch.empty((frames, N), device=logits.device, dtype=torch.uint8)
    m2 = torch.empty((frames, N), device=logits.device, dtype=torch.uint8)
    check(_lib.lib().dvid_topk_mask(ptr(logits), frames, N, C, k1, k2, ptr(m1), ptr(m2), cur_stream()),
          "dvid_topk_mask")
    _cnt()
    return m1, m2


def gather_masked_rows(src, mask, k):
    _chk(src, F32, "src")
    frames, N = mask.shape
    dst = torch.empty((frames * k, 256), device=src.device, dtype=F32)
    check(_lib.lib().dvid_gather_masked_rows(ptr(src), ptr(mask), frames, N, k, ptr(dst), cur_stream()),
          "dvid_gather_masked_rows")
    _cnt()
    return dst


NMS_WS_PER_FRAME = 1024 * 8 + 1024 * 16 + 1024 * 16 * 8   # DVID_NMS_WORKSPACE_PER_FRAME


def nms(boxes, scores, labels=None, counts=None, n=None, thr=0.5, plus_one=False, ge=False, ascending_out=False,
        clip_wh=None, want_compact=True):
    """boxes [frames][cap][4], scores [frames][cap], labels int32 [frames][cap] or None.
    Returns dict(keep int64 [frames][cap], count int32 [frames], boxes/scores/labels compacted)."""
    _chk(boxes, F32, "boxes"); _chk(scores, F32, "scores"); _chk(labels, torch.int32, "labels")
    _chk(counts, torch.int32, "counts")
    frames, cap = scores.shape
    if n is None:
        n = cap
    dev = boxes.device
    keep = torch.full((frames, cap), -1, device=dev, dtype=torch.int64)
    count = torch.empty((frames,), device=dev, dtype=torch.int32)
    ob = torch.empty((frames, cap, 4), device=dev, dtype=F32) if want_compact else None
    os_ = torch.empty((frames, cap), device=dev, dtype=F32) if want_compact else None
    ol = torch.empty((frames, cap), device=dev, dtype=torch.int32) if (want_compact and labels is not None) else None
    cw, ch = clip_wh if clip_wh is not None else (0.0, 0.0)
    ws_bytes = frames * NMS_WS_PER_FRAME
    ws = torch.empty((ws_bytes,), device=dev, dtype=torch.uint8)
    check(_lib.lib().dvid_nms(ptr(boxes), ptr(scores), ptr(labels), ptr(counts), n, cap, frames, thr, int(plus_one),
                              int(ge), int(ascending_out), cw, ch, ptr(keep), ptr(ob), ptr(os_), ptr(ol), ptr(count),
                              ptr(ws), ws_bytes, cur_stream()), "dvid_nms")
    _cnt(3)
    return dict(keep=keep, count=count, boxes=ob, scores=os_, labels=ol)


# ------------------------------------------------------------------------------------------------ Swin backbone
def swin_rows(B, Hh, W, C, x=None, write_x=False, add=None, add_mode=0, ln=None, out_f16=False, out_f32=False,
              out_mode=1, shift=0):
    """Row kernel over the fp32 residual stream (see dvid_swin_rows in include/dvid_b200.h).  Returns (out16, out32)."""
    _chk(x, F32, "x"); _chk(add, H, "add")
    g, b = ln if ln is not None else (None, None)
    _chk(g, F32, "gamma"); _chk(b, F32, "beta")
    dev = x.device if x is not None else add.device
    o16 = o32 = None
    if out_f16:
        rows = B * ((Hh + 6) // 7) * ((W + 6) // 7) * 49 if out_mode == 2 else B * Hh * W
        o16 = torch.empty((rows, C), device=dev, dtype=H)
    if out_f32:
        o32 = torch.empty((B, Hh, W, C), device=dev, dtype=F32)
    check(_lib.lib().dvid_swin_rows(ptr(x), int(write_x), ptr(add), add_mode, ptr(g), ptr(b), ptr(o16), ptr(o32),
                                    out_mode, B, Hh, W, C, shift, cur_stream()), "dvid_swin_rows")
    _cnt()
    return o16, o32


def swin_patch_merge(x, ln):
    _chk(x, F32, "x")
    B, Hh, W, C = x.shape
    out = torch.empty((B * ((Hh + 1) // 2) * ((W + 1) // 2), 4 * C), device=x.device, dtype=H)
    check(_lib.lib().dvid_swin_patch_merge(ptr(x), B, Hh, W, C, ptr(ln[0]), ptr(ln[1]), ptr(out), cur_stream()),
          "dvid_swin_patch_merge")
    _cnt()
    return out


def swin_patch_gather(img, mean, std):
    u8 = img.dtype == torch.uint8
    _chk(img, torch.uint8 if u8 else F32, "img")
    B, _, Hh, W = img.shape
    out = torch.empty((B * (Hh // 4) * (W // 4), 64), device=img.device, dtype=H)
    m = (ctypes.c_float * 3)(*mean)
    s = (ctypes.c_float * 3)(*std)
    fn = _lib.lib().dvid_swin_patch_gather_u8 if u8 else _lib.lib().dvid_swin_patch_gather
    check(fn(ptr(img), ptr(out), B, Hh, W, m, s, cur_stream()), "dvid_swin_patch_gather")
    _cnt()
    return out


# default: the mma.sync kernel - one 49x49x32 window per CTA finishes in registers; the tcgen05 variant (two windows per
# 128-row tile, TMEM round trips, per-CTA TMEM allocation) measured 2.5-3.3x slower at the Swin-B shapes
# (profiles/r02_attention_ab.json), so it is the opt-in
SWIN_ATTENTION_TC = bool(int(__import__("os").environ.get("DVID_SWIN_ATTN_TC", "0")))


def swin_window_attention(qkv, bias, B, Hh, W, C, heads, shift, tc=None):
    """tc: tcgen05 kernel (dvid_swin_window_attention_tc) or the mma.sync one; None = module default."""
    tc = SWIN_ATTENTION_TC if tc is None else tc
    _chk(qkv, H, "qkv"); _chk(bias, F32, "bias")
    out = torch.empty((qkv.shape[0], C), device=qkv.device, dtype=H)
    nw = B * ((Hh + 6) // 7) * ((W + 6) // 7)
    with _prof("attention", 4.0 * nw * heads * 49 * 49 * 32, 2.0 * qkv.numel() + 2.0 * out.numel()):
        fn = _lib.lib().dvid_swin_window_attention_tc if tc else _lib.lib().dvid_swin_window_attention
        check(fn(ptr(qkv), ptr(bias), ptr(out), B, Hh, W, C, heads, shift, cur_stream()),
              "dvid_swin_window_attention")
    _cnt()
    return out


# ------------------------------------------------------------------------------------------------ global memory
def cdist(x):
    _chk(x, F32, "x")
    n, d = x.shape
    out = torch.empty((n, n), device=x.device, dtype=F32)
    check(_lib.lib().dvid_cdist_f32(ptr(x), ptr(out), n, d, cur_stream()), "dvid_cdist_f32")
    _cnt()
    return out


def furthest_point_sampling(b, n, m, dist, temp, idx):
    """Same signature as mega_core._C.furthest_point_sampling (mega_core/csrc/fps.h:15-36); returns 1."""
    _chk(dist, F32, "points"); _chk(temp, F32, "temp"); _chk(idx, torch.int32, "idx")
    check(_lib.lib().dvid_furthest_point_sampling(b, n, m, ptr(dist), ptr(temp), ptr(idx), cur_stream()),
          "dvid_furthest_point_sampling")
    _cnt()
    return 1


def roi_align_legacy_forward(inp, rois, spatial_scale, pooled_height, pooled_width, sampling_ratio):
    """mega_core._C.roi_align_forward (csrc/ROIAlign.h:11-27): input (N,C,H,W) fp32, rois (n,5) fp32 ->
    (n,C,ph,pw) fp32, maskrcnn-benchmark semantics (no half-pixel shift, roi size >= 1)."""
    _chk(inp, F32, "input"); _chk(rois, F32, "rois")
    if inp.dim() != 4 or rois.dim() != 2 or rois.shape[1] != 5:
        raise _lib.DvidError("roi_align_forward: input (N,C,H,W), rois (n,5)")
    n = rois.shape[0]
    N, C, Hh, Ww = inp.shape
    out = torch.empty((n, C, int(pooled_height), int(pooled_width)), device=inp.device, dtype=F32)
    check(_lib.lib().dvid_roi_align_legacy_forward(ptr(inp), ptr(rois), n, C, Hh, Ww, float(spatial_scale),
                                                   int(pooled_height), int(pooled_width), int(sampling_ratio),
                                                   ptr(out), cur_stream()), "dvid_roi_align_legacy_forward")
    _cnt()
    return out


def vid_match(pred_boxes, pred_labels, order, pred_off, gt_boxes, gt_labels, gt_ignore, gt_off, iou_thresh,
              empty_weight):
    """vid_eval.py:167-291 greedy matching of all images in one launch (dvid_vid_match).  Packed device tensors: boxes
    fp32 [n,4], labels int32, offsets int32 [images+1], order int32 (per image, descending score), gt_ignore uint8.
    Returns (hit uint8 [Np], weight fp64 [Np]) indexed by position in `order`."""
    _chk(pred_boxes, F32, "pred_boxes"); _chk(gt_boxes, F32, "gt_boxes")
    for t, n in ((pred_labels, "pred_labels"), (order, "order"), (pred_off, "pred_off"), (gt_labels, "gt_labels"),
                 (gt_off, "gt_off")):
        _chk(t, torch.int32, n)
    _chk(gt_ignore, torch.uint8, "gt_ignore")
    n_img = pred_off.numel() - 1
    if gt_off.numel() != n_img + 1 or order.numel() != pred_labels.numel() or pred_boxes.shape[0] != order.numel() \
            or gt_boxes.shape[0] != gt_labels.numel() or gt_ignore.numel() != gt_labels.numel():
        raise _lib.DvidError("vid_match: inconsistent packed sizes")
    dev = pred_boxes.device
    taken = torch.zeros(max(gt_labels.numel(), 1), device=dev, dtype=torch.uint8)
    hit = torch.empty(max(order.numel(), 1), device=dev, dtype=torch.uint8)
    weight = torch.empty(max(order.numel(), 1), device=dev, dtype=torch.float64)
    check(_lib.lib().dvid_vid_match(ptr(pred_boxes), ptr(pred_labels), ptr(order), ptr(pred_off), ptr(gt_boxes),
                                    ptr(gt_labels), ptr(gt_ignore), ptr(gt_off), n_img, float(iou_thresh),
                                    float(empty_weight), ptr(taken), ptr(hit), ptr(weight), cur_stream()),
          "dvid_vid_match")
    _cnt()
    return hit[:order.numel()], weight[:order.numel()]


def decode_jpeg(data, device="cuda"):
    """One JPEG file (bytes) -> uint8 [H, W, 3] RGB on the device (nvJPEG behind dvid_jpeg_decode_rgb): what
    `Image.open(f).convert("RGB")` returns in the reference's datasets, decoded on the GPU."""
    buf = (ctypes.c_ubyte * len(data)).from_buffer_copy(data)
    w, h = ctypes.c_int(0), ctypes.c_int(0)
    check(_lib.lib().dvid_jpeg_info(buf, len(data), ctypes.byref(w), ctypes.byref(h)), "dvid_jpeg_info")
    out = torch.empty((h.value, w.value, 3), device=device, dtype=torch.uint8)
    check(_lib.lib().dvid_jpeg_decode_rgb(buf, len(data), ptr(out), w.value, h.value, cur_stream()),
          "dvid_jpeg_decode_rgb")
    _cnt()
    return out

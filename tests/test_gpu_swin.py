"""GPU parity of the Swin-B backbone path (BASELINE config 3, SURVEY.md 8a rows a3', a17): C-ABI kernels vs torch
restatements, the assembled body vs the REFERENCE-RUN golden features (tests/golden/ref_swin_small.pt) and vs the
fp16-emulating oracle, Swin + FPN, and a whole clip with the Swin config (INFER_BATCH 4, ALL_FRAME_INTERVAL 4).

Tolerances: fp16 storage of every GEMM output / LayerNorm output (the residual stream is fp32): body features within
2e-2 absolute of the fp32 reference values (|x| ~ 1 after the output LayerNorm; 99.9% of elements within 1e-2)."""
import os

import pytest
import torch
import torch.nn.functional as F

from diffusionvid_b200 import model as pm, ops, structures, synth
from oracle import model as om, swin as osw
from tests.parity_util import match_fraction

pytestmark = pytest.mark.gpu

SW = dict(embed=128, depths=(2, 2, 2, 2), heads=(4, 8, 16, 32))
HP = dict(num_proposals=100, num_classes=30, hidden=256, nheads=8, dim_dynamic=64, dim_ff=2048, num_heads=3,
          num_heads_local=1, num_cls=1, num_reg=3, sample_step=1, snr_scale=2.0, use_nms=True, infer_batch=4,
          all_frame_interval=4, key_frame_location=0, global_enable=True, mem_size=200, mem_size2=50,
          topk=(75, 25), pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375),
          blocks=(1, 1, 1, 1), swin=SW, device="cuda")


def test_swin_rows_modes(cuda):
    g = torch.Generator().manual_seed(0)
    B, Hh, W, C, sh = 2, 10, 16, 256, 3
    X = torch.randn(B, Hh, W, C, generator=g)
    add = (0.5 * torch.randn(B * Hh * W, C, generator=g)).half()
    gam, bet = 1 + 0.1 * torch.randn(C, generator=g), 0.1 * torch.randn(C, generator=g)
    # token-order add + LN -> windowed (pad + shift + partition) fp16
    xd = X.to(cuda).clone()
    o16, _ = ops.swin_rows(B, Hh, W, C, x=xd, write_x=True, add=add.to(cuda), add_mode=1, ln=(gam.to(cuda), bet.to(cuda)),
                           out_f16=True, out_mode=2, shift=sh)
    v = X + add.float().view(B, Hh, W, C)
    assert torch.allclose(xd.cpu(), v, atol=1e-6)
    ln = F.layer_norm(v, (C,), gam, bet, 1e-5)
    pad = F.pad(ln, (0, 0, 0, (7 - W % 7) % 7, 0, (7 - Hh % 7) % 7))
    ref = osw.window_partition(torch.roll(pad, (-sh, -sh), (1, 2)), 7).reshape(-1, C)
    assert o16.shape == ref.shape
    assert (o16.float().cpu() - ref).abs().max().item() <= 4e-3
    # windowed add (reverse + un-shift + crop) + LN -> token order
    Hp, Wp = pad.shape[1], pad.shape[2]
    addw = (0.5 * torch.randn(ref.shape[0], C, generator=g)).half()
    xd = X.to(cuda).clone()
    o16, _ = ops.swin_rows(B, Hh, W, C, x=xd, write_x=True, add=addw.to(cuda), add_mode=2, ln=(gam.to(cuda), bet.to(cuda)),
                           out_f16=True, shift=sh)
    rev = torch.roll(osw.window_reverse(addw.float().view(-1, 7, 7, C), 7, Hp, Wp), (sh, sh), (1, 2))[:, :Hh, :W]
    v = X + rev
    assert torch.allclose(xd.cpu(), v, atol=1e-6)
    assert (o16.float().cpu() - F.layer_norm(v, (C,), gam, bet, 1e-5).view(-1, C)).abs().max().item() <= 4e-3
    # patch merging with odd sizes
    Xo = torch.randn(2, 5, 7, 128, generator=g)
    g4, b4 = 1 + 0.1 * torch.randn(512, generator=g), 0.1 * torch.randn(512, generator=g)
    m = ops.swin_patch_merge(Xo.to(cuda), (g4.to(cuda), b4.to(cuda)))
    xp = F.pad(Xo, (0, 0, 0, 1, 0, 1))
    cat = torch.cat([xp[:, 0::2, 0::2], xp[:, 1::2, 0::2], xp[:, 0::2, 1::2], xp[:, 1::2, 1::2]], -1).reshape(-1, 512)
    assert (m.float().cpu() - F.layer_norm(cat, (512,), g4, b4, 1e-5)).abs().max().item() <= 4e-3


@pytest.mark.parametrize("tc", [False, True])
@pytest.mark.parametrize("shift", [0, 3])
@pytest.mark.parametrize("B", [2, 1, 3])        # 12, 6 and 18 windows; with B = 3 x (1 x 3) windows below: odd count
def test_window_attention_matches_reference_math(cuda, shift, tc, B):
    g = torch.Generator().manual_seed(1)
    Hh, W, C, nh = (10, 16, 128, 4) if B != 3 else (7, 16, 128, 4)
    nwy, nwx = (Hh + 6) // 7, (W + 6) // 7
    rows = B * nwy * nwx * 49
    qkv = (0.7 * torch.randn(rows, 3 * C, generator=g)).half()
    bias = 0.5 * torch.randn(nh, 49, 49, generator=g)
    out = ops.swin_window_attention(qkv.to(cuda), bias.to(cuda), B, Hh, W, C, nh, shift, tc=tc)
    q, k, v = qkv.float().view(-1, 49, 3, nh, 32).permute(2, 0, 3, 1, 4)
    att = (q * 32 ** -0.5) @ k.transpose(-2, -1) + bias[None]
    if shift:
        mask = osw.shift_mask(Hh, W, 7, 3)
        att = (att.view(B, nwy * nwx, nh, 49, 49) + mask[None, :, None]).view(-1, nh, 49, 49)
    ref = (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(rows, C)
    assert (out.float().cpu() - ref).abs().max().item() <= 3e-3


def test_gemm_gelu_epilogue(cuda):
    g = torch.Generator().manual_seed(2)
    a = torch.randn(300, 128, generator=g).half()
    w = (torch.randn(512, 128, generator=g) / 11).half()
    b = 0.1 * torch.randn(512, generator=g)
    out = ops.gemm(a.to(cuda), w.to(cuda), b.to(cuda), relu=2)
    ref = F.gelu(F.linear(a.float(), w.float(), b))
    assert (out.float().cpu() - ref).abs().max().item() <= 4e-3


def _model(seed=91, **over):
    hp = dict(HP, **over)
    sd = synth.make_state_dict(seed=seed, swin=hp["swin"])
    m = pm.DiffusionDet(hp)
    m.load_state_dict(sd, strict=False)
    m.to("cuda")
    return hp, sd, m


def test_swin_body_matches_reference_golden(cuda):
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ref_swin_small.pt"), weights_only=False)
    meta = gold["meta"]
    assert meta["cfg"]["embed"] == 128
    hp, sd, m = _model(seed=meta["weight_seed"], swin=dict(meta["cfg"]))
    m._pack()
    x = torch.randn(*meta["shape"], generator=torch.Generator().manual_seed(meta["input_seed"]))
    # the body consumes [0,1] images and normalises itself: feed un-normalised input that normalises to x
    mean = torch.tensor(hp["pixel_mean"]).view(1, 3, 1, 1) / 255.
    std = torch.tensor(hp["pixel_std"]).view(1, 3, 1, 1) / 255.
    got = m._swin_body((x * std + mean).to(cuda))
    o16 = osw.swin_body(om.Ctx(sd, om.Quant(True)), x)
    for f, key in zip(got, ("swin1", "swin2", "swin3")):
        f = f.float().cpu().permute(0, 3, 1, 2)
        ref = gold["out"][key]
        assert f.shape == ref.shape
        d = (f - ref).abs()
        assert d.max().item() <= 3e-2 and (d <= 1e-2).float().mean().item() >= 0.999, (key, d.max().item())
        assert (f - o16[key]).abs().max().item() <= 2e-2, key


def test_swin_fpn_features_match_oracle(cuda):
    hp, sd, m = _model()
    imgs = synth.make_clip(2, 160, 224, seed=3)
    o = om.OracleDiffusionVID(sd, dict(num_proposals=hp["num_proposals"]), fp16=True)
    ref = o.backbone(imgs)
    m._pack()
    got = m.extract_features(imgs.to(cuda))
    for gt, r in zip(got, ref):
        gt = gt.float().cpu().permute(0, 3, 1, 2)
        assert gt.shape == r.shape
        scale = r.abs().max().item()
        assert (gt - r).abs().max().item() <= 2e-2 * scale
        assert (gt - r).abs().mean().item() <= 2e-3 * scale


def test_swin_clip_end_to_end(cuda):
    """vid_Swin_B_DiffusionVID.yaml protocol at small size: INFER_BATCH 4, ALL_FRAME_INTERVAL 4, MAX_OFFSET 3, T=1."""
    h, w, L = 128, 192, 10
    hp, sd, m = _model()
    noise = om.NoiseSource(9, hp["num_proposals"])
    m.noise = noise
    ocfg = {k: hp[k] for k in ("num_proposals", "sample_step", "mem_size", "mem_size2", "topk", "infer_batch",
                               "all_frame_interval")}
    o = om.OracleDiffusionVID(sd, ocfg, fp16=True, noise=noise)
    frames = synth.make_clip(L, h, w, seed=6)
    samples = synth.clip_samples(frames, [7, 2, 5], h, w, infer_batch=4, max_offset=3)
    fracs = []
    for s in samples:
        ref = o.forward(s)
        got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                     ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                     ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                     frame_id=s["frame_id"], start_id=0, end_id=L - 1, seg_len=L,
                     frame_category=s["frame_category"], video_id=0))
        assert len(got) == len(ref)
        for g_, r in zip(got, ref):
            fracs.append(match_fraction(g_.bbox.cpu(), g_.get_field("scores").cpu(), g_.get_field("labels").cpu(),
                                        r["boxes"], r["scores"], r["labels"], max(h, w), box_tol=2e-3, score_tol=4e-3))
    assert len(fracs) == L
    fr = sorted(fracs)
    assert fr[len(fr) // 2] >= 0.95 and sum(fracs) / len(fracs) >= 0.9, fracs

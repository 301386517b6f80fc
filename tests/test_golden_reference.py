"""The oracle against outputs of the REFERENCE'S OWN CODE (tests/golden/ref_diffusionvid_small.pt).

The fixture was produced by tests/golden/make_golden.py, which executes the unmodified reference sources
mega_core/modeling/detector/diffusion_det.py (_forward_test, model_predictions, inference, update_erase_memory),
mega_core/modeling/roi_heads/box_head/box_head.py (DynamicHead, RCNNHead, RCNNHead_cond, DynamicConv, time MLP),
box_head/loss.py (box format helpers) and mega_core/structures/* on the CPU, with torchvision's real roi_align /
batched_nms behind a restated detectron2 ROIPooler, on seeded synthetic weights / clip / noise.  This pins rows a1, a4-a14
and a16 of SURVEY.md 8(a); the detectron2 backbone (a3) and the CUDA FPS kernel (a15) stay restatement-only.

Tolerances: fp32 vs fp32 with different summation orders (nn.MultiheadAttention vs einsum, bmm vs matmul):
1e-4 absolute on logits / features, 1e-3 px on boxes; clip detections are matched as sets (parity_util).
"""
import os

import pytest
import torch

from diffusionvid_b200 import synth
from oracle import model as om
from tests.parity_util import match_fraction

PATH = os.path.join(os.path.dirname(__file__), "golden", "ref_diffusionvid_small.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(PATH, weights_only=False)


@pytest.fixture(scope="module")
def setup(gold):
    m = gold["meta"]
    sd = synth.make_state_dict(seed=m["weight_seed"], blocks=tuple(m["blocks"]))
    frames = synth.make_clip(m["L"], m["h"], m["w"], seed=m["clip_seed"])
    return m, sd, frames


def test_schedule_matches_reference(gold):
    ac = om.cosine_alphas_cumprod()
    assert torch.equal(ac, gold["schedule"]["alphas_cumprod"])
    assert torch.equal(torch.sqrt(1. / ac), gold["schedule"]["sqrt_recip"])
    assert torch.equal(torch.sqrt(1. / ac - 1), gold["schedule"]["sqrt_recipm1"])


def test_time_embedding_matches_reference(gold, setup):
    m, sd, _ = setup
    c = om.Ctx(sd, om.Quant(False))
    e = om.time_embedding(c, torch.tensor([999, 749, 0]))
    assert (e - gold["head"]["time_emb"]).abs().max().item() <= 2e-5


def test_backbone_standin_matches_oracle(gold, setup):
    """the golden script's plain-conv backbone and the oracle's folded-BN backbone agree (both restate detectron2)."""
    m, sd, frames = setup
    o = om.OracleDiffusionVID(sd, dict(num_proposals=m["N"]), fp16=False)
    f = o.backbone(frames[:2])
    for got, key in zip(f, ("p3", "p4", "p5")):
        ref = gold["head"][key]
        assert got.shape == ref.shape
        assert (got - ref).abs().max().item() <= 2e-4 * max(1.0, ref.abs().max().item())


def test_dynamic_head_extraction_matches_reference(gold, setup):
    """DynamicHead.forward(box_extract=1): 3 x RCNNHead (ROIPooler, self-attention, DynamicConv, FFN, time modulation,
    towers, apply_deltas) + per-frame top-75 / top-25 selection, on the reference's own feature maps."""
    m, sd, _ = setup
    N, h, w = m["N"], m["h"], m["w"]
    g = gold["head"]
    o = om.OracleDiffusionVID(sd, dict(num_proposals=N), fp16=False, noise=om.NoiseSource(m["noise_seed"], N))
    feats = [g["p3"], g["p4"], g["p5"]]
    whwh = torch.tensor([w, h, w, h], dtype=torch.float32)[None].expand(2, -1)
    x = o.noise.get("init", 0, 0, 0, 2)
    boxes = o._x_to_boxes(x, whwh)
    temb = om.time_embedding(o.c, torch.full((2,), 999, dtype=torch.long))
    lg, bx, obj = om.head_base_stages(o.c, feats, boxes, temb, o.cfg)
    assert (lg - g["logits"]).abs().max().item() <= 1e-4
    assert (bx - g["boxes"]).abs().max().item() <= 1e-3
    assert (obj - g["obj"].reshape(-1, 256)).abs().max().item() <= 1e-4
    k1, k2 = om.select_topk_feats(lg, obj, 2, N, [min(75, N), min(25, N)])
    assert k1.shape == g["k1"].shape and k2.shape == g["k2"].shape
    assert (k1 - g["k1"]).abs().max().item() <= 1e-4
    assert (k2 - g["k2"]).abs().max().item() <= 1e-4


@pytest.mark.parametrize("T", [1, 4])
def test_clip_matches_reference_forward_test(gold, setup, T):
    """whole clips through DiffusionDet._forward_test of the reference vs the oracle: global memory (FPS), cached-stage
    (T=1) and re-run (T=4) branches, DDIM update with renewal, ensemble without the last step, top-k, batched NMS,
    clip_to_image."""
    m, sd, frames = setup
    N, h, w, L = m["N"], m["h"], m["w"], m["L"]
    o = om.OracleDiffusionVID(sd, dict(num_proposals=N, sample_step=T, mem_size=m["mem_size"]), fp16=False,
                              noise=om.NoiseSource(m["noise_seed"], N))
    samples = synth.clip_samples(frames, m["global_idx"], h, w)
    got = []
    for s in samples:
        got += o.forward(s)
    ref = gold["clip_T%d" % T]
    assert len(got) == len(ref) == L
    mem = gold["mem_T%d" % T]
    assert o.mem[0].shape == mem[0].shape and (o.mem[0] - mem[0]).abs().max().item() <= 1e-4
    assert o.mem[1].shape == mem[1].shape and (o.mem[1] - mem[1]).abs().max().item() <= 1e-4
    fracs = []
    for g_, r in zip(got, ref):
        assert tuple(r["size"]) == (w, h)
        assert r["labels"].min().item() >= 1 and r["labels"].max().item() <= 30
        fracs.append(match_fraction(g_["boxes"], g_["scores"], g_["labels"], r["boxes"], r["scores"], r["labels"],
                                    max(h, w), box_tol=1e-4, score_tol=1e-5))
        fracs.append(match_fraction(r["boxes"], r["scores"], r["labels"], g_["boxes"], g_["scores"], g_["labels"],
                                    max(h, w), box_tol=1e-4, score_tol=1e-5))
    assert min(fracs) >= 0.98 and sum(fracs) / len(fracs) >= 0.995, fracs
    same = sum(int(g_["scores"].numel() == r["scores"].numel()) for g_, r in zip(got, ref))
    assert same >= L - 1


def test_swin_body_matches_reference_swintransformer():
    """oracle.swin.swin_body vs the reference's own swintransformer.py (tests/golden/make_golden_swin.py): patch embed,
    W-MSA / SW-MSA with zero-padded windows (24x40 tokens -> 28x42), relative position bias, shift mask, patch merging
    with odd sizes (3x5 at the last stage), per-output LayerNorm."""
    from oracle import swin
    gold = torch.load(os.path.join(os.path.dirname(__file__), "golden", "ref_swin_small.pt"), weights_only=False)
    m = gold["meta"]
    sd = synth.make_state_dict(seed=m["weight_seed"], swin=m["cfg"])
    x = torch.randn(*m["shape"], generator=torch.Generator().manual_seed(m["input_seed"]))
    out = swin.swin_body(om.Ctx(sd, om.Quant(False)), x)
    assert set(out) == set(gold["out"])
    for k, ref in gold["out"].items():
        assert out[k].shape == ref.shape
        assert (out[k] - ref).abs().max().item() <= 2e-4, k

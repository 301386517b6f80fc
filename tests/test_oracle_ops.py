"""CPU tests that pin the oracle's operator restatements (oracle/ops.py, oracle/model.py) to the libraries the reference
actually calls (torchvision roi_align / batched_nms, torch.nn.MultiheadAttention) and to the constants quoted from the
reference in SURVEY.md Appendix B.  No GPU needed."""
import math

import numpy as np
import pytest
import torch
import torchvision

from oracle import model as om
from oracle import ops as oo


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


@pytest.mark.parametrize("scale,hw", [(1 / 8., (76, 128)), (1 / 32., (19, 32))])
def test_roi_align_matches_torchvision(scale, hw):
    g = gen(1)
    feat = torch.randn(2, 16, *hw, generator=g)
    n = 64
    c = torch.rand(n, 4, generator=g)
    boxes = om.box_cxcywh_to_xyxy(c) * torch.tensor([1000., 600., 1000., 600.])
    boxes[0] = torch.tensor([10., 10., 10., 10.])              # zero area
    boxes[1] = torch.tensor([-300., -200., 1500., 900.])       # far outside
    boxes[2] = torch.tensor([990., 590., 1000., 600.])
    boxes[3] = torch.tensor([1200., 700., 1300., 800.])        # fully outside -> zeros
    rois = torch.cat([torch.randint(0, 2, (n, 1), generator=g).float(), boxes], dim=1)
    ref = torchvision.ops.roi_align(feat, rois, 7, scale, 2, True)
    got = oo.roi_align(feat, rois, 7, scale, 2)
    assert (got - ref).abs().max().item() <= 2e-5


def test_roi_pooler_level_assignment():
    # sqrt(area)=224 -> level 4 (index 1); 112 -> 3; 448 -> 5; tiny/zero -> 3; huge -> 5   
    b = torch.tensor([[0., 0., 224., 224.], [0., 0., 112., 112.], [0., 0., 448., 448.], [5., 5., 5., 5.],
                      [0., 0., 2000., 2000.], [0., 0., 223.9, 223.9], [0., 0., 447.9, 447.9]])
    assert oo.assign_levels(b).tolist() == [1, 0, 2, 0, 2, 0, 1]


def test_mha_matches_torch():
    g = gen(2)
    E, nh = 256, 8
    m = torch.nn.MultiheadAttention(E, nh, dropout=0.0)
    with torch.no_grad():
        for p in m.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    q = torch.randn(30, 3, E, generator=g)
    kv = torch.randn(45, 3, E, generator=g)
    ref = m(q, kv, kv)[0]
    got, _ = oo.mha(q, kv, kv, m.in_proj_weight, m.in_proj_bias, m.out_proj.weight, m.out_proj.bias, nh)
    assert (got - ref).abs().max().item() <= 1e-5
    # the model-level attention (fp32 mode) is the same function
    sd = {"a.in_proj_weight": m.in_proj_weight.detach(), "a.in_proj_bias": m.in_proj_bias.detach(),
          "a.out_proj.weight": m.out_proj.weight.detach(), "a.out_proj.bias": m.out_proj.bias.detach()}
    got2 = om.attention(om.Ctx(sd, om.Quant(False)), q, kv, "a", nh)
    assert (got2 - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("n,ncls", [(300, 30), (900, 30), (50, 1)])
def test_batched_nms_matches_torchvision(n, ncls):
    g = gen(n)
    ctr = torch.rand(n // 3 + 1, 4, generator=g)[torch.randint(0, n // 3 + 1, (n,), generator=g)]
    c = (ctr + 0.02 * torch.randn(n, 4, generator=g)).clamp(0.02, 0.98)
    boxes = om.box_cxcywh_to_xyxy(torch.stack([c[:, 0], c[:, 1], 0.1 + 0.3 * c[:, 2], 0.1 + 0.3 * c[:, 3]], 1))
    boxes = boxes * torch.tensor([1000., 600., 1000., 600.])
    scores = torch.rand(n, generator=g)
    labels = torch.randint(1, ncls + 1, (n,), generator=g)
    ref = torchvision.ops.batched_nms(boxes, scores, labels, 0.5)
    got = oo.batched_nms(boxes, scores, labels, 0.5)
    assert torch.equal(got, ref)
    assert 0 < got.numel() < n


def test_fps_no_ties_equals_plain_greedy():
    g = gen(3)
    x = torch.randn(300, 16, generator=g)
    d = oo.cdist_l2(x).numpy()
    got = oo.fps(d, 40)
    temp = np.full(300, 1e10, dtype=np.float32)
    picks = [0]
    for _ in range(39):
        temp = np.minimum(temp, d[picks[-1]])
        picks.append(int(np.argmax(temp)))
    assert got.tolist() == picks


def test_fps_tie_rule_closed_form():
    """The literal emulation of the reference's scan + tree (oracle.ops.fps) picks, among equal maxima, the candidate
    with the smallest (bit_reverse(k mod bs), k): two tied slots first meet in the tree at their lowest differing bit
    and the lower slot wins (mega_core/csrc/cuda/fps.cu:60-136).  The CUDA kernel implements the closed form."""
    n = 1500                                            # reference block size 1024
    d = np.ones((n, n), dtype=np.float32)
    np.fill_diagonal(d, 0.0)
    got = oo.fps(d, 12).tolist()

    def brev10(s):
        return int(format(s, "010b")[::-1], 2)
    picked = [0]
    for _ in range(11):
        cand = [k for k in range(n) if k not in picked]
        picked.append(min(cand, key=lambda k: (brev10(k % 1024), k)))
    assert got == picked
    assert got[:6] == [0, 1024, 512, 256, 1280, 768]
    assert oo.fps_block_size(1800) == 1024 and oo.fps_block_size(600) == 512 and oo.fps_block_size(37) == 32


def test_schedule_constants_from_reference():
    """SURVEY.md Appendix B quotes the T=4 DDIM constants derived from diffusion_det.py:50-61,578-584."""
    ac = om.cosine_alphas_cumprod()
    assert ac.dtype == torch.float32 and ac.shape == (1000,)
    assert abs(ac[999].item() - 2.4288e-9) / 2.4288e-9 < 1e-3
    expect = {(999, 749): (0.144272, 0.925056, 1.1103e-4, 0.379832),
              (749, 499): (0.493844, 0.647065, 0.295742, 0.702740),
              (499, 249): (0.847012, 0.355003, 0.164197, 0.920333)}
    for (t, tn), (a_next, sigma, c, sq) in expect.items():
        a = ac[t].double(); an = ac[tn].double()
        s2 = (1 - a / an) * (1 - an) / (1 - a)
        assert abs(an.item() - a_next) < 2e-6
        assert abs(s2.sqrt().item() - sigma) < 2e-6
        assert abs((1 - an - s2).sqrt().item() - c) / c < 2e-3
        assert abs(an.sqrt().item() - sq) < 2e-6
    assert abs(torch.sqrt(1. / ac[999]).item() - 20291.17) / 20291.17 < 1e-3


def test_apply_deltas_identity_and_clamp():
    boxes = torch.tensor([[10., 20., 110., 220.]])
    out = om.apply_deltas(torch.zeros(1, 4), boxes)
    assert torch.allclose(out, boxes)
    big = om.apply_deltas(torch.tensor([[0., 0., 50., 50.]]), boxes)
    assert abs((big[0, 2] - big[0, 0]).item() - 100.0 * 100000.0 / 16) / (100.0 * 100000.0 / 16) < 1e-5


def test_cdist_direct_vs_the_reference_call_and_fps_picks():
    """The reference calls torch.cdist(feats, feats, p=2.0) with the DEFAULT compute mode (diffusion_det.py:880), which
    for n > 25 takes the matmul route (|a|^2 + |b|^2 - 2ab, clamp, sqrt); the oracle and the sm_100a kernel take direct
    differences (oracle/ops.py::cdist_l2).  The two differ by fp32 cancellation error only.  Bound it on memory-like
    rows (LayerNorm outputs, 256 wide, as update_erase_memory sees them: 600 -> 150 and 1800 -> 900), and check that
    farthest-point sampling picks the same rows from either matrix up to near-tie swaps."""
    for seed, (n, m) in enumerate([(600, 150), (1800, 900)]):
        x = torch.nn.functional.layer_norm(torch.randn(n, 256, generator=gen(40 + seed)) * 1.5, (256,))
        direct = oo.cdist_l2(x)
        ref = torch.cdist(x, x, p=2.0)                       # 'use_mm_for_euclid_dist_if_necessary' -> mm path here
        off = ~torch.eye(n, dtype=torch.bool)
        rel = ((direct - ref).abs()[off] / ref[off]).max().item()
        assert rel <= 2e-5, rel                              # fp32 cancellation of the mm route at |x|^2 = 256
        assert direct.diagonal().abs().max().item() == 0.0   # exact zeros on the diagonal (the mm route leaves ~1e-3)
        a = set(oo.fps(direct.numpy(), m).tolist())
        b = set(oo.fps(ref.numpy(), m).tolist())
        assert len(a & b) >= 0.97 * m, (len(a & b), m)


@pytest.mark.parametrize("sampling_ratio", [2, 0])
def test_legacy_roi_align_restatement_matches_torchvision(sampling_ratio):
    """oracle/legacy.py::roi_align_legacy (csrc/cuda/ROIAlign_cuda.cu:65-125) against the installed torchvision
    roi_align(aligned=False): same algorithm (both derive from caffe2's), incl. malformed / out-of-image rois."""
    from oracle import legacy
    g = gen(77)
    feat = torch.randn(2, 5, 20, 30, generator=g)
    rois = torch.tensor([[0, 10., 12., 100., 90.], [1, -20., -8., 40., 33.], [0, 50., 50., 50., 50.],
                         [1, 200., 100., 260., 170.], [0, 0., 0., 239., 159.], [1, 30.5, 20.25, 28., 19.]])
    got = legacy.roi_align_legacy(feat.numpy(), rois.numpy(), 0.125, 7, 7, sampling_ratio)
    ref = torchvision.ops.roi_align(feat, rois, (7, 7), 0.125, sampling_ratio, False)
    assert np.abs(got - ref.numpy()).max() <= 2e-6

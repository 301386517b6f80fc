"""Known-answer tests of the legacy native NMS (mega_core._C.nms) from the reference's own tests/test_nms.py
(vectors in tests/golden/nms_vectors.json, extracted by executing that file - tests/golden/extract_nms_vectors.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle.legacy import nms_legacy

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "nms_vectors.json")))["cases"]


def test_vectors_present():
    assert len(CASES) == 6 and max(len(c["scores"]) for c in CASES) == 53


@pytest.mark.parametrize("i", range(len(CASES)))
def test_oracle_legacy_nms_matches_reference_vectors(i):
    c = CASES[i]
    keep = nms_legacy(np.array(c["boxes"], np.float32), np.array(c["scores"], np.float32), c["thresh"], ge=True)
    assert np.sort(keep).tolist() == c["expected"]


@pytest.mark.gpu
@pytest.mark.parametrize("i", range(len(CASES)))
def test_c_shim_nms_matches_reference_vectors(cuda, i):
    from diffusionvid_b200 import _C_shim
    c = CASES[i]
    keep = _C_shim.nms(torch.tensor(c["boxes"], device=cuda), torch.tensor(c["scores"], device=cuda), c["thresh"])
    assert keep.dtype == torch.int64
    assert keep.cpu().tolist() == c["expected"]            # ascending original indices, like cuda/nms.cu:127-130


@pytest.mark.gpu
def test_c_shim_nms_random_boxes_match_oracle(cuda):
    from diffusionvid_b200 import _C_shim
    g = torch.Generator().manual_seed(5)
    for n in (1, 2, 37, 300, 1000):
        xy = torch.rand(n, 2, generator=g) * 200
        wh = torch.rand(n, 2, generator=g) * 60 + 1
        boxes = torch.cat([xy, xy + wh], 1)
        scores = torch.rand(n, generator=g)
        ref = nms_legacy(boxes.numpy(), scores.numpy(), 0.4, ge=False)
        got = _C_shim.nms(boxes.to(cuda), scores.to(cuda), 0.4)
        assert got.cpu().tolist() == ref.tolist()
    assert _C_shim.nms(torch.zeros(0, 4, device=cuda), torch.zeros(0, device=cuda), 0.5).numel() == 0
    with pytest.raises(NotImplementedError):
        _C_shim.roi_pool_forward()


@pytest.mark.gpu
@pytest.mark.parametrize("sampling_ratio", [2, 0])
def test_c_shim_legacy_roi_align_forward(cuda, sampling_ratio):
    """mega_core._C.roi_align_forward (csrc/cuda/ROIAlign_cuda.cu:65-125) over libdvid_b200.so vs the numpy restatement
    (oracle/legacy.py, itself pinned to torchvision aligned=False) - known-answer shapes: regular, partly and wholly
    outside the map, zero-size (clamped to 1x1), inverted, plus a bulk random set against torchvision on the CPU."""
    import torchvision
    from diffusionvid_b200 import _C_shim
    from oracle import legacy
    g = torch.Generator().manual_seed(77)
    feat = torch.randn(2, 37, 20, 30, generator=g)
    rois = torch.tensor([[0, 10., 12., 100., 90.], [1, -20., -8., 40., 33.], [0, 50., 50., 50., 50.],
                         [1, 200., 100., 260., 170.], [0, 0., 0., 239., 159.], [1, 30.5, 20.25, 28., 19.]])
    got = _C_shim.roi_align_forward(feat.to(cuda), rois.to(cuda), 0.125, 7, 7, sampling_ratio)
    assert got.shape == (6, 37, 7, 7) and got.dtype == torch.float32
    ref = legacy.roi_align_legacy(feat.numpy(), rois.numpy(), 0.125, 7, 7, sampling_ratio)
    assert np.abs(got.cpu().numpy() - ref).max() <= 1e-6       # same fp32 operation order; division by count differs by <= 1 ulp
    # bulk: 500 random rois on a 256-channel map, rectangular pooling, against torchvision's CPU kernel
    feat = torch.randn(3, 256, 38, 64, generator=g)
    xy = torch.rand(500, 2, generator=g) * torch.tensor([900., 500.])
    wh = torch.rand(500, 2, generator=g) * 300
    rois = torch.cat([torch.randint(0, 3, (500, 1), generator=g).float(), xy, xy + wh], 1)
    got = _C_shim.roi_align_forward(feat.to(cuda), rois.to(cuda), 1 / 16., 7, 14, sampling_ratio)
    ref = torchvision.ops.roi_align(feat, rois, (7, 14), 1 / 16., sampling_ratio, False)
    assert (got.cpu() - ref).abs().max().item() <= 1e-5
    assert _C_shim.roi_align_forward(feat.to(cuda), torch.zeros(0, 5, device=cuda), 0.5, 7, 7, 2).shape == (0, 256, 7, 7)


@pytest.mark.gpu
def test_c_shim_fps_signature(cuda):
    """same call as diffusion_det.py:892-896: points (1,n,n), temp (1,n) = 1e10, idx (1,k) int32 -> returns 1."""
    from diffusionvid_b200 import _C_shim
    from oracle import ops as oo
    x = torch.randn(70, 16, generator=torch.Generator().manual_seed(1))
    d = oo.cdist_l2(x)
    temp = torch.full((1, 70), 1e10, device=cuda)
    idx = torch.zeros((1, 20), dtype=torch.int32, device=cuda)
    assert _C_shim.furthest_point_sampling(1, 70, 20, d.to(cuda)[None].contiguous(), temp, idx) == 1
    assert idx[0].cpu().tolist() == oo.fps(d.numpy(), 20).tolist()
    assert _C_shim.furthest_point_sampling(1, 70, 20, d[None], temp, idx) == -1      # CPU tensor: -1 like fps.h:35

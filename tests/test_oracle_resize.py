"""Pins oracle/resize.py (restatement of Pillow's 8-bit antialiased bilinear resample, which the reference reaches
through transforms.py:31-67 -> torchvision F.resize on a PIL image) against the Pillow installed here, bit for bit,
and the size rule against the reference's arithmetic on the VID frame sizes."""
import numpy as np
import pytest

from oracle import resize as orz

PIL = pytest.importorskip("PIL")
from PIL import Image  # noqa: E402


@pytest.mark.parametrize("h,w,oh,ow", [(72, 128, 56, 100), (90, 160, 75, 133), (60, 100, 60, 100), (50, 80, 75, 120),
                                       (97, 131, 40, 131), (64, 64, 23, 77), (33, 47, 99, 20), (720, 1280, 562, 1000)])
def test_resize_matches_pillow_bit_exact(h, w, oh, ow):
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    img[: h // 4] = 255; img[h // 4: h // 2, : w // 2] = 0            # saturated regions exercise the clamp
    want = np.asarray(Image.fromarray(img).resize((ow, oh), Image.BILINEAR))
    got = orz.resize_bilinear_u8(img, oh, ow)
    assert got.shape == want.shape and np.array_equal(got, want)


def test_torchvision_resize_on_pil_is_the_same_call():
    tv = pytest.importorskip("torchvision")
    import torchvision.transforms.functional as F
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, size=(90, 160, 3), dtype=np.uint8)
    want = np.asarray(F.resize(Image.fromarray(img), (56, 100)))
    assert np.array_equal(orz.resize_bilinear_u8(img, 56, 100), want)


def test_size_rule():
    """transforms.py:38-59 with MIN_SIZE_TEST 600 / MAX_SIZE_TEST 1000 (configs/BASE_RCNN_*gpu.yaml)."""
    assert orz.get_size((1280, 720)) == (562, 999)        # capped: size = round(1000 * 720 / 1280) = 562, ow = int(562 * 1280 / 720)
    assert orz.get_size((640, 480)) == (600, 800)
    assert orz.get_size((480, 640)) == (800, 600)
    assert orz.get_size((1000, 600)) == (600, 1000)
    assert orz.get_size((500, 500)) == (600, 600)


def test_clip_loader_size_rule_and_argument_checks():
    """Host side of diffusionvid_b200/clip_loader.py: the size rule equals Resize.get_size (restated in the oracle) on
    a grid of frame sizes; non-uint8 / non-HWC input is rejected before anything touches the device."""
    import torch
    from diffusionvid_b200 import clip_loader
    for w in (320, 500, 640, 1000, 1280, 1920):
        for h in (180, 240, 480, 500, 600, 720, 1080):
            for mn, mx in ((600, 1000), (150, 250), (800, 1333)):
                assert clip_loader.get_size((w, h), mn, mx) == orz.get_size((w, h), mn, mx)
    t = clip_loader.GpuFrameTransform(600, 1000, 32, device="cpu")
    with pytest.raises(ValueError):
        t(torch.zeros(2, 3, 8, 8, dtype=torch.uint8))            # CHW, not the decoder's HWC
    with pytest.raises(ValueError):
        t(torch.zeros(8, 8, 3, dtype=torch.float32))

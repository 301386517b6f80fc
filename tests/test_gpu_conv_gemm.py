"""GPU parity: tcgen05 implicit-GEMM conv / GEMM (dvid_conv2d_nhwc_f16, dvid_gemm_f16) vs a plain fp32 PyTorch
reference of the same op on fp16-rounded inputs. Tolerance: fp32 accumulation, fp16 output rounding -> 2e-3 relative
to the output scale (one fp16 ulp is 9.8e-4)."""
import ctypes

import pytest
import torch

from diffusionvid_b200 import _lib

pytestmark = pytest.mark.gpu


def _gemm(a, w, bias=None, resid=None, relu=0, splits=0):
    L = _lib.lib()
    m, k = a.shape
    n = w.shape[0]
    if splits:
        used = ctypes.c_int(0)
        part = torch.empty(splits, m, n, device=a.device, dtype=torch.float32)
        _lib.check(L.dvid_gemm_f16(_lib.ptr(a), _lib.ptr(w), None, None, None, _lib.ptr(part), m, n, k, 0, splits,
                                   ctypes.byref(used), _lib.cur_stream()), "gemm")
        return part[:used.value].sum(0)
    out = torch.empty(m, n, device=a.device, dtype=torch.float16)
    _lib.check(L.dvid_gemm_f16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(resid), _lib.ptr(out), None, m, n, k,
                               relu, 1, None, _lib.cur_stream()), "gemm")
    return out


@pytest.mark.parametrize("m,n,k", [(128, 64, 64), (128, 256, 256), (2400, 768, 256), (2400, 256, 2048),
                                   (300, 32768, 256), (1000, 136, 72), (77, 256, 12544)])
def test_gemm_plain(cuda, m, n, k):
    g = torch.Generator(device="cpu").manual_seed(m * 31 + n * 7 + k)
    a = (torch.randn(m, k, generator=g) * 0.5).half().to(cuda)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(cuda)
    out = _gemm(a, w)
    ref = a.float() @ w.float().t()
    torch.cuda.synchronize()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


def test_gemm_epilogue(cuda):
    g = torch.Generator(device="cpu").manual_seed(5)
    m, n, k = 2400, 512, 256
    a = torch.randn(m, k, generator=g).half().to(cuda)
    w = (torch.randn(n, k, generator=g) / 16).half().to(cuda)
    bias = torch.randn(n, generator=g).to(cuda)
    resid = torch.randn(m, n, generator=g).half().to(cuda)
    out = _gemm(a, w, bias, resid, relu=1)
    ref = torch.relu(a.float() @ w.float().t() + bias + resid.float())
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("splits", [2, 8, 49])
def test_gemm_splitk(cuda, splits):
    g = torch.Generator(device="cpu").manual_seed(9)
    m, n, k = 2400, 256, 12544
    a = (torch.randn(m, k, generator=g) * 0.1).half().to(cuda)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(cuda)
    out = _gemm(a, w, splits=splits)
    ref = a.float() @ w.float().t()
    err = (out - ref).abs().max().item()
    assert err <= 1e-4 * max(1.0, ref.abs().max().item()), err


def _conv(x_nhwc, w_k, bias, resid, cout, R, S, stride, pad, resid_shift, relu):
    L = _lib.lib()
    n, h, w, cin = x_nhwc.shape
    ho = (h + 2 * pad - R) // stride + 1
    wo = (w + 2 * pad - S) // stride + 1
    out = torch.full((n, ho, wo, cout), float("nan"), device=x_nhwc.device, dtype=torch.float16)
    _lib.check(L.dvid_conv2d_nhwc_f16(_lib.ptr(x_nhwc), _lib.ptr(w_k), _lib.ptr(bias), _lib.ptr(resid), _lib.ptr(out),
                                      n, h, w, cin, cout, R, S, stride, pad, resid_shift, relu, _lib.cur_stream()),
               "conv")
    return out


@pytest.mark.parametrize("n,h,w,cin,cout,R,stride,pad", [
    (2, 19, 32, 64, 64, 3, 1, 1),
    (2, 38, 64, 256, 256, 3, 1, 1),
    (1, 76, 128, 128, 128, 3, 1, 1),
    (3, 10, 10, 512, 512, 3, 1, 1),
    (2, 40, 40, 256, 1024, 1, 1, 0),
    (1, 20, 20, 1024, 256, 1, 1, 0),
    (2, 152, 256, 64, 64, 1, 1, 0),
])
def test_conv_stride1(cuda, n, h, w, cin, cout, R, stride, pad):
    g = torch.Generator(device="cpu").manual_seed(n + h * 3 + cin)
    x = torch.randn(n, cin, h, w, generator=g).half().to(cuda)
    wt = (torch.randn(cout, cin, R, R, generator=g) / (cin * R * R) ** 0.5).half().to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    resid = torch.randn(n, cout, h, w, generator=g).half().to(cuda)
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_k = wt.permute(0, 2, 3, 1).contiguous().view(cout, -1)
    resid_nhwc = resid.permute(0, 2, 3, 1).contiguous()
    out = _conv(x_nhwc, w_k, bias, resid_nhwc, cout, R, R, stride, pad, 0, 1)
    ref = torch.relu(torch.nn.functional.conv2d(x.float(), wt.float(), bias, stride=stride, padding=pad) + resid.float())
    ref = ref.permute(0, 2, 3, 1)
    assert not torch.isnan(out).any()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


def test_conv_fpn_topdown(cuda):
    """lateral 1x1 conv + nearest-x2 upsampled coarser map (FPN sum fuse), no activation."""
    g = torch.Generator(device="cpu").manual_seed(3)
    n, h, w, cin, cout = 2, 38, 64, 1024, 256
    x = torch.randn(n, cin, h, w, generator=g).half().to(cuda)
    wt = (torch.randn(cout, cin, 1, 1, generator=g) / cin ** 0.5).half().to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    top = torch.randn(n, cout, h // 2, w // 2, generator=g).half().to(cuda)
    out = _conv(x.permute(0, 2, 3, 1).contiguous(), wt.view(cout, cin).contiguous(), bias,
                top.permute(0, 2, 3, 1).contiguous(), cout, 1, 1, 1, 0, 1, 0)
    ref = torch.nn.functional.conv2d(x.float(), wt.float(), bias) + \
        torch.nn.functional.interpolate(top.float(), scale_factor=2, mode="nearest")
    err = (out.float() - ref.permute(0, 2, 3, 1)).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("n,h,w,cin,cout,R,pad", [(2, 40, 40, 128, 128, 3, 1), (1, 152, 256, 128, 128, 3, 1),
                                                   (2, 38, 64, 512, 1024, 1, 0)])
def test_conv_stride2_tma(cuda, n, h, w, cin, cout, R, pad):
    """stride-2 convolution through TMA element strides (no im2col)."""
    g = torch.Generator(device="cpu").manual_seed(17)
    x = torch.randn(n, cin, h, w, generator=g).half().to(cuda)
    wt = (torch.randn(cout, cin, R, R, generator=g) / (cin * R * R) ** 0.5).half().to(cuda)
    out = _conv(x.permute(0, 2, 3, 1).contiguous(), wt.permute(0, 2, 3, 1).contiguous().view(cout, -1), None, None,
                cout, R, R, 2, pad, 0, 0)
    ref = torch.nn.functional.conv2d(x.float(), wt.float(), None, stride=2, padding=pad).permute(0, 2, 3, 1)
    assert not torch.isnan(out).any()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("m,n,k", [(2400, 32768, 256), (2400, 4096, 128), (5000, 1024, 192), (40000, 256, 64)])
def test_gemm_weight_stationary_walk(cuda, m, n, k):
    """shapes that take the B-stationary tile walk (K <= 256, >= 2 tiles per SM): bias + ReLU epilogue, ragged M."""
    g = torch.Generator(device="cpu").manual_seed(m + n + k)
    a = (torch.randn(m, k, generator=g) * 0.5).half().to(cuda)
    w = (torch.randn(n, k, generator=g) / k ** 0.5).half().to(cuda)
    bias = torch.randn(n, generator=g).to(cuda)
    out = _gemm(a, w, bias, None, relu=1)
    # reference in chunks of columns to bound memory
    err, scale = 0.0, 1.0
    for c0 in range(0, n, 4096):
        ref = torch.relu(a.float() @ w[c0:c0 + 4096].float().t() + bias[c0:c0 + 4096])
        err = max(err, (out[:, c0:c0 + 4096].float() - ref).abs().max().item())
        scale = max(scale, ref.abs().max().item())
    assert err <= 2e-3 * scale, err


def test_conv1x1_weight_stationary_with_residual(cuda):
    """res4-style conv3: 256 -> 1024 on 8 x 38 x 64 pixels + residual + ReLU (152 x 4 tiles -> B-stationary walk)."""
    g = torch.Generator(device="cpu").manual_seed(23)
    n, h, w, cin, cout = 8, 38, 64, 256, 1024
    x = torch.randn(n, h, w, cin, generator=g).half().to(cuda)
    wt = (torch.randn(cout, cin, generator=g) / cin ** 0.5).half().to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    resid = torch.randn(n, h, w, cout, generator=g).half().to(cuda)
    out = _conv(x, wt, bias, resid, cout, 1, 1, 1, 0, 0, 1)
    ref = torch.relu(x.float().view(-1, cin) @ wt.float().t() + bias + resid.float().view(-1, cout)).view(n, h, w, cout)
    assert not torch.isnan(out).any()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err


@pytest.mark.parametrize("cin,cout,R", [(256, 256, 3), (1024, 256, 1), (512, 256, 3)])
def test_conv_partial_last_wave(cuda, cin, cout, R):
    """res4-style layers: 8 x 38 x 64 pixels = 152 tiles on 148 SMs (a full wave plus four tiles): the persistent tile
    walk must cover the partial second wave; repeated launches give identical results."""
    g = torch.Generator(device="cpu").manual_seed(cin + R)
    n, h, w = 8, 38, 64
    x = torch.randn(n, h, w, cin, generator=g).half().to(cuda)
    wt = (torch.randn(cout, R, R, cin, generator=g) / (cin * R * R) ** 0.5).half().to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias,
                                                padding=R // 2)).permute(0, 2, 3, 1)
    first = _conv(x, wt.view(cout, -1), bias, None, cout, R, R, 1, R // 2, 0, 1)
    err = (first.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err
    for _ in range(2):
        assert torch.equal(_conv(x, wt.view(cout, -1), bias, None, cout, R, R, 1, R // 2, 0, 1), first)


@pytest.fixture
def streamk():
    L = _lib.lib()
    _lib.check(L.dvid_conv_streamk(1), "dvid_conv_streamk")
    yield
    _lib.check(L.dvid_conv_streamk(0), "dvid_conv_streamk")


@pytest.mark.parametrize("n,h,w,cin,cout,R,topdown", [
    (8, 38, 64, 256, 256, 3, False),     # res4 conv2: 152 tiles x 36 k-blocks, BN=256
    (8, 38, 64, 1024, 256, 1, False),    # res4 conv1: 152 tiles x 16 k-blocks
    (8, 38, 64, 1024, 256, 1, True),     # FPN lateral4 + nearest-x2 top-down residual (epilogue residual path)
    (8, 76, 128, 128, 128, 3, False),    # res3 conv2: 608 tiles x 18 k-blocks, BN=128 (4.1 waves)
    (5, 37, 61, 512, 192, 3, False),     # ragged image / channel tails, BN=64 or 128
])
def test_conv_stream_k_schedule(cuda, streamk, n, h, w, cin, cout, R, topdown):
    """Stream-K schedule (dvid_conv_streamk): the (tile, k-block) space is cut into one equal range per CTA, partial
    accumulator tiles are parked in the workspace and finished by the CTA holding the tile's first k-blocks.  Must
    match the fp32 reference like the whole-tile schedule, agree with it to fp32 re-association noise (<= 1 fp16 ulp
    of the output scale), be repeatable, and leave the flags re-armed for the next launch."""
    g = torch.Generator(device="cpu").manual_seed(cin + R + n)
    x = torch.randn(n, h, w, cin, generator=g).half().to(cuda)
    wt = (torch.randn(cout, R, R, cin, generator=g) / (cin * R * R) ** 0.5).half().to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    resid = torch.randn(n, (h + 1) // 2, (w + 1) // 2, cout, generator=g).half().to(cuda) if topdown else None
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias,
                                     padding=R // 2)
    if topdown:
        ref = ref + torch.nn.functional.interpolate(resid.float().permute(0, 3, 1, 2), scale_factor=2,
                                                    mode="nearest")[:, :, :h, :w]
    else:
        ref = torch.relu(ref)
    ref = ref.permute(0, 2, 3, 1)
    run = lambda: _conv(x, wt.view(cout, -1), bias, resid, cout, R, R, 1, R // 2, 1 if topdown else 0,
                        0 if topdown else 1)
    first = run()
    assert not torch.isnan(first).any()
    scale = max(1.0, ref.abs().max().item())
    assert (first.float() - ref).abs().max().item() <= 2e-3 * scale
    for _ in range(3):
        assert torch.equal(run(), first)
    _lib.check(_lib.lib().dvid_conv_streamk(0), "dvid_conv_streamk")
    plain = run()
    _lib.check(_lib.lib().dvid_conv_streamk(1), "dvid_conv_streamk")
    assert (first.float() - plain.float()).abs().max().item() <= 1e-3 * scale
    assert torch.equal(run(), first)


@pytest.mark.parametrize("n,h,w,cin,cout,R,stride", [
    (9, 38, 64, 256, 256, 3, 1),      # 171 M tiles: odd, the last pair has an out-of-range partner tile
    (8, 38, 64, 256, 384, 3, 1),      # second N tile half empty: the peer CTA's half of the weight rows is out of range
    (8, 76, 128, 256, 256, 3, 2),     # stride 2 (res4.0 conv2)
    (8, 19, 32, 2048, 512, 1, 1),     # 1x1, 32 k-blocks, 40 M tiles x 2 N tiles
    (8, 37, 61, 256, 256, 3, 1),      # ragged image: tile tails in x and y
])
def test_conv_cta_pair_variant(cuda, n, h, w, cin, cout, R, stride):
    """Shapes that take the cta_group::2 pair variant of the conv kernel (plain 256-wide tiles, >= 16 k-blocks; default
    on, DVID_CONV_CTA2=0 switches it off): pairs of adjacent M tiles, each CTA loading half of the weight rows."""
    g = torch.Generator(device="cpu").manual_seed(n * 100 + cout + R)
    x = torch.randn(n, h, w, cin, generator=g).half().to(cuda)
    wt = (torch.randn(cout, R, R, cin, generator=g) / (cin * R * R) ** 0.5).half().to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    ref = torch.relu(torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias,
                                                stride=stride, padding=R // 2)).permute(0, 2, 3, 1)
    out = _conv(x, wt.view(cout, -1), bias, None, cout, R, R, stride, R // 2, 0, 1)
    assert not torch.isnan(out).any()
    err = (out.float() - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), err
    assert torch.equal(_conv(x, wt.view(cout, -1), bias, None, cout, R, R, stride, R // 2, 0, 1), out)

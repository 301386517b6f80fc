"""Size-independent properties of the CUDA kernels at the headline sizes (BASELINE config 2: 8 frames of 608x1024,
N=300, T=4 -> 900 NMS candidates per frame, 1800 -> 900 memory rows): what must hold whatever the data, checked with
plain tensor algebra on the device instead of the CPU oracle (which takes minutes at these sizes).

  * convolution / GEMM: a centre-tap identity filter returns the input bit for bit; identity 1x1 + residual is one
    exact fp16 addition; linearity in the weights.
  * ROIAlign: partition of unity (a constant map pools to the constant for boxes inside the image), linearity.
  * attention: rows of the softmax sum to one (constant V -> constant output), key-permutation invariance.
  * top-k: output scores sorted, the k-th score equals the k-th largest sigmoid of the input.
  * NMS: output sorted by score, survivors mutually below the threshold per class, every suppressed candidate
    overlaps a better survivor of its class, idempotence (NMS of the survivors keeps all, in order).
  * farthest-point sampling: distinct picks starting at row 0, every pick is the arg-max of the distance to the set
    picked so far, the pick distances never increase.
  * DDIM / box maps: round trip boxes -> noise-space -> boxes.
"""
import pytest
import torch
import torch.nn.functional as F

from diffusionvid_b200 import ops

pytestmark = pytest.mark.gpu

W_IMG, H_IMG = 1000.0, 600.0


def gen(seed):
    return torch.Generator(device="cpu").manual_seed(seed)


# ------------------------------------------------------------------------------------------------ conv / GEMM
@pytest.mark.parametrize("shape", [(8, 38, 64, 256), (8, 76, 128, 128), (3, 19, 32, 512)])
def test_conv3x3_centre_tap_identity_returns_the_input(cuda, shape):
    n, h, w, c = shape
    x = torch.randn(shape, generator=gen(1)).half().to(cuda)
    wt = torch.zeros(c, 9, c, dtype=torch.float16)
    wt[:, 4, :] = torch.eye(c, dtype=torch.float16)               # tap (1,1) of [cout][R*S][cin]
    out = ops.conv2d(x, wt.view(c, 9 * c).contiguous().to(cuda), torch.zeros(c, device=cuda), c, 3, 3, 1, 1, False)
    assert torch.equal(out, x)
    # a shifted tap: out[y][x] = in[y][x+1], zero in the last column (padding)
    wt.zero_()
    wt[:, 5, :] = torch.eye(c, dtype=torch.float16)
    out = ops.conv2d(x, wt.view(c, 9 * c).contiguous().to(cuda), torch.zeros(c, device=cuda), c, 3, 3, 1, 1, False)
    assert torch.equal(out[:, :, :-1], x[:, :, 1:]) and not out[:, :, -1].any()


def test_identity_1x1_with_residual_and_relu_is_one_fp16_rounding(cuda):
    n, h, w, c = 8, 38, 64, 1024                                 # res4 conv3 shape: residual through the tensor core
    g = gen(2)
    x = torch.randn(n, h, w, c, generator=g).half().to(cuda)
    r = torch.randn(n, h, w, c, generator=g).half().to(cuda)
    eye = torch.eye(c, dtype=torch.float16).to(cuda)
    bias = (torch.randint(-8, 9, (c,), generator=g).float() / 4).to(cuda)
    out = ops.conv2d(x, eye, bias, c, 1, 1, 1, 0, True, resid=r)
    want = F.relu(x.float() + r.float() + bias).half()
    # x * 1.0 and the residual are exact in the fp32 accumulator; only the order of the two fp32 additions is free, so
    # a result may differ by one fp16 ulp where the fp32 sum falls on a rounding boundary
    diff = (out.float() - want.float()).abs()
    assert float((diff == 0).float().mean()) >= 0.9999 and float(diff.max()) <= 2.0 ** -7


def test_gemm_is_linear_in_the_weights(cuda):
    g = gen(3)
    a = torch.randn(2400, 256, generator=g).half().to(cuda)
    # weights on a coarse grid: products and sums are exact in fp32, so linearity holds up to the final fp16 rounding
    w1 = (torch.randint(-4, 5, (2048, 256), generator=g).float() / 8).half().to(cuda)
    w2 = (torch.randint(-4, 5, (2048, 256), generator=g).float() / 8).half().to(cuda)
    o1, _ = ops.gemm_partials(a, w1, 1)
    o2, _ = ops.gemm_partials(a, w2, 1)
    o12, _ = ops.gemm_partials(a, (w1 + w2).contiguous(), 1)
    assert (o12[0] - (o1[0] + o2[0])).abs().max().item() <= 1e-3      # fp32 accumulation order only
    # split-K is a re-association of the same sum
    o4, used = ops.gemm_partials(a, w1, 4)
    assert used == 4 and (o4[:used].sum(0) - o1[0]).abs().max().item() <= 1e-3


# ------------------------------------------------------------------------------------------------ ROIAlign
def _inside_boxes(g, frames, n):
    c = torch.rand(frames, n, 2, generator=g) * torch.tensor([W_IMG - 200, H_IMG - 200]) + 100
    wh = torch.rand(frames, n, 2, generator=g) * 180 + 8
    return torch.cat([c - wh / 2, c + wh / 2], -1).contiguous()


def test_roi_align_partition_of_unity_and_linearity(cuda):
    frames, n = 8, 300
    g = gen(4)
    hw = ((76, 128), (38, 64), (19, 32))
    boxes = _inside_boxes(g, frames, n).to(cuda)
    const = [torch.full((frames, h, w, 256), 0.75, dtype=torch.float16, device=cuda) for h, w in hw]
    roi, mean32, _ = ops.roi_align(ops.Levels(const), boxes, n)
    assert torch.equal(roi, torch.full_like(roi, 0.75))               # weights of every bin sum to exactly 1 here
    assert (mean32 - 0.75).abs().max().item() <= 1e-6
    f1 = [torch.randn(frames, h, w, 256, generator=g).half().to(cuda) for h, w in hw]
    f2 = [torch.randn(frames, h, w, 256, generator=g).half().to(cuda) for h, w in hw]
    r1 = ops.roi_align(ops.Levels(f1), boxes, n)[0].float()
    r2 = ops.roi_align(ops.Levels(f2), boxes, n)[0].float()
    f12 = [(a.float() + b.float()).half() for a, b in zip(f1, f2)]
    r12 = ops.roi_align(ops.Levels(f12), boxes, n)[0].float()
    # three fp16 roundings (two outputs, the summed map) of values up to ~4
    assert (r12 - (r1 + r2)).abs().max().item() <= 8e-3


# ------------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("tc", [False, True])
def test_attention_rows_sum_to_one_and_ignore_key_order(cuda, tc):
    batch, n, heads = 8, 300, 8
    g = gen(5)
    q = (torch.randn(batch * n, 256, generator=g)).half().to(cuda)
    k = (torch.randn(batch * n, 256, generator=g)).half().to(cuda)
    v = torch.randn(batch * n, 256, generator=g).half().to(cuda)
    out = torch.empty(batch * n, 256, dtype=torch.float16, device=cuda)

    def run(kk, vv):
        o = torch.empty_like(out)
        ops.attention(q, kk, vv, o, batch, heads, n, n, 256, 256, 256, 256, n * 256, n * 256, n * 256, n * 256, tc=tc)
        return o
    cv = torch.full_like(v, 0.5)
    assert (run(k, cv).float() - 0.5).abs().max().item() <= 1e-3      # sum of the probabilities is one
    base = run(k, v)
    perm = torch.randperm(n, generator=g).to(cuda)
    kp = k.view(batch, n, 256)[:, perm].reshape(batch * n, 256).contiguous()
    vp = v.view(batch, n, 256)[:, perm].reshape(batch * n, 256).contiguous()
    assert (run(kp, vp).float() - base.float()).abs().max().item() <= 4e-3   # summation order + fp16 output


# ------------------------------------------------------------------------------------------------ top-k / NMS
def test_topk_output_is_sorted_and_cuts_at_the_kth_score(cuda):
    frames, N, C, k = 8, 300, 30, 300
    g = gen(6)
    logits = (torch.randn(frames, N, C, generator=g) * 2 - 1).to(cuda)
    boxes = _inside_boxes(g, frames, N).to(cuda)
    ob = torch.zeros(frames, 900, 4, device=cuda)
    osc = torch.zeros(frames, 900, device=cuda)
    ol = torch.zeros(frames, 900, device=cuda, dtype=torch.int32)
    ops.topk_scores(logits, boxes, k, ob, osc, ol, 0)
    s = osc[:, :k]
    assert bool((s[:, :-1] >= s[:, 1:]).all())
    kth = torch.sigmoid(logits.view(frames, -1)).topk(k, dim=1)[0]
    assert (s - kth).abs().max().item() <= 1e-6
    lab = ol[:, :k].long()
    assert int(lab.min()) >= 1 and int(lab.max()) <= C                  # labels are 1-based class ids
    # every output box is one of the frame's input boxes
    d = (ob[:, :k, None, :] - boxes[:, None, :, :]).abs().sum(-1).min(-1)[0]
    assert float(d.max()) == 0.0


def _pair_iou(b):
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(b[:, None, :2], b[None, :, :2])
    rb = torch.min(b[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    inter = wh[..., 0] * wh[..., 1]
    return inter / (area[:, None] + area[None, :] - inter)


def test_nms_invariants_at_900_candidates(cuda):
    frames, n, ncls, thr = 8, 900, 30, 0.5
    g = gen(7)
    ctr = torch.rand(frames, 60, 4, generator=g)
    pick = torch.randint(0, 60, (frames, n), generator=g)
    c = (torch.gather(ctr, 1, pick[..., None].expand(-1, -1, 4)) + torch.randn(frames, n, 4, generator=g) * 0.02)
    c = c.clamp(0.02, 0.98)
    wh = 0.1 + 0.3 * c[..., 2:]
    boxes = (torch.cat([c[..., :2] - wh / 2, c[..., :2] + wh / 2], -1) *
             torch.tensor([W_IMG, H_IMG, W_IMG, H_IMG])).contiguous().to(cuda)
    scores = torch.rand(frames, n, generator=g).to(cuda)
    labels = torch.randint(1, ncls + 1, (frames, n), generator=g).int().to(cuda)
    r = ops.nms(boxes, scores, labels, thr=thr)
    for f in range(frames):
        cnt = int(r["count"][f])
        keep = r["keep"][f, :cnt]
        assert 0 < cnt < n and keep.unique().numel() == cnt
        ks = scores[f][keep]
        assert bool((ks[:-1] >= ks[1:]).all())
        iou = _pair_iou(boxes[f])
        same = labels[f][:, None] == labels[f][None, :]
        kk = iou[keep][:, keep]
        kk.fill_diagonal_(0)
        assert float((kk * same[keep][:, keep]).max()) <= thr + 1e-5                # survivors do not overlap
        dropped = torch.ones(n, dtype=torch.bool, device=cuda)
        dropped[keep] = False
        cover = (iou[:, keep] > thr - 1e-5) & same[:, keep] & (scores[f][keep][None, :] >= scores[f][:, None])
        assert bool(cover[dropped].any(1).all())                                    # every drop has a better cause
        # idempotence: the survivors alone survive again, in the same order
        again = ops.nms(boxes[f][keep][None].contiguous(), ks[None].contiguous(),
                        labels[f][keep][None].contiguous(), thr=thr)
        assert int(again["count"][0]) == cnt
        assert torch.equal(again["keep"][0, :cnt], torch.arange(cnt, device=cuda))


# ------------------------------------------------------------------------------------------------ memory sampling
def test_fps_greedy_property_at_1800_to_900(cuda):
    n, m = 1800, 900
    x = torch.randn(n, 256, generator=gen(8)).to(cuda)
    dist = ops.cdist(x)
    assert float(dist.diagonal().abs().max()) == 0.0 and float((dist - dist.t()).abs().max()) <= 1e-5
    temp = torch.full((1, n), 1e10, device=cuda)
    idx = torch.zeros((1, m), dtype=torch.int32, device=cuda)
    ops.furthest_point_sampling(1, n, m, dist, temp, idx)
    picks = idx[0].long()
    assert int(picks[0]) == 0 and picks.unique().numel() == m
    # replay: distance of every point to the picked set before each pick; the pick must be its arg-max
    rows = dist[picks]                                              # [m][n]
    run_min = torch.cummin(rows, 0)[0]                              # after pick j: min over picks 0..j
    before = run_min[:-1]                                           # state seen by pick j+1
    chosen = before.gather(1, picks[1:, None])[:, 0]
    assert torch.equal(chosen, before.max(1)[0])
    assert bool((chosen[:-1] >= chosen[1:]).all())                  # the covering radius never grows


# ------------------------------------------------------------------------------------------------ box maps
def test_noise_to_boxes_round_trip(cuda):
    frames, n, scale = 8, 300, 2.0
    x = torch.randn(frames, n, 4, generator=gen(9)).to(cuda)
    boxes = ops.noise_to_boxes(x, scale, W_IMG, H_IMG)
    assert boxes.shape == (frames, n, 4)
    # invert: xyxy/size -> cxcywh -> (v*2-1)*scale; clamped coordinates come back at +-scale
    b = boxes / torch.tensor([W_IMG, H_IMG, W_IMG, H_IMG], device=cuda)
    cxcywh = torch.cat([(b[..., :2] + b[..., 2:]) / 2, b[..., 2:] - b[..., :2]], -1)
    back = (cxcywh * 2 - 1) * scale
    assert (back - x.clamp(-scale, scale)).abs().max().item() <= 1e-4

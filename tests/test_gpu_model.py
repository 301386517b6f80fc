"""GPU parity of the assembled path (backbone, decoder heads, whole clips) against the fp16-emulating CPU oracle
(oracle/model.py, Quant(True)) and sanity bounds against the fp32 oracle, on the same seeded weights / clip / noise.

Tolerances (fp16 storage between kernels, fp32 accumulation):
  feature maps   : max |d| <= 1% of the map's max (a handful of fp16 ulps after ~20 convs), mean |d| <= 1e-3 of max
  head outputs   : >= 99% of the boxes within 2e-3 of the image size (boxes) / 2e-2 (logits); the remaining rows are
                   boxes whose FPN level / 0.5 threshold flipped on an ulp (see tests/parity_util.py)
  whole clips    : per frame, detections matched by label + box (2e-3 of image size) + score (4e-3); the median frame
                   must match >= 95%, and T=1 (no renewal feedback) must match >= 95% on average.
"""
import pytest
import torch

from diffusionvid_b200 import model as pm, ops, structures, synth
from oracle import model as om
from tests.parity_util import match_fraction, rows_within

pytestmark = pytest.mark.gpu

HP = dict(num_proposals=100, num_classes=30, hidden=256, nheads=8, dim_dynamic=64, dim_ff=2048, num_heads=3,
          num_heads_local=1, num_cls=1, num_reg=3, sample_step=4, snr_scale=2.0, use_nms=True, infer_batch=8,
          all_frame_interval=8, key_frame_location=0, global_enable=True, mem_size=300, mem_size2=50,
          topk=(75, 25), pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375),
          blocks=(2, 2, 3, 2), device="cuda")


def _models(T, seed=21, hp_over=None):
    hp = dict(HP, sample_step=T)
    if hp_over:
        hp.update(hp_over)
    sd = synth.make_state_dict(seed=seed, blocks=hp["blocks"])
    m = pm.DiffusionDet(hp)
    m.load_state_dict(sd, strict=False)
    m.to("cuda")
    noise = om.NoiseSource(9, hp["num_proposals"])
    m.noise = noise
    ocfg = {k: hp[k] for k in ("num_proposals", "sample_step", "mem_size", "mem_size2", "topk")}
    return hp, sd, m, noise, ocfg


def test_backbone_r101_matches_oracle(cuda):
    """full R-101 + FPN depth (104 convs) at a small image: product NHWC fp16 maps vs the fp16 oracle."""
    hp, sd, m, _, ocfg = _models(1, hp_over=dict(blocks=(3, 4, 23, 3)))
    imgs = synth.make_clip(2, 160, 224, seed=8)
    o = om.OracleDiffusionVID(sd, ocfg, fp16=True)
    ref = o.backbone(imgs)
    m._pack()
    got = m.extract_features(imgs.to(cuda))
    for g, r in zip(got, ref):
        g = g.float().cpu().permute(0, 3, 1, 2)
        assert g.shape == r.shape
        scale = r.abs().max().item()
        assert (g - r).abs().max().item() <= 1e-2 * scale
        assert (g - r).abs().mean().item() <= 1e-3 * scale


def test_head_stages_match_oracle(cuda):
    """head_series[0..2] on oracle features: logits / boxes / object features per stage chain."""
    hp, sd, m, noise, ocfg = _models(1)
    h, w = 192, 256
    imgs = synth.make_clip(3, h, w, seed=4)
    o = om.OracleDiffusionVID(sd, ocfg, fp16=True, noise=noise)
    feats = o.backbone(imgs)                                     # fp16-rounded values, NCHW fp32
    B, N = 3, hp["num_proposals"]
    whwh = torch.tensor([w, h, w, h], dtype=torch.float32)[None].expand(B, -1)
    x = noise.get("init", 0, 0, 0, B)
    boxes = o._x_to_boxes(x, whwh)
    temb = om.time_embedding(o.c, torch.full((B,), 999, dtype=torch.long))
    lg_r, bx_r, obj_r = om.head_base_stages(o.c, feats, boxes, temb, o.cfg)
    m._pack()
    lv = ops.Levels([f.permute(0, 2, 3, 1).contiguous().half().to(cuda) for f in feats])
    bd = ops.noise_to_boxes(x.to(cuda), 2.0, float(w), float(h))
    assert torch.equal(bd.cpu(), boxes)
    lg, bx, o32, o16 = m._base_stages(lv, bd, 999)
    assert rows_within(bx.cpu(), bx_r, 2e-3 * max(h, w)) >= 0.99
    assert rows_within(lg.cpu(), lg_r, 2e-2) >= 0.99
    assert rows_within(o32.cpu(), obj_r, 3e-2) >= 0.99
    # sanity vs the fp32 oracle: fp16 storage costs about 1e-2 on logits
    o32o = om.OracleDiffusionVID(sd, ocfg, fp16=False, noise=noise)
    lg_f, bx_f, _ = om.head_base_stages(o32o.c, o32o.backbone(imgs), boxes,
                                        om.time_embedding(o32o.c, torch.full((B,), 999, dtype=torch.long)), o32o.cfg)
    assert rows_within(lg.cpu(), lg_f, 1e-1) >= 0.97


def _run_clip(m, o, samples, h, w, L):
    fracs, counts_equal, n_out = [], 0, 0
    for s in samples:
        ref = o.forward(s)
        got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                     ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                     ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                     frame_id=s["frame_id"], start_id=0, end_id=s["end_id"], seg_len=L,
                     frame_category=s["frame_category"], video_id=0))
        assert len(got) == len(ref)
        for g, r in zip(got, ref):
            n_out += 1
            assert g.bbox.is_cuda and g.size == (w, h)
            counts_equal += int(len(g) == r["scores"].numel())
            fracs.append(match_fraction(g.bbox.cpu(), g.get_field("scores").cpu(), g.get_field("labels").cpu(),
                                        r["boxes"], r["scores"], r["labels"], max(h, w), box_tol=2e-3, score_tol=4e-3))
    return fracs, counts_equal, n_out


@pytest.mark.parametrize("T", [1, 4])
def test_clip_end_to_end_matches_oracle(cuda, T):
    """19-frame clip (ragged last batch of 3), 4 global frames, T=1 and T=4."""
    h, w, L = 192, 256, 19
    hp, sd, m, noise, ocfg = _models(T)
    o = om.OracleDiffusionVID(sd, ocfg, fp16=True, noise=noise)
    frames = synth.make_clip(L, h, w, seed=6)
    samples = synth.clip_samples(frames, [17, 3, 9, 12], h, w)
    fracs, counts_equal, n_out = _run_clip(m, o, samples, h, w, L)
    assert n_out == L
    fr = sorted(fracs)
    assert fr[len(fr) // 2] >= 0.95, fracs
    if T == 1:
        assert sum(fracs) / len(fracs) >= 0.95, fracs
        # the global memory (farthest-point sampled) of product and oracle hold the same rows up to fp16 noise
        pm_mem = m.proposal_feats_global[0].cpu()
        assert pm_mem.shape == o.mem[0].shape
    else:
        assert sum(f >= 0.95 for f in fracs) >= 0.5 * L, fracs


@pytest.mark.parametrize("T", [1, 4])
def test_clip_with_the_tcgen05_decoder_kernels(cuda, T, monkeypatch):
    """The opt-in tcgen05 variants of the decoder's small contractions inside the whole model: DynamicConv bmm pair
    (roi_dynconv_tc_kernel, hp dynconv_tc / DVID_DYNCONV_TC) and head-dim-32 attention (attention_tc_kernel,
    ops.ATTENTION_TC / DVID_ATTN_TC).  Same clip, same bar against the oracle as the default kernels."""
    monkeypatch.setattr(ops, "ATTENTION_TC", True)
    h, w, L = 192, 256, 19
    hp, sd, m, noise, ocfg = _models(T, hp_over=dict(dynconv_tc=1))
    assert m.dynconv_tc
    o = om.OracleDiffusionVID(sd, ocfg, fp16=True, noise=noise)
    frames = synth.make_clip(L, h, w, seed=6)
    samples = synth.clip_samples(frames, [17, 3, 9, 12], h, w)
    l0 = ops.LAUNCHES
    fracs, counts_equal, n_out = _run_clip(m, o, samples, h, w, L)
    assert n_out == L and ops.LAUNCHES > l0
    fr = sorted(fracs)
    assert fr[len(fr) // 2] >= 0.95, fracs
    if T == 1:
        assert sum(fracs) / len(fracs) >= 0.95, fracs


def test_skipping_the_unused_base_stages_changes_nothing(cuda):
    """hp skip_unused_base (dead-code elimination, off by default): with SAMPLE_STEP > 1 the t=999 base stages of the
    LOCAL frames are computed by the reference (diffusion_det.py:438-460) but read only by the SAMPLE_STEP == 1 branch
    (box_head.py:300-302).  Skipping them must give bit-identical detections, device-resident and host-fed."""
    h, w, L = 192, 256, 27
    outs = []
    for skip in (0, 1):
        for host in (False, True):
            hp, sd, m, noise, ocfg = _models(4, hp_over=dict(skip_unused_base=skip))
            m.host_results = host
            frames = synth.make_clip(L, h, w, seed=6)
            frames = frames.pin_memory() if host else frames.to(cuda)
            res = []
            l0 = ops.LAUNCHES
            for s in synth.clip_samples(frames, [17, 3, 9, 12], h, w):
                got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                             ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                             ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                             frame_id=s["frame_id"], start_id=0, end_id=s["end_id"], seg_len=L,
                             frame_category=s["frame_category"], video_id=0))
                res += [(b.bbox.cpu(), b.get_field("scores").cpu(), b.get_field("labels").cpu()) for b in got]
            outs.append((res, ops.LAUNCHES - l0))
    assert all(len(o[0]) == L for o in outs)
    for k in (0, 1):        # same feeding mode, with / without the dead work
        for (b0, s0, l0), (b1, s1, l1) in zip(outs[k][0], outs[2 + k][0]):
            assert torch.equal(b0, b1) and torch.equal(s0, s1) and torch.equal(l0, l1)
        assert outs[2 + k][1] < outs[k][1]            # and it really launches fewer kernels


def test_single_frame_config_without_global_memory(cuda):
    """BASELINE config[0] shape: vid_R_101_DiffusionDET.yaml semantics - 4 base heads, no cond head / memory, T=1,
    N=100, 2 frames of 300x300 (padded to 320x320)."""
    h, w, L = 300, 300, 2
    hp, sd, m, noise, ocfg = _models(1, hp_over=dict(num_heads=4, num_heads_local=0, global_enable=False))
    ocfg.update(num_heads=4, num_heads_local=0, global_enable=False)
    sd = synth.make_state_dict(seed=21, blocks=hp["blocks"], num_heads=4, num_heads_local=0, global_enable=False)
    m.load_state_dict(sd, strict=False)
    o = om.OracleDiffusionVID(sd, ocfg, fp16=True, noise=noise)
    frames = synth.make_clip(L, h, w, seed=2)
    assert frames.shape[-2:] == (320, 320)
    samples = synth.clip_samples(frames, [], h, w)
    fracs, counts_equal, n_out = _run_clip(m, o, samples, h, w, L)
    assert n_out == L and sum(fracs) / len(fracs) >= 0.95, fracs


@pytest.mark.parametrize("T", [1, 4])
def test_host_fed_pipeline_equals_device_resident_path(cuda, T):
    """Frames fed from (pinned) host memory go through the upload pipeline (backbone on the frames already on the
    device while the later ones are in flight, DiffusionDet._extract_pipelined); frames already in HBM go through the
    one-unit path.  Both must return the same detections (nothing in the backbone or the heads mixes frames)."""
    h, w, L = 192, 256, 19
    outs = []
    for host in (False, True):
        hp, sd, m, noise, ocfg = _models(T)
        m.host_results = host
        frames = synth.make_clip(L, h, w, seed=6)
        frames = frames.pin_memory() if host else frames.to(cuda)
        res = []
        for s in synth.clip_samples(frames, [17, 3, 9, 12], h, w):
            got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                         ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                         ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                         frame_id=s["frame_id"], start_id=0, end_id=s["end_id"], seg_len=L,
                         frame_category=s["frame_category"], video_id=0))
            res += [(b.bbox.cpu(), b.get_field("scores").cpu(), b.get_field("labels").cpu()) for b in got]
        outs.append(res)
        if host:
            assert m.io_bytes["h2d"] >= (L + 4) * 3 * h * w * 4
    assert len(outs[0]) == len(outs[1]) == L
    # the backbone sees different batch compositions on the two paths (tile-shape / split-K choices depend on the
    # number of tiles), so the results agree to fp16 noise, not bit for bit: same bar as against the oracle
    fracs = [match_fraction(b1, s1, l1, b0, s0, l0, max(h, w), box_tol=2e-3, score_tol=4e-3)
             for (b0, s0, l0), (b1, s1, l1) in zip(*outs)]
    fr = sorted(fracs)
    assert fr[len(fr) // 2] >= 0.95, fracs
    if T == 1:
        assert sum(fracs) / len(fracs) >= 0.95, fracs


def test_uint8_frames_equal_fp32_frames(cuda):
    """Clip-loader mode: the same clip delivered as decoded 8-bit frames (uint8 ImageLists in pinned
    host memory, ToTensor fused into dvid_preprocess_u8) must give exactly the detections of the reference protocol
    (ToTensor on the host, fp32 ImageLists) - the first kernel's output is bit-identical and everything after it runs
    on the same batch compositions - while a quarter of the bytes cross PCIe."""
    h, w, L = 192, 256, 19
    u8 = (synth.make_clip(L, h, w, seed=6) * 255.0).round().clamp(0, 255).to(torch.uint8)
    outs, sent = [], []
    for frames in (u8.to(torch.float32).div(255).pin_memory(), u8.pin_memory()):
        hp, sd, m, noise, ocfg = _models(4)
        m.host_results = True
        res = []
        for s in synth.clip_samples(frames, [17, 3, 9, 12], h, w):
            got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                         ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                         ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                         frame_id=s["frame_id"], start_id=0, end_id=s["end_id"], seg_len=L,
                         frame_category=s["frame_category"], video_id=0))
            res += [(b.bbox.cpu(), b.get_field("scores").cpu(), b.get_field("labels").cpu()) for b in got]
        outs.append(res)
        sent.append(m.io_bytes["h2d"])
    assert len(outs[0]) == len(outs[1]) == L
    assert sent[0] == 4 * sent[1] and sent[1] >= (L + 4) * 3 * h * w and sent[1] % (3 * h * w) == 0
    for (b0, s0, l0), (b1, s1, l1) in zip(*outs):
        assert torch.equal(b0, b1) and torch.equal(s0, s1) and torch.equal(l0, l1)


def test_stream_k_schedule_inside_the_model(cuda):
    """Opt-in stream-K schedule (hp streamk=1 / DVID_STREAMK=1): at 600x1000 the res4 3x3 convolutions of an 8-frame
    batch are 152 tiles on 148 SMs and take the stream-K variant inside the captured extract unit; the library switch
    is toggled around parallel stream branches.  Detections must agree with the default schedule to fp16 noise (the
    k-blocks of a shared tile are summed in two parts) and the switch must be off again for the next model."""
    h, w, L = 600, 1000, 9
    outs = []
    for sk in (0, 1):
        hp, sd, m, noise, ocfg = _models(1, hp_over=dict(streamk=sk, num_proposals=100))
        frames = synth.make_clip(L, h, w, seed=6).to(cuda)
        res = []
        for s in synth.clip_samples(frames, [3, 7], h, w):
            got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                         ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                         ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                         frame_id=s["frame_id"], start_id=0, end_id=s["end_id"], seg_len=L,
                         frame_category=s["frame_category"], video_id=0))
            res += [(b.bbox.cpu(), b.get_field("scores").cpu(), b.get_field("labels").cpu()) for b in got]
        outs.append(res)
    assert len(outs[0]) == len(outs[1]) == L
    fracs = sorted(match_fraction(b1, s1, l1, b0, s0, l0, max(h, w), box_tol=2e-3, score_tol=4e-3)
                   for (b0, s0, l0), (b1, s1, l1) in zip(*outs))
    assert fracs[len(fracs) // 2] >= 0.95, fracs


def test_deferred_host_results_equal_synchronous_ones(cuda):
    """host_results mode returns BoxLists that wait for their asynchronous device->host copy on first access
    (structures.BoxList.deferred): the call itself does not block, `.to("cpu")` does not block (the reference loop,
    engine/inference.py:75), a read blocks and yields exactly the detections of the synchronous copy; results stay
    valid after the pinned ring buffer has been reused by later batches."""
    h, w, L = 192, 256, 43                      # 6 key batches: the 4-deep ring of pinned buffers wraps
    outs = []
    for deferred in (False, True):
        hp, sd, m, noise, ocfg = _models(1)
        m.host_results = True
        m.deferred_results = deferred
        frames = synth.make_clip(L, h, w, seed=6).pin_memory()
        kept = []
        for s in synth.clip_samples(frames, [17, 3, 9, 12], h, w):
            got = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                         ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                         ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                         frame_id=s["frame_id"], start_id=0, end_id=s["end_id"], seg_len=L,
                         frame_category=s["frame_category"], video_id=0))
            if got:
                assert all(g.is_pending == deferred for g in got)
                moved = [g.to("cpu") for g in got]
                assert all(a is b for a, b in zip(moved, got)) or not deferred
                kept += moved
        assert len(kept) == L
        res = [(b.bbox.clone(), b.get_field("scores").clone(), b.get_field("labels").clone()) for b in kept]
        assert all(not b.is_pending and not b.bbox.is_cuda and len(b) == b.get_field("scores").numel() for b in kept)
        outs.append(res)
    for (b0, s0, l0), (b1, s1, l1) in zip(*outs):
        assert torch.equal(b0, b1) and torch.equal(s0, s1) and torch.equal(l0, l1)
    import pickle
    back = pickle.loads(pickle.dumps(kept[0]))
    assert torch.equal(back.bbox, kept[0].bbox) and back.size == kept[0].size and back.fields() == kept[0].fields()

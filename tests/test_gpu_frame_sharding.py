"""Frame sharding (SURVEY.md 8e mode B, BASELINE config 5) on the GPU kernels: the frames of one clip dealt to `world`
ranks must give, ON EVERY RANK, BoxLists that are BIT-IDENTICAL to the single-process run of the same clip (same
weights, same explicit noise) - the "8-GPU result == 1-GPU result" requirement of SURVEY.md 8(e).

That holds because every kernel on the path is batch-invariant: convolutions / GEMMs accumulate each output element
over K in the same k-block order whatever the number of frames or tiles (split-K factors are fixed per call site, the
tile WIDTH chosen by the wave cost model changes which CTA computes an element, not how), attention / top-k / NMS /
DDIM are per frame, ROIAlign + DynamicConv per box, and the farthest-point sampling runs redundantly on identical
gathered candidates.

Both granularities of DiffusionDet.set_frame_sharding are covered: "frames" (frame i of every call -> rank i % world, every
rank returns the whole clip) and "batches" (key batch k -> rank k % world, a rank returns its own batches).

Two variants:
  * two processes sharing cuda:0 over the `gloo` backend (device tensors, host transport) - runs on the one-GPU box;
  * two / four ranks, one GPU each, over NCCL - skipped unless that many devices are visible (gpurun --gpus N).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionvid_b200 import model as pm, structures, synth
from oracle import model as om

pytestmark = pytest.mark.gpu

H, W, L = 192, 256, 19          # ragged last batch of 3 frames
GIDX = [17, 3, 9, 12, 5]        # 5 global frames: uneven over 2 and 4 ranks
HP = dict(num_proposals=100, num_classes=30, hidden=256, nheads=8, dim_dynamic=64, dim_ff=2048, num_heads=3,
          num_heads_local=1, num_cls=1, num_reg=3, sample_step=4, snr_scale=2.0, use_nms=True, infer_batch=8,
          all_frame_interval=8, key_frame_location=0, global_enable=True, mem_size=300, mem_size2=50,
          topk=(75, 25), pixel_mean=(123.675, 116.280, 103.530), pixel_std=(58.395, 57.120, 57.375),
          blocks=(2, 2, 3, 2))


def _run(rank, world, T, port, out_dir, backend, shape=None, hp_over=None, mode="frames"):
    h, w, frames_n = shape or (H, W, L)
    dev = "cuda:%d" % (rank if backend == "nccl" else 0)
    torch.cuda.set_device(dev)
    hp = dict(HP, sample_step=T, device=dev, **(hp_over or {}))
    m = pm.DiffusionDet(hp)
    m.load_state_dict(synth.make_state_dict(seed=21, blocks=hp["blocks"]), strict=False)
    m.to(dev)
    m.noise = om.NoiseSource(9, hp["num_proposals"])
    if world > 1:
        dist.init_process_group(backend, init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        m.set_frame_sharding(rank, world, mode=mode)
    frames = synth.make_clip(frames_n, h, w, seed=6).to(dev)
    res = {}
    with torch.no_grad():
        for s in synth.clip_samples(frames, [g % frames_n for g in GIDX], h, w):
            out = m(dict(cur=structures.ImageList(s["cur"], [(h, w)]),
                         ref_l=[structures.ImageList(t, [(h, w)]) for t in s["ref_l"]],
                         ref_g=[structures.ImageList(t, [(h, w)]) for t in s["ref_g"]],
                         frame_id=s["frame_id"], start_id=0, end_id=frames_n - 1, seg_len=frames_n,
                         frame_category=s["frame_category"], video_id=0))
            for i, b in enumerate(out):       # a key call returns the frames frame_id .. frame_id + len(out) - 1
                res[s["frame_id"] + i] = (b.bbox.cpu(), b.get_field("scores").cpu(), b.get_field("labels").cpu())
    torch.save(dict(res=res, mem=m.proposal_feats_global[0].cpu(), comm=dict(m.comm_bytes)),
               os.path.join(out_dir, "r%d_w%d.pt" % (rank, world)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _check(tmp_path, world, frames_n=L, mode="frames"):
    ref = torch.load(os.path.join(tmp_path, "r0_w1.pt"))
    assert sorted(ref["res"]) == list(range(frames_n))
    seen = {}
    for rank in range(world):
        got = torch.load(os.path.join(tmp_path, "r%d_w%d.pt" % (rank, world)))
        if mode == "frames":
            assert sorted(got["res"]) == list(range(frames_n))       # every rank holds the whole clip
            assert got["comm"]["results"] > 0
        else:
            # whole key batches: batch k (frames 8k .. 8k+7) is returned by rank k % world and by nobody else
            assert sorted(got["res"]) == [f for f in range(frames_n) if (f // 8) % world == rank], sorted(got["res"])
            assert got["comm"]["results"] == 0
        assert torch.equal(got["mem"], ref["mem"])                   # same farthest-point picks, same bits
        for i, (gb, gs, gl) in got["res"].items():
            rb, rs, rl = ref["res"][i]
            assert gb.shape == rb.shape, (rank, i, gb.shape, rb.shape)
            assert torch.equal(gb, rb) and torch.equal(gs, rs) and torch.equal(gl, rl), (rank, i)
            seen[i] = seen.get(i, 0) + 1
        assert got["comm"]["memory"] > 0
    assert sorted(seen) == list(range(frames_n))


@pytest.mark.parametrize("mode", ["frames", "batches"])
@pytest.mark.parametrize("T", [1, 4])
def test_two_ranks_on_one_gpu_bit_identical_to_one_rank(cuda, tmp_path, T, mode):
    _run(0, 1, T, 0, str(tmp_path), "gloo")
    mp.spawn(_run, args=(2, T, _free_port(), str(tmp_path), "gloo", None, None, mode), nprocs=2, join=True)
    _check(str(tmp_path), 2, mode=mode)


@pytest.mark.parametrize("mode", ["frames", "batches"])
@pytest.mark.parametrize("world", [2, 4])
def test_nccl_ranks_bit_identical_to_one_rank(cuda, tmp_path, world, mode):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs (gpurun --gpus %d)" % (world, world))
    _run(0, 1, 4, 0, str(tmp_path), "nccl")
    mp.spawn(_run, args=(world, 4, _free_port(), str(tmp_path), "nccl", None, None, mode), nprocs=world, join=True)
    _check(str(tmp_path), world, mode=mode)


def test_batch_invariance_at_the_headline_shape(cuda, tmp_path):
    """Same property at 1000x600 with the full R-101 (152 res4 tiles for 8 frames, 76 for 4: different tile widths and
    wave counts in the two runs), one 8-frame key batch + 4 global frames, two ranks on one GPU."""
    shape = (600, 1000, 8)
    hp_full = dict(blocks=(3, 4, 23, 3), num_proposals=300, mem_size=200, mem_size2=60)
    _run(0, 1, 4, 0, str(tmp_path), "gloo", shape, hp_full)
    mp.spawn(_run, args=(2, 4, _free_port(), str(tmp_path), "gloo", shape, hp_full), nprocs=2, join=True)
    _check(str(tmp_path), 2, 8)

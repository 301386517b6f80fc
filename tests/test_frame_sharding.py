"""Multi-rank frame sharding (SURVEY.md 8e mode B) on CPU: world_size-2 `gloo` processes, the test-only op shim standing
in for the CUDA library, must return - on EVERY rank - the detections of the single-process run of the same clip
(same weights, same explicit noise).  Covers the per-video all-gather of memory candidates, the per-batch result
all-reduce, a ragged last batch (fewer frames than ranks own evenly) and T=1 / T=4."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionvid_b200 import model as pm, structures, synth
from oracle import model as om
from tests import cpu_ops_shim
from tests.parity_util import match_fraction
from tests.test_host_logic import SMALL

H, W, L = 64, 96, 11
GIDX = [9, 2, 5]


def _run(rank, world, T, port, out_dir, mode="frames"):
    real_ops, real_threads = pm.ops, torch.get_num_threads()
    try:
        _run_impl(rank, world, T, port, out_dir, mode)
    finally:            # the world == 1 run happens inside the pytest process: do not leak the shim to other tests
        pm.ops = real_ops
        torch.set_num_threads(real_threads)


def _run_impl(rank, world, T, port, out_dir, mode="frames"):
    torch.set_num_threads(1)       # same CPU kernels in every process: results must not depend on the world size
    pm.ops = cpu_ops_shim
    hp = dict(SMALL, sample_step=T)
    m = pm.DiffusionDet(hp)
    m.load_state_dict(synth.make_state_dict(seed=11, blocks=hp["blocks"]), strict=False)
    m.noise = om.NoiseSource(3, hp["num_proposals"])
    if world > 1:
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
        m.set_frame_sharding(rank, world, mode=mode)
    frames = synth.make_clip(L, H, W, seed=5)
    res = []
    own_frames = []
    for s in synth.clip_samples(frames, GIDX, H, W):
        out = m(dict(cur=structures.ImageList(s["cur"], [(H, W)]),
                     ref_l=[structures.ImageList(t, [(H, W)]) for t in s["ref_l"]],
                     ref_g=[structures.ImageList(t, [(H, W)]) for t in s["ref_g"]],
                     frame_id=s["frame_id"], start_id=0, end_id=L - 1, seg_len=L, frame_category=s["frame_category"],
                     video_id=0))
        res += [(b.bbox.clone(), b.get_field("scores").clone(), b.get_field("labels").clone()) for b in out]
        own_frames += [s["frame_id"] + i for i in range(len(out))]
    torch.save(dict(res=res, frames=own_frames, mem=m.proposal_feats_global[0]),
               os.path.join(out_dir, "r%d_w%d.pt" % (rank, world)))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("T", [1, 4])
def test_two_rank_frame_sharding_matches_single_process(tmp_path, T):
    _run(0, 1, T, 0, str(tmp_path))
    mp.spawn(_run, args=(2, T, _free_port(), str(tmp_path)), nprocs=2, join=True)
    ref = torch.load(os.path.join(tmp_path, "r0_w1.pt"))
    for rank in (0, 1):
        got = torch.load(os.path.join(tmp_path, "r%d_w2.pt" % rank))
        assert len(got["res"]) == len(ref["res"]) == L
        assert got["mem"].shape == ref["mem"].shape
        # the farthest-point-sampled memory is replicated: same rows on every rank as in the single-process run
        assert (got["mem"] - ref["mem"]).abs().max().item() <= 1e-4
        fr = []
        for (gb, gs, gl), (rb, rs, rl) in zip(got["res"], ref["res"]):
            fr.append(match_fraction(gb, gs, gl, rb, rs, rl, max(H, W), box_tol=1e-3, score_tol=1e-3))
        assert min(fr) >= 0.9 and sum(fr) / len(fr) >= 0.97, fr
    a = torch.load(os.path.join(tmp_path, "r0_w2.pt"))["res"]
    b = torch.load(os.path.join(tmp_path, "r1_w2.pt"))["res"]
    for x, y in zip(a, b):        # both ranks hold the identical all-reduced result
        assert all(torch.equal(p, q) for p, q in zip(x, y))


def test_two_rank_key_batch_sharding_matches_single_process(tmp_path):
    """mode="batches": key batch k of the clip belongs to rank k % 2 and is returned by that rank only; together the two
    ranks hold the single-process detections (L = 11: batch 0 = frames 0..7 on rank 0, batch 1 = frames 8..10 on rank 1);
    the replicated global memory equals the single-process one."""
    T = 4
    _run(0, 1, T, 0, str(tmp_path))
    mp.spawn(_run, args=(2, T, _free_port(), str(tmp_path), "batches"), nprocs=2, join=True)
    ref = torch.load(os.path.join(tmp_path, "r0_w1.pt"))
    got = [torch.load(os.path.join(tmp_path, "r%d_w2.pt" % r)) for r in (0, 1)]
    assert got[0]["frames"] == list(range(8)) and got[1]["frames"] == list(range(8, L))
    for g in got:
        assert (g["mem"] - ref["mem"]).abs().max().item() <= 1e-4
        fr = [match_fraction(gb, gs, gl, *ref["res"][f], max(H, W), box_tol=1e-3, score_tol=1e-3)
              for f, (gb, gs, gl) in zip(g["frames"], g["res"])]
        assert min(fr) >= 0.9 and sum(fr) / len(fr) >= 0.97, fr

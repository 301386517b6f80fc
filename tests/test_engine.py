"""Inference driver: loop bookkeeping, tensor-format prediction gather over a world_size-2 gloo group,
and the predictions.pth artefact (mega_core/engine/inference.py:22-116,161-168)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from diffusionvid_b200 import engine
from diffusionvid_b200.structures import BoxList, ImageList


def _boxlist(seed, n, size=(96, 64)):
    g = torch.Generator().manual_seed(seed)
    xy = torch.rand(n, 2, generator=g) * 40
    b = BoxList(torch.cat([xy, xy + 1 + torch.rand(n, 2, generator=g) * 20], 1), size, mode="xyxy")
    b.add_field("scores", torch.rand(n, generator=g))
    b.add_field("labels", torch.randint(1, 31, (n,), generator=g))
    return b


def _all_predictions(n_images=13):
    return {i: _boxlist(100 + i, (i * 7) % 5 if i != 4 else 0, (96 + i, 64)) for i in range(n_images)}


def _same(a, b):
    return (a.size == b.size and torch.equal(a.bbox, b.bbox) and torch.equal(a.get_field("scores"), b.get_field("scores"))
            and torch.equal(a.get_field("labels"), b.get_field("labels"))
            and b.get_field("labels").dtype == torch.int64)


def test_pack_roundtrip_with_empty_images():
    preds = _all_predictions()
    back = engine.unpack_predictions(*engine.pack_predictions(preds))
    assert sorted(back) == sorted(preds)
    assert all(_same(preds[k], back[k]) for k in preds)
    assert engine.pack_predictions({})[1].shape == (0, 4)


def _worker(rank, world, port, out_dir):
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    full = _all_predictions()
    # contiguous per-rank id ranges like VIDTestDistributedSampler (samplers/distributed.py:83-95); rank 1 owns more
    mine = {k: v for k, v in full.items() if (k < 5) == (rank == 0)}
    res = engine.accumulate_predictions(mine)
    if rank == 0:
        engine.save_predictions(res, out_dir)
    else:
        assert res is None
    dist.barrier()
    dist.destroy_process_group()


def test_gather_world2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = torch.load(os.path.join(str(tmp_path), "predictions.pth"), weights_only=False)
    full = _all_predictions()
    assert isinstance(got, list) and len(got) == len(full)
    assert all(_same(full[i], got[i]) for i in range(len(full)))


def test_single_process_and_noncontiguous_warning(caplog):
    full = _all_predictions(6)
    del full[3]
    with caplog.at_level("WARNING", logger="diffusionvid_b200.inference"):
        res = engine.accumulate_predictions(full)
    assert len(res) == 5 and "not a contiguous set" in caplog.text


class _FakeModel(torch.nn.Module):
    """Returns one BoxList per call at frame_category 0/1 like the reference's clip protocol; records what it saw."""

    def __init__(self):
        super().__init__()
        self.seen = []

    def forward(self, images):
        assert not self.training
        self.seen.append((images["frame_id"], images["cur"].tensors.device.type, len(images["ref_l"])))
        return [_boxlist(images["frame_id"], 3)]


def test_compute_on_dataset_maps_ids_and_moves_to_cpu():
    def loader():
        for i in range(4):
            img = ImageList(torch.zeros(1, 3, 8, 8), [(8, 8)])
            yield (dict(cur=img, ref_l=[ImageList(torch.zeros(1, 3, 8, 8), [(8, 8)])], ref_g=[], frame_id=i), None,
                   [[10 + i]])
    m = _FakeModel().train()
    out = engine.compute_on_dataset(m, loader(), torch.device("cpu"))
    assert sorted(out) == [10, 11, 12, 13]
    assert [s[0] for s in m.seen] == [0, 1, 2, 3]
    assert all(_same(out[10 + i], _boxlist(i, 3)) for i in range(4))

"""Inference driver around the model call: the loop, the cross-rank prediction gather and the
``predictions.pth`` artefact, mirroring ``mega_core/engine/inference.py`` for ``MODEL.VID.METHOD = "diffusion"``.

  compute_on_dataset            inference.py:22-93    batches of (images, targets, image_ids); images is the clip dict
                                                      (cur / ref_l / ref_g ImageLists + frame bookkeeping); the result
                                                      maps ``image_ids[0][i]`` to the i-th returned BoxList (on the CPU)
  accumulate_predictions        inference.py:96-116   merge the per-rank dicts on the main process, warn when the ids
                                                      are not contiguous, return the list ordered by id
  save_predictions / inference  inference.py:161-168  ``torch.save(list[BoxList], <output_folder>/predictions.pth)``

The reference gathers pickled BoxLists through byte tensors (``comm.py:54-94``).  Here every rank packs its detections
into four flat tensors (per-image header, boxes, scores, labels) and the exchange is three ``all_gather`` calls of
padded tensors on the process group's own device - NCCL over NVLink for GPU jobs, gloo on CPU - with no pickling.
"""
import logging
import os

import torch
import torch.distributed as dist

from .structures import BoxList

_HDR = 4   # per image: id, number of boxes, width, height


def compute_on_dataset(model, data_loader, device, timer=None):
    """Runs ``model(images)`` over the loader (inference.py:22-93, "diffusion" branch without seq-NMS)."""
    model.eval()
    results = {}
    cpu = torch.device("cpu")
    for images, _targets, image_ids in data_loader:
        with torch.no_grad():
            if timer:
                timer.tic()
            images["cur"] = images["cur"].to(device)
            for key in ("ref", "ref_l", "ref_m", "ref_g"):
                if key in images:
                    images[key] = [img.to(device) for img in images[key]]
            output = model(images)
            if timer:
                if torch.device(device).type != "cpu":
                    torch.cuda.synchronize()
                timer.toc()
            output = [o.to(cpu) for o in output]
        results.update({img_id: result for img_id, result in zip(image_ids[0], output)})
    return results


def pack_predictions(predictions):
    """dict {image id: BoxList(xyxy, fields scores/labels)} -> (header int64 [n,4], boxes f32 [m,4], scores f32 [m],
    labels int64 [m]) in ascending id order."""
    ids = sorted(predictions)
    hdr = torch.zeros(len(ids), _HDR, dtype=torch.int64)
    boxes, scores, labels = [], [], []
    for i, k in enumerate(ids):
        b = predictions[k]
        if b.mode != "xyxy":
            b = b.convert("xyxy")
        hdr[i] = torch.tensor([int(k), len(b), int(b.size[0]), int(b.size[1])])
        boxes.append(b.bbox.reshape(-1, 4).float().cpu())
        scores.append(b.get_field("scores").reshape(-1).float().cpu())
        labels.append(b.get_field("labels").reshape(-1).long().cpu())
    cat = lambda xs, shape, dt: torch.cat(xs) if xs else torch.zeros(shape, dtype=dt)   # noqa: E731
    return hdr, cat(boxes, (0, 4), torch.float32), cat(scores, (0,), torch.float32), cat(labels, (0,), torch.int64)


def unpack_predictions(hdr, boxes, scores, labels, into=None):
    out = {} if into is None else into
    o = 0
    for k, n, w, h in hdr.tolist():
        b = BoxList(boxes[o:o + n].clone(), (w, h), mode="xyxy")
        b.add_field("scores", scores[o:o + n].clone())
        b.add_field("labels", labels[o:o + n].clone())
        out[k] = b
        o += n
    return out


def _all_gather_ragged(t, dev, group=None):
    """all_gather of tensors whose first dimension differs per rank (padded to the maximum)."""
    world = dist.get_world_size(group)
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=dev)
    ns = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(ns, n, group=group)
    ns = [int(x) for x in ns]
    pad = torch.zeros((max(ns + [1]),) + tuple(t.shape[1:]), dtype=t.dtype, device=dev)
    pad[:t.shape[0]] = t.to(dev)
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)
    return [b[:k].cpu() for b, k in zip(bufs, ns)]


def accumulate_predictions(predictions_per_rank, group=None):
    """inference.py:96-116.  Returns the id-ordered list on the main process (rank 0) and None elsewhere; a
    single-process run returns the list directly.  Later ranks win on duplicate ids, as ``dict.update`` does there."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" \
            else torch.device("cpu")
        hdr, boxes, scores, labels = pack_predictions(predictions_per_rank)
        # boxes, scores and labels travel as one f32 [m,6] tensor (labels are 1..num_classes: exact in f32)
        rows = torch.cat([boxes, scores[:, None], labels[:, None].float()], dim=1)
        hdrs = _all_gather_ragged(hdr, dev, group)
        rowss = _all_gather_ragged(rows, dev, group)
        if dist.get_rank(group) != 0:
            return None
        predictions = {}
        for h, r in zip(hdrs, rowss):
            unpack_predictions(h, r[:, :4], r[:, 4], r[:, 5].long(), into=predictions)
    else:
        predictions = dict(predictions_per_rank)
    image_ids = sorted(predictions)
    if image_ids and len(image_ids) != image_ids[-1] + 1:
        logging.getLogger("diffusionvid_b200.inference").warning(
            "Number of images that were gathered from multiple processes is not a contiguous set. "
            "Some images might be missing from the evaluation")
    return [predictions[i] for i in image_ids]


def save_predictions(predictions, output_folder, name="predictions.pth"):
    """inference.py:165-168: the artefact ``tools/test_prediction.py`` and ``vid_eval`` read back."""
    os.makedirs(output_folder, exist_ok=True)
    path = os.path.join(output_folder, name)
    torch.save(predictions, path)
    return path


def inference(model, data_loader, device="cuda", output_folder=None, timer=None, evaluate=False, logger=None):
    """inference.py:119-168: run the loader, gather on rank 0, write predictions.pth and - with `evaluate` and a
    dataset that carries ground truth (get_img_info / get_groundtruth / map_class_id_to_class_name) - the VID AP50 /
    CorLoc report of data/datasets/evaluation/vid/vid_eval.py (diffusionvid_b200/evaluation.py).  Returns the
    prediction list, or (predictions, metrics) when evaluating."""
    device = torch.device(device)
    predictions = compute_on_dataset(model, data_loader, device, timer)
    if dist.is_available() and dist.is_initialized():
        dist.barrier()
    predictions = accumulate_predictions(predictions)
    if predictions is None:
        return None
    if output_folder:
        save_predictions(predictions, output_folder)
    if evaluate:
        from . import evaluation
        # matching on the GPU (one dvid_vid_match launch over all frames) when the model ran on one
        return predictions, evaluation.do_vid_evaluation(data_loader.dataset, predictions, output_folder, logger,
                                                         device=device if device.type == "cuda" else "cpu")
    return predictions

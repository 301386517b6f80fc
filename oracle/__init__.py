"""CPU restatement oracle of the DiffusionVID inference hot path.

TEST INFRASTRUCTURE ONLY. Nothing under diffusionvid_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or the
timed CPU baseline - never as a compute path of the product.

PARITY PINNING: the reference (sdroh1027/DiffusionVID @ 8375542) cannot be imported or built in this container
(torch._six, THC, apex, detectron2, timm, yacs, fvcore are all absent) and it ships no test
or golden vector for any DiffusionDet code, so the model-level restatement in oracle/model.py is "parity unpinned"
by the reference itself.  What *is* pinned:
  * oracle.ops.roi_align / batched_nms / mha  against torchvision.ops.roi_align(aligned=True), torchvision.ops.
    batched_nms and torch.nn.MultiheadAttention (the exact library calls the reference makes through detectron2),
  * oracle.legacy (C restatement of mega_core/csrc/cpu/nms_cpu.cpp) against the golden vectors of the reference's
    tests/test_nms.py (committed under tests/golden/ with the script that extracted them),
  * oracle.ops.fps against a literal emulation of mega_core/csrc/cuda/fps.cu's thread-strided scan + tree reduction.
"""

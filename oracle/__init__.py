"""CPU restatement oracle of the DiffusionVID inference hot path.

TEST INFRASTRUCTURE ONLY. Nothing under diffusionvid_b200/ may import this package; only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it, and only as the checker or the
timed CPU baseline - never as a compute path of the product.

PARITY PINNING: the reference (sdroh1027/DiffusionVID @ 8375542) cannot be imported as a package or built in this
container (torch._six, THC, apex, detectron2, timm, yacs, fvcore are all absent) and it ships no
test or golden vector for any DiffusionDet code.  The pins are therefore:
  * oracle.model against OUTPUTS OF THE REFERENCE'S OWN SOURCES run here on the CPU: tests/golden/make_golden.py loads
    the unmodified diffusion_det.py / box_head.py / loss.py / structures/*.py by path (third-party imports satisfied by
    the real torchvision ops and inert placeholders) and tests/test_golden_reference.py checks the oracle against the
    committed results (tests/golden/ref_diffusionvid_small.pt);
  * oracle.ops.roi_align / batched_nms / mha against torchvision.ops.roi_align(aligned=True), torchvision.ops.
    batched_nms and torch.nn.MultiheadAttention (the exact library calls the reference makes through detectron2);
  * oracle.legacy.nms_legacy (mega_core/csrc/cpu/nms_cpu.cpp) against the vectors of the reference's tests/test_nms.py
    (tests/golden/nms_vectors.json; tests/golden/extract_nms_vectors.py executes that test file against it);
  * oracle.ops.fps is a literal emulation of mega_core/csrc/cuda/fps.cu's thread-strided scan + tree reduction -
    restatement only ("parity unpinned": that CUDA kernel cannot be built or run here), as is the detectron2
    R-101+FPN backbone in oracle.model.resnet_fpn (source absent from /root/reference).
"""

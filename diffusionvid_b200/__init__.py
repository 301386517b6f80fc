"""diffusionvid_b200 — B200-native (sm_100a) implementation of the DiffusionVID inference hot path.

Host code is Python/PyTorch (device memory, streams, torch.distributed); all compute is hand-written CUDA behind the
C ABI declared in include/dvid_b200.h (libdvid_b200.so, loaded by diffusionvid_b200._lib). There is no CPU fallback:
using any operator without the built extension raises.
"""
__version__ = "0.1.0"

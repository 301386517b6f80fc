"""ctypes loader for libdvid_b200.so (the C ABI in include/dvid_b200.h).

The library is built in-tree by build.sh / __graft_entry__.build(). Loading fails loudly when it is missing: the
product path has no CPU or PyTorch fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_C", "libdvid_b200.so")

ERRORS = {1: "DVID_ERR_SHAPE", 2: "DVID_ERR_CUDA", 3: "DVID_ERR_DRIVER", 4: "DVID_ERR_ARG"}

_lib = None


class DvidError(RuntimeError):
    pass


def lib():
    """Return the loaded CDLL (loading it on first use)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise DvidError(
                f"{LIB_PATH} not found: build it with ./build.sh (or __graft_entry__.build()). "
                "diffusionvid_b200 has no fallback path.")
        _lib = ctypes.CDLL(LIB_PATH)
    return _lib


def check(code, what):
    if code != 0:
        raise DvidError(f"{what} failed: {ERRORS.get(code, code)}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

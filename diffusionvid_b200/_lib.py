"""ctypes loader for libdvid_b200.so (the C ABI in include/dvid_b200.h).

The library is built in-tree by build.sh / __graft_entry__.build(). Loading fails loudly when it is missing: the
product path has no CPU or PyTorch fallback.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DVID_LIB_PATH") or os.path.join(_HERE, "_C", "libdvid_b200.so")   # override: A/B of builds

ERRORS = {1: "DVID_ERR_SHAPE", 2: "DVID_ERR_CUDA", 3: "DVID_ERR_DRIVER", 4: "DVID_ERR_ARG"}

_lib = None

P = ctypes.c_void_p
I = ctypes.c_int
L = ctypes.c_long
F = ctypes.c_float

# name -> argtypes; every entry point returns int (dvid_resize_workspace_bytes: long).  Kept in the order of include/dvid_b200.h.
SIGNATURES = {
    "dvid_abi_version": [],
    "dvid_conv_streamk": [I],
    "dvid_num_sms": [],
    "dvid_conv2d_nhwc_f16": [P, P, P, P, P, I, I, I, I, I, I, I, I, I, I, I, P],
    "dvid_stem_conv_f16": [P, P, P, P, I, I, I, I, I, P],
    "dvid_gemm_f16": [P, P, P, P, P, P, I, I, I, I, I, P, P],
    "dvid_preprocess": [P, P, I, I, I, I, I, I, P, P, P],
    "dvid_preprocess_u8": [P, P, I, I, I, I, I, I, P, P, P],
    "dvid_maxpool3x3s2_nhwc_f16": [P, P, I, I, I, I, P],
    "dvid_attention_hd32": [P, P, P, P, I, I, I, I, L, L, L, L, L, L, L, L, P],
    "dvid_attention_hd32_tc": [P, P, P, P, I, I, I, I, L, L, L, L, L, L, L, L, P],
    "dvid_roi_align": [P, P, P, P, P, I, I, P, P, P, P],
    "dvid_roi_dynconv": [P, P, P, P, P, I, I, P, P, P, P, P, P, P, P],
    "dvid_roi_dynconv_tc": [P, P, P, P, P, I, I, P, P, P, P, P, P, P, P],
    "dvid_gemm256_row": [P, P, P, P, P, P, I, P, P, I, P],
    "dvid_row_post": [P, I, L, P, P, P, P, I, P, P, P, I, I, P, P, P, P, I, I, I, I, P, I, P],
    "dvid_small_linear": [P, P, P, P, I, I, I, I, I, P],
    "dvid_time_sinusoid": [P, P, P, I, P],
    "dvid_head_final": [P, I, P, I, P, I, P, P, P, P, I, P],
    "dvid_head_tail": [P, P, P, P, P, P, I, P, P, P, P, P, P, P, P, P, P, P, P, P, P, I, P],
    "dvid_noise_to_boxes": [P, P, I, F, F, F, P],
    "dvid_ddim_step": [P, I, P, P, P, P, P, P, P, I, I, F, F, F, F, F, F, F, F, P],
    "dvid_topk_scores": [P, P, I, I, I, I, P, P, P, I, I, P],
    "dvid_topk_mask": [P, I, I, I, I, I, P, P, P],
    "dvid_gather_masked_rows": [P, P, I, I, I, P, P],
    "dvid_nms": [P, P, P, P, I, I, I, F, I, I, I, F, F, P, P, P, P, P, P, L, P],
    "dvid_cdist_f32": [P, P, I, I, P],
    "dvid_furthest_point_sampling": [I, I, I, P, P, P, P],
    "dvid_roi_align_legacy_forward": [P, P, I, I, I, I, F, I, I, I, P, P],
    "dvid_vid_match": [P, P, P, P, P, P, P, P, I, F, ctypes.c_double, P, P, P, P],
    "dvid_swin_rows": [P, I, P, I, P, P, P, P, I, I, I, I, I, I, P],
    "dvid_swin_patch_merge": [P, I, I, I, I, P, P, P, P],
    "dvid_swin_patch_gather": [P, P, I, I, I, P, P, P],
    "dvid_swin_patch_gather_u8": [P, P, I, I, I, P, P, P],
    "dvid_resize_workspace_bytes": [I, I, I, I, I],
    "dvid_resize_bilinear_u8": [P, I, I, I, I, I, P, I, I, P, L, P],
    "dvid_swin_window_attention": [P, P, P, I, I, I, I, I, I, P],
    "dvid_jpeg_info": [P, L, P, P],
    "dvid_jpeg_decode_rgb": [P, L, P, I, I, P],
    "dvid_swin_window_attention_tc": [P, P, P, I, I, I, I, I, I, P],
}


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
        cdll = ctypes.CDLL(LIB_PATH)
        for name, argtypes in SIGNATURES.items():
            fn = getattr(cdll, name)      # AttributeError here = header / library mismatch: fail loudly
            fn.argtypes = argtypes
            fn.restype = ctypes.c_long if name == "dvid_resize_workspace_bytes" else ctypes.c_int
        _lib = cdll
    return _lib


def check(code, what):
    if code != 0:
        raise DvidError(f"{what} failed: {ERRORS.get(code, code)}")


def ptr(t):
    """Device pointer of a torch tensor (or None) as c_void_p."""
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def cur_stream():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

"""The drop-in boundary is a C ABI (include/dvid_b200.h -> diffusionvid_b200/_C/libdvid_b200.so).  CPU-only checks: the
library loads without a GPU, exports every entry point the header declares, the ctypes binding table
(diffusionvid_b200/_lib.py) covers exactly those entry points with the right arity, and no torch / C++ types leak
into the signatures.  No compute call is made here (the -m gpu tests call through the same table)."""
import ctypes
import os
import re

import pytest

from diffusionvid_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dvid_b200.h")


def _declarations():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)          # drop comments
    decls = {}
    for m in re.finditer(r"DVID_API\s+(?:int|long)\s+(\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = (n, args)
    return decls


def test_header_declares_the_hot_path_entry_points():
    d = _declarations()
    assert len(d) >= 30
    for name in ("dvid_conv2d_nhwc_f16", "dvid_gemm_f16", "dvid_preprocess", "dvid_preprocess_u8", "dvid_roi_dynconv",
                 "dvid_head_tail", "dvid_ddim_step", "dvid_topk_scores", "dvid_nms", "dvid_furthest_point_sampling",
                 "dvid_swin_window_attention", "dvid_conv_streamk"):
        assert name in d, name


def test_library_loads_and_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        pytest.fail("libdvid_b200.so is not built (run ./build.sh or __graft_entry__.build())")
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _declarations():
        assert hasattr(lib, name), "header declares %s but the library does not export it" % name
    lib.dvid_abi_version.restype = ctypes.c_int
    assert lib.dvid_abi_version() >= 9


def test_binding_table_matches_the_header():
    d = _declarations()
    assert set(_lib.SIGNATURES) == set(d), (sorted(set(d) - set(_lib.SIGNATURES)), sorted(set(_lib.SIGNATURES) - set(d)))
    for name, argtypes in _lib.SIGNATURES.items():
        assert len(argtypes) == d[name][0], (name, len(argtypes), d[name])


def test_signatures_are_plain_c():
    for name, (_, args) in _declarations().items():
        for bad in ("torch", "at::", "std::", "Tensor", "&", "cudaStream_t"):
            assert bad not in args, (name, bad)
    src = open(HEADER).read()
    assert 'extern "C"' in src

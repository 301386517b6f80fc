#!/usr/bin/env python
"""Extract the known-answer vectors of the reference's tests/test_nms.py by EXECUTING that file (unmodified) with
`mega_core.layers.nms` bound to a recorder, and write them to tests/golden/nms_vectors.json.

The reference asserts `np.sort(box_nms(boxes, scores, thresh)) == gt`; the recorder returns the oracle's answer and a
patched numpy.testing.assert_array_equal captures the expected array of every call, so the fixture holds
(boxes, scores, thresh, expected) exactly as the reference's test states them."""
import importlib.util
import json
import os
import sys
import types
import unittest

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle.legacy import nms_legacy  # noqa: E402

REF = os.environ.get("DVID_REFERENCE", "/root/reference")
calls, cases = [], []


def recorder(boxes, scores, thresh):
    calls.append((boxes.numpy().copy(), scores.numpy().copy(), float(thresh)))
    return torch.from_numpy(nms_legacy(boxes.numpy(), scores.numpy(), thresh))


layers = types.ModuleType("mega_core.layers"); layers.nms = recorder
pkg = types.ModuleType("mega_core"); pkg.__path__ = []; pkg.layers = layers
sys.modules["mega_core"] = pkg; sys.modules["mega_core.layers"] = layers

real_assert = np.testing.assert_array_equal


def capture(actual, expected, *a, **k):
    b, s, t = calls[-1]
    cases.append(dict(boxes=b.tolist(), scores=s.tolist(), thresh=t, expected=np.asarray(expected).tolist()))
    return real_assert(actual, expected, *a, **k)


np.testing.assert_array_equal = capture
spec = importlib.util.spec_from_file_location("ref_test_nms", os.path.join(REF, "tests", "test_nms.py"))
mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
res = unittest.TextTestRunner(verbosity=0).run(unittest.defaultTestLoader.loadTestsFromModule(mod))
assert res.wasSuccessful(), "the oracle's legacy NMS fails the reference's own test"
out = os.path.join(HERE, "nms_vectors.json")
json.dump(dict(source="sdroh1027/DiffusionVID@8375542 tests/test_nms.py", cases=cases), open(out, "w"))
print("wrote", out, len(cases), "cases")

"""numpy restatement of the reference's legacy native NMS (mega_core._C.nms).  TEST INFRASTRUCTURE ONLY.

Follows mega_core/csrc/cpu/nms_cpu.cpp:5-65 (CPU: areas and overlaps with the legacy +1 pixel convention, suppress when
`ovr >= threshold`, return the surviving ORIGINAL indices in ascending order) and mega_core/csrc/cuda/nms.cu:13-21,
:112-130 (CUDA: same IoU, suppress when `> threshold`).  Pinned by the reference's own tests/test_nms.py vectors
(tests/golden/nms_vectors.json, extracted by tests/golden/extract_nms_vectors.py which executes that test file)."""
import numpy as np


def nms_legacy(dets, scores, threshold, ge=True):
    dets = np.asarray(dets, dtype=np.float32)
    scores = np.asarray(scores, dtype=np.float32)
    n = dets.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    x1, y1, x2, y2 = dets[:, 0], dets[:, 1], dets[:, 2], dets[:, 3]
    areas = (x2 - x1 + np.float32(1)) * (y2 - y1 + np.float32(1))
    order = np.argsort(-scores, kind="stable")
    suppressed = np.zeros(n, dtype=np.uint8)
    thr = np.float32(threshold)
    for _i in range(n):
        i = order[_i]
        if suppressed[i]:
            continue
        for _j in range(_i + 1, n):
            j = order[_j]
            if suppressed[j]:
                continue
            w = max(np.float32(0), min(x2[i], x2[j]) - max(x1[i], x1[j]) + np.float32(1))
            h = max(np.float32(0), min(y2[i], y2[j]) - max(y1[i], y1[j]) + np.float32(1))
            inter = np.float32(w * h)
            ovr = inter / (areas[i] + areas[j] - inter)
            if (ovr >= thr) if ge else (ovr > thr):
                suppressed[j] = 1
    return np.nonzero(suppressed == 0)[0].astype(np.int64)

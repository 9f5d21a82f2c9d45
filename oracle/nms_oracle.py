"""Oracle for FCOS post-processing NMS.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

The reference calls ``torchvision.ops.boxes.batched_nms(boxes, scores, labels, 0.3)``
(fcos_utils/fcos.py:635).  torchvision is a third-party dependency (pinned 0.11.3 by
scripts/init_env.sh:25, 0.26.0 installed); its CPU kernel is only present as a binary,
so the published algorithm is restated here in numpy float32:

* ``torchvision/csrc/ops/cpu/nms_kernel.cpp``: stable descending sort of the scores,
  greedy scan, a box j is suppressed by a kept box i when
  ``inter / (area_i + area_j - inter) > iou_threshold`` with every quantity a float32
  and the comparison carried out against the threshold as a double.
* ``torchvision/ops/boxes.py:51-121``: ``batched_nms`` uses the "coordinate trick"
  (boxes shifted by ``label * (max_coord + 1)``) when ``boxes.numel() <= 4000`` on the
  CPU and per-class NMS followed by a descending score sort above that.

Pinned against ``torchvision.ops.nms`` / ``batched_nms`` in tests/test_oracle_golden.py.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def nms_order(scores: np.ndarray) -> np.ndarray:
    """Stable descending order of float32 scores (ties keep ascending index)."""
    scores = np.asarray(scores, dtype=F32)
    # stable sort on the negated key keeps equal scores in index order
    return np.argsort(-scores.astype(np.float64), kind="stable")


def nms(boxes: np.ndarray, scores: np.ndarray, iou_threshold: float) -> np.ndarray:
    """Greedy NMS, float32 arithmetic, returns kept indices in score-descending order."""
    boxes = np.ascontiguousarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    n = boxes.shape[0]
    if n == 0:
        return np.zeros((0,), dtype=np.int64)
    order = nms_order(scores)
    b = boxes[order]
    x1, y1, x2, y2 = b[:, 0], b[:, 1], b[:, 2], b[:, 3]
    areas = ((x2 - x1).astype(F32) * (y2 - y1).astype(F32)).astype(F32)
    suppressed = np.zeros(n, dtype=bool)
    thr = np.float64(iou_threshold)
    keep = []
    for i in range(n):
        if suppressed[i]:
            continue
        keep.append(order[i])
        if i + 1 == n:
            break
        xx1 = np.maximum(x1[i], x1[i + 1:])
        yy1 = np.maximum(y1[i], y1[i + 1:])
        xx2 = np.minimum(x2[i], x2[i + 1:])
        yy2 = np.minimum(y2[i], y2[i + 1:])
        w = np.maximum(F32(0), (xx2 - xx1).astype(F32))
        h = np.maximum(F32(0), (yy2 - yy1).astype(F32))
        inter = (w * h).astype(F32)
        denom = ((areas[i] + areas[i + 1:]).astype(F32) - inter).astype(F32)
        with np.errstate(divide="ignore", invalid="ignore"):
            ovr = (inter / denom).astype(F32)
        suppressed[i + 1:] |= ovr.astype(np.float64) > thr
    return np.asarray(keep, dtype=np.int64)


def coordinate_trick_boxes(boxes: np.ndarray, labels: np.ndarray) -> np.ndarray:
    """boxes + label * (max + 1), all in float32 (torchvision/ops/boxes.py:87-104)."""
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    max_coordinate = boxes.max()
    offsets = (labels.astype(F32) * F32(max_coordinate + F32(1))).astype(F32)
    return (boxes + offsets[:, None]).astype(F32)


def batched_nms(boxes, scores, labels, iou_threshold: float, cpu_switch_numel: int = 4000) -> np.ndarray:
    """Class-aware NMS with torchvision's CPU strategy switch."""
    boxes = np.asarray(boxes, dtype=F32).reshape(-1, 4)
    scores = np.asarray(scores, dtype=F32).reshape(-1)
    labels = np.asarray(labels).reshape(-1)
    if boxes.size == 0:
        return np.zeros((0,), dtype=np.int64)
    if boxes.size > cpu_switch_numel:
        keep_mask = np.zeros(scores.shape[0], dtype=bool)
        for c in np.unique(labels):
            idx = np.nonzero(labels == c)[0]
            k = nms(boxes[idx], scores[idx], iou_threshold)
            keep_mask[idx[k]] = True
        keep_idx = np.nonzero(keep_mask)[0]
        return keep_idx[nms_order(scores[keep_idx])]
    return nms(coordinate_trick_boxes(boxes, labels), scores, iou_threshold)

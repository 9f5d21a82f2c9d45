"""Oracle for HandNet.forward (SURVEY.md section 8a rows S1, S2 and the pipeline glue).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Restates handnet_pipeline/handnet_pipeline.py:58-116 over the FCOS and A2J oracles.
Pinned against the real reference by oracle/make_golden.py -> tests/golden/*.pt.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import numpy as np
import torch

from . import a2j_oracle, fcos_oracle

CROP = 176          # handnet_pipeline.py:101
PAD_FRAC = 0.4      # handnet_pipeline.py:93
F32 = np.float32


def pad_box(box_xyxy: np.ndarray, img_h: int, img_w: int) -> np.ndarray:
    """handnet_pipeline.py:88-97.  The box is truncated to int64; ``percent * w`` is a python
    float times a 0-dim int64 tensor, which torch evaluates in float32; ``box[k] - that`` is
    float32 as well; python ``max(0, t)`` / ``min(W, t)`` keep ``t`` unless the bound wins; the
    assignment back into the int64 tensor truncates toward zero."""
    b = np.trunc(np.asarray(box_xyxy, dtype=F32)).astype(np.int64)
    w = b[2] - b[0]
    h = b[3] - b[1]
    pw = F32(F32(PAD_FRAC) * F32(w))
    ph = F32(F32(PAD_FRAC) * F32(h))
    out = b.copy()
    v = F32(F32(b[0]) - pw)
    out[0] = int(np.trunc(v)) if v > 0 else 0
    v = F32(F32(b[1]) - ph)
    out[1] = int(np.trunc(v)) if v > 0 else 0
    v = F32(F32(b[2]) + pw)
    out[2] = int(np.trunc(v)) if v < img_w else img_w
    v = F32(F32(b[3]) + ph)
    out[3] = int(np.trunc(v)) if v < img_h else img_h
    return out


def nearest_src_index(dst: int, in_size: int, out_size: int) -> int:
    """ATen upsample_nearest (legacy 'nearest'): min(floor(dst * float32(in/out)), in - 1)."""
    scale = F32(in_size) / F32(out_size)
    return min(int(np.floor(F32(dst) * scale)), in_size - 1)


def crop_resize(depth: torch.Tensor, box: np.ndarray, out: int = CROP) -> torch.Tensor:
    """handnet_pipeline.py:101: depth[:, y1:y2+1, x1:x2+1] -> F.interpolate(size=(176,176)),
    default mode 'nearest'.  depth is [C, H, W]; python slicing clamps the inclusive end."""
    c, hh, ww = depth.shape
    x1, y1, x2, y2 = (int(v) for v in box)
    ys, ye = min(max(y1, 0), hh), min(max(y2 + 1, 0), hh)
    xs, xe = min(max(x1, 0), ww), min(max(x2 + 1, 0), ww)
    ih, iw = ye - ys, xe - xs
    if ih <= 0 or iw <= 0:
        raise ValueError("empty crop")
    ry = torch.tensor([ys + nearest_src_index(d, ih, out) for d in range(out)])
    rx = torch.tensor([xs + nearest_src_index(d, iw, out) for d in range(out)])
    return depth[:, ry][:, :, rx]


def handnet_forward(fcos_sd: Dict[str, torch.Tensor], a2j_sd: Dict[str, torch.Tensor],
                    images: Sequence[torch.Tensor], depth_images: torch.Tensor, num_classes: int = 3,
                    min_size: int = 800, max_size: int = 1333, emulate_bf16: bool = False,
                    detections: Optional[List[Dict[str, torch.Tensor]]] = None, max_hands: int = 1):
    """Ensemble branch of HandNet.forward (is_detect=False, is_3D=False).

    Returns (final_results [B,21,3], depth_batch [n,1,176,176], crops [n,4] int64, hit_mask [B]).
    Where the reference raises on mixed hit/miss batches (torch.stack over a list holding
    None, handnet_pipeline.py:111) the oracle keeps only the hits.

    ``max_hands = H > 1`` is NOT reference behaviour (the reference keeps ``boxes[:1]``,
    handnet_pipeline.py:84-85): it restates the extension of the B200 path (BASELINE.json config 5,
    "up to 4 hands/frame") as the reference's own per-box steps (:88-102, then A2J) applied to
    ``boxes[:H]``; final_results is then [B,H,21,3], hit_mask [B,H], and depth_batch / crops list the
    hits in (frame, hand) order.  Column 0 of every output equals the max_hands = 1 result."""
    if detections is None:
        detections = fcos_oracle.fcos_forward(fcos_sd, images, num_classes, ext=False, min_size=min_size,
                                              max_size=max_size, emulate_bf16=emulate_bf16)
    bsz = len(images)
    final = torch.zeros((bsz, max_hands, 21, 3))
    hit = torch.zeros((bsz, max_hands), dtype=torch.bool)
    crops, depth_batch = [], []
    for i, det in enumerate(detections):
        boxes = det["boxes"][det["labels"] == num_classes - 1]
        h, w = images[i].shape[-2:]
        for k in range(min(max_hands, boxes.shape[0])):
            box = pad_box(boxes[k].numpy(), int(h), int(w))
            hit[i, k] = True
            crops.append(torch.from_numpy(box))
            depth_batch.append(crop_resize(depth_images[i], box))
    if max_hands == 1:
        final, hit = final[:, 0], hit[:, 0]
    if not depth_batch:
        return final, torch.zeros_like(depth_images), torch.zeros((bsz, 4)), hit
    depth_batch = torch.stack(depth_batch)
    crops = torch.stack(crops)
    final[hit] = a2j_oracle.a2j_forward(a2j_sd, depth_batch, emulate_bf16=emulate_bf16)
    return final, depth_batch, crops, hit

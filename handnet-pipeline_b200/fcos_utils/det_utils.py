"""Box coder of the FCOS detector (reference: fcos_utils/det_utils.py:220-294).

Only the inference half (``decode_single``) is on the hot path; on the GPU it is fused into
``hn_fcos_decode_select``.  This class keeps the reference's name and call signature for callers
that decode on the host, and is what the CUDA kernel is tested against.  The training-only coders
and matchers of the reference file (BoxCoder, Matcher, SSDMatcher, samplers) are out of scope.
"""
from __future__ import annotations

import torch
from torch import Tensor


class BoxLinearCoder:
    """Distances (left, top, right, bottom) from an anchor centre, optionally in units of the anchor size."""

    def __init__(self, normalize_by_size: bool = True) -> None:
        self.normalize_by_size = normalize_by_size

    def decode_single(self, rel_codes: Tensor, boxes: Tensor) -> Tensor:
        boxes = boxes.to(rel_codes.dtype)
        cx = 0.5 * (boxes[:, 0] + boxes[:, 2])
        cy = 0.5 * (boxes[:, 1] + boxes[:, 3])
        if self.normalize_by_size:
            w = boxes[:, 2] - boxes[:, 0]
            h = boxes[:, 3] - boxes[:, 1]
            rel_codes = rel_codes * torch.stack((w, h, w, h), dim=1)
        return torch.stack((cx - rel_codes[:, 0], cy - rel_codes[:, 1], cx + rel_codes[:, 2], cy + rel_codes[:, 3]),
                           dim=1)

    def encode_single(self, reference_boxes: Tensor, proposals: Tensor) -> Tensor:
        cx = 0.5 * (reference_boxes[:, 0] + reference_boxes[:, 2])
        cy = 0.5 * (reference_boxes[:, 1] + reference_boxes[:, 3])
        t = torch.stack((cx - proposals[:, 0], cy - proposals[:, 1], proposals[:, 2] - cx, proposals[:, 3] - cy), dim=1)
        if self.normalize_by_size:
            w = reference_boxes[:, 2] - reference_boxes[:, 0]
            h = reference_boxes[:, 3] - reference_boxes[:, 1]
            t = t / torch.stack((w, h, w, h), dim=1)
        return t

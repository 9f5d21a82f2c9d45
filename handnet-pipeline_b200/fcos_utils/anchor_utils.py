"""Anchor generator of the FCOS detector (reference: fcos_utils/anchor_utils.py:10-132).

FCOS uses one square anchor per feature-map cell whose centre is the FCOS "point".  On the GPU the
anchors are never materialised: ``hn_fcos_decode_select`` regenerates them from the cell index.  This
module keeps the reference's class for host-side callers and for the parity tests.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
from torch import Tensor, nn


class AnchorGenerator(nn.Module):
    def __init__(self, sizes=((128, 256, 512),), aspect_ratios=((0.5, 1.0, 2.0),)):
        super().__init__()
        if not isinstance(sizes[0], (list, tuple)):
            sizes = tuple((s,) for s in sizes)
        if not isinstance(aspect_ratios[0], (list, tuple)):
            aspect_ratios = (aspect_ratios,) * len(sizes)
        assert len(sizes) == len(aspect_ratios)
        self.sizes = sizes
        self.aspect_ratios = aspect_ratios
        self.cell_anchors = [self.generate_anchors(s, a) for s, a in zip(sizes, aspect_ratios)]

    @staticmethod
    def generate_anchors(scales, aspect_ratios, dtype=torch.float32, device="cpu") -> Tensor:
        scales = torch.as_tensor(scales, dtype=dtype, device=device)
        ratios = torch.as_tensor(aspect_ratios, dtype=dtype, device=device)
        hr = torch.sqrt(ratios)
        wr = 1 / hr
        ws = (wr[:, None] * scales[None, :]).view(-1)
        hs = (hr[:, None] * scales[None, :]).view(-1)
        return (torch.stack([-ws, -hs, ws, hs], dim=1) / 2).round()

    def num_anchors_per_location(self) -> List[int]:
        return [len(s) * len(a) for s, a in zip(self.sizes, self.aspect_ratios)]

    def grid_anchors(self, grid_sizes: Sequence[Sequence[int]], strides: Sequence[Sequence[int]]) -> List[Tensor]:
        out = []
        for (gh, gw), (sh, sw), base in zip(grid_sizes, strides, self.cell_anchors):
            xs = torch.arange(0, gw, dtype=torch.int32, device=base.device) * int(sw)
            ys = torch.arange(0, gh, dtype=torch.int32, device=base.device) * int(sh)
            yy, xx = torch.meshgrid(ys, xs, indexing="ij")
            shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
            out.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
        return out

    def forward(self, image_list, feature_maps: List[Tensor]) -> List[Tensor]:
        """image_list: anything with ``.tensors`` ([B,C,H,W]) and ``.image_sizes``."""
        grid_sizes = [fm.shape[-2:] for fm in feature_maps]
        image_size = image_list.tensors.shape[-2:]
        dtype, device = feature_maps[0].dtype, feature_maps[0].device
        strides = [[image_size[0] // g[0], image_size[1] // g[1]] for g in grid_sizes]
        self.cell_anchors = [c.to(dtype=dtype, device=device) for c in self.cell_anchors]
        per_level = self.grid_anchors(grid_sizes, strides)
        return [torch.cat(per_level) for _ in range(len(image_list.image_sizes))]

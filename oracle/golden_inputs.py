"""Seeded input generators shared by oracle/make_golden.py and the tests.

TEST INFRASTRUCTURE ONLY.  Fixtures store seeds, not inputs; both sides regenerate the inputs
with these functions (torch CPU generators are deterministic for a given torch build).
"""
from __future__ import annotations

import torch


def sample(t, n=4096):
    """Evenly strided sample of a tensor, enough to pin it without committing megabytes."""
    f = t.reshape(-1)
    step = max(1, f.numel() // n)
    return f[::step].clone(), torch.tensor([f.numel(), step])


def inputs_images(seed, n, h, w):
    g = torch.Generator().manual_seed(seed)
    return [torch.rand(3, h, w, generator=g) for _ in range(n)]


def stress_head_tensors(seed, batch, locs, num_classes=3, mu=-0.35):
    """BASELINE.json config 4: raise logits so ~half the locations pass 0.7.  (The generator itself lives
    in hn_b200.synth so that bench.py can build the workload without importing oracle/.)"""
    from hn_b200 import synth
    return synth.stress_head_tensors(seed, batch, locs, num_classes, mu)


def pad_crop_inputs(seed, nb, hh, ww):
    """Boxes (one per frame) and depth maps for the S1/S2 pad + crop cases."""
    g = torch.Generator().manual_seed(seed)
    ctr = torch.rand(nb, 2, generator=g) * torch.tensor([ww, hh])
    half = torch.rand(nb, 2, generator=g) * torch.tensor([ww * 0.6, hh * 0.6]) + 0.3
    boxes = torch.cat((ctr - half, ctr + half), dim=1)
    boxes[0] = torch.tensor([10.2, 5.7, 10.9, 6.1])              # collapses to zero width after truncation
    boxes[1] = torch.tensor([-20.5, -3.2, 50.9, 40.0])           # negative corner (truncation toward zero)
    boxes[2] = torch.tensor([0.0, 0.0, float(ww), float(hh)])    # whole image
    boxes[3] = torch.tensor([100.0, 60.0, 400.0, 300.0])         # overhangs the far corner
    depth = torch.rand(nb, 1, hh, ww, generator=g)
    return boxes, depth

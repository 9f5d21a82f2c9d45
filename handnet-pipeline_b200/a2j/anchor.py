"""A2J anchors and anchor post-process (reference: a2j/anchor.py).

``generate_anchors`` / ``shift`` are one-time host code and keep the reference's numpy semantics;
``post_process.forward`` (anchor.py:57-82) runs as one fused CUDA kernel (``hn_a2j_aggregate``): softmax over
the anchors of each joint and the weighted (anchor + offset) / depth sums, for the whole batch at once.
``A2J_loss`` is training-only and out of scope (a stub keeps the buffers for state-dict compatibility).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from hn_b200 import ops


def generate_anchors(P_h=None, P_w=None):
    """In-cell anchor offsets, row i*len + j = (P_h[i], P_w[j])  (anchor.py:7-24)."""
    P_h = np.array([2, 6, 10, 14]) if P_h is None else np.asarray(P_h)
    P_w = np.array([2, 6, 10, 14]) if P_w is None else np.asarray(P_w)
    a = np.zeros((len(P_h) * len(P_h), 2))
    a[:, 0] = np.repeat(P_h, len(P_h))[: a.shape[0]]
    a[:, 1] = np.tile(P_w, len(P_w))[: a.shape[0]]
    return a


def shift(shape, stride, anchors):
    """All anchors of a feature map: cell k = w*shape[0] + h holds anchors + (stride*h, stride*w)  (anchor.py:26-42)."""
    sh = np.arange(0, shape[0]) * stride
    sw = np.arange(0, shape[1]) * stride
    gh, gw = np.meshgrid(sh, sw)                        # 'xy' indexing: gh[w, h] = sh[h]
    shifts = np.stack((gh.ravel(), gw.ravel()), axis=1)
    return (anchors[None, :, :] + shifts[:, None, :]).reshape(-1, 2)


class post_process(nn.Module):
    def __init__(self, P_h=[2, 6], P_w=[2, 6], shape=[48, 26], stride=8, thres=8, is_3D=True):
        super().__init__()
        anchors = generate_anchors(P_h=P_h, P_w=P_w)
        self.register_buffer("all_anchors", torch.from_numpy(shift(shape, stride, anchors)).float())
        self.register_buffer("thres", torch.from_numpy(np.array(thres)).float())
        self.is_3D = is_3D

    def forward(self, heads, voting=False):
        """heads = (cls [n,A,J], reg [n,A,J,2], depth [n,A,J]) on the device -> [n, J, 3] on the device
        (``[n, J, 2]`` when ``is_3D`` is False).  ``voting`` is unused, as in the reference."""
        if self.is_3D:
            cls, reg, dep = heads
        else:
            cls, reg = heads
            dep = torch.zeros_like(cls)
        out = ops.a2j_aggregate(cls.contiguous().float(), reg.contiguous().float(), dep.contiguous().float(),
                                self.all_anchors.float().contiguous())
        return out if self.is_3D else out[..., :2]


class A2J_loss(nn.Module):
    """Training loss of the reference (anchor.py:84-153): buffers only, for checkpoint compatibility."""

    def __init__(self, P_h=[2, 6], P_w=[2, 6], shape=[8, 4], stride=8, thres=[10.0, 20.0], spatialFactor=0.1,
                 img_shape=[0, 0], is_3D=True):
        super().__init__()
        anchors = generate_anchors(P_h=P_h, P_w=P_w)
        self.register_buffer("all_anchors", torch.from_numpy(shift(shape, stride, anchors)).float())
        self.register_buffer("thres", torch.from_numpy(np.array(thres)).float())
        self.spatialFactor, self.img_shape, self.is_3D = spatialFactor, img_shape, is_3D

    def forward(self, heads, annotations):
        raise NotImplementedError("A2J training loss is outside the scope of the B200 inference build")

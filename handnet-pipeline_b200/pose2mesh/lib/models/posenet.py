"""2D -> 3D pose lifter of pose2mesh (the residual MLP of Martinez et al.) on the fp32 linear kernel.

Call surface and parameter names of /root/reference/pose2mesh/lib/models/posenet.py (LinearModel :44-93, Linear :11-41,
get_model :95-98) so that its checkpoints load unchanged.  Forward (eval mode): y = w1(x); per stage y = y + w2(relu(bn2(
w1(relu(bn1(y)))))) (:25-39; dropout is the identity in eval mode); out = w2(y) (:75-87; batch_norm1 of the top module is a
parameter the reference never applies, :62,77).  Each BatchNorm + ReLU runs as the input transform of the linear layer that
follows it, the residual add as its epilogue: five launches, all weight streaming (4096 x 4096 fp32 = 67 MB per layer)."""
from __future__ import annotations

import torch
import torch.nn as nn

from hn_b200 import ops

from .backbones.cheby_graph_conv import bn_affine


class Linear(nn.Module):
    def __init__(self, linear_size, p_dropout=0.5):
        super().__init__()
        self.l_size = linear_size
        self.w1 = nn.Linear(linear_size, linear_size)
        self.batch_norm1 = nn.BatchNorm1d(linear_size)
        self.w2 = nn.Linear(linear_size, linear_size)
        self.batch_norm2 = nn.BatchNorm1d(linear_size)

    def forward(self, x):
        h = ops.linear_f32([x], _w(self.w1), _b(self.w1), in_affine=bn_affine(self.batch_norm1))
        return ops.linear_f32([h], _w(self.w2), _b(self.w2), in_affine=bn_affine(self.batch_norm2), res=x)


def _w(lin):
    return lin.weight.detach().float().contiguous()


def _b(lin):
    return None if lin.bias is None else lin.bias.detach().float().contiguous()


class LinearModel(nn.Module):
    def __init__(self, num_joint, linear_size=4096, num_stage=2, p_dropout=0.5, pretrained=False):
        super().__init__()
        self.linear_size, self.p_dropout, self.num_stage = linear_size, p_dropout, num_stage
        self.input_size, self.output_size = num_joint * 2, num_joint * 3
        self.w1 = nn.Linear(self.input_size, linear_size)
        self.batch_norm1 = nn.BatchNorm1d(linear_size)             # (kept for checkpoint compatibility; unused, as upstream)
        self.linear_stages = nn.ModuleList([Linear(linear_size, p_dropout) for _ in range(num_stage)])
        self.w2 = nn.Linear(linear_size, self.output_size)
        if pretrained:
            raise NotImplementedError("load the PoseNet weights through FlatPose2Mesh.load_state_dict (ros_demo.py:146-147)")

    def forward(self, x):
        if self.training:
            raise NotImplementedError("pose2mesh on the B200 build is inference only: call .eval()")
        y = ops.linear_f32([x.contiguous().float()], _w(self.w1), _b(self.w1))
        for st in self.linear_stages:
            y = st(y)
        return ops.linear_f32([y], _w(self.w2), _b(self.w2))


def get_model(num_joint, hid_dim, num_layer, p_dropout, pretrained=False):
    return LinearModel(num_joint, hid_dim, num_layer, p_dropout, pretrained)

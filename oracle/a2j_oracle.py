"""Oracle for the A2J pose network.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Functional torch-CPU restatement driven by a state dict with the reference's parameter
names (a2j/a2j.py:212-224 builds the modules).  ``emulate_bf16`` as in fcos_oracle.
Pinned against the real reference by oracle/make_golden.py -> tests/golden/*.pt.
"""
from __future__ import annotations

from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # nn.BatchNorm2d default (a2j/resnet.py:34, a2j/a2j.py:52)


def _q(x, emulate):
    return x.to(torch.bfloat16).to(torch.float32) if emulate else x


def _bn(sd, prefix):
    w, b = sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float()
    rm, rv = sd[prefix + ".running_mean"].float(), sd[prefix + ".running_var"].float()
    scale = w * (rv + BN_EPS).rsqrt()
    return scale, b - rm * scale


def _conv_bn(sd, x, conv, bn, stride, pad, dil, emulate, relu, residual=None):
    scale, shift = _bn(sd, bn)
    if (conv + ".bias") in sd:                       # tower convs carry a bias (a2j/a2j.py:50)
        shift = shift + scale * sd[conv + ".bias"].float()
    y = F.conv2d(x, _q(sd[conv + ".weight"].float(), emulate), None, stride=stride, padding=pad, dilation=dil)
    y = y * scale[None, :, None, None] + shift[None, :, None, None]
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    return _q(y, emulate)


def backbone(sd, x, emulate: bool, channel_in: int = 1, prefix: str = "Backbone.model."):
    """ResNetBackBone.forward (a2j/a2j.py:194-210) over a2j/resnet.py's ResNet-50:
    Bottleneck with the stride on the 3x3 (resnet.py:68); layer4 has stride 1 and
    dilation 2 on blocks 1-2 only (resnet.py:112,142,145).  Returns (C4, C5)."""
    x = x[:, 0:channel_in]
    w = sd[prefix + "conv1.weight"].float()
    if channel_in == 1:
        # expand(n,3,h,w) of one depth channel == convolving with the weights summed over Cin
        x = x.expand(x.shape[0], 3, x.shape[2], x.shape[3])
    s, b = _bn(sd, prefix + "bn1")
    y = F.conv2d(x, w, None, stride=2, padding=3)     # stem runs on the fp32 crop with fp32 weights
    y = _q(F.relu(y * s[None, :, None, None] + b[None, :, None, None]), emulate)
    y = F.max_pool2d(y, kernel_size=3, stride=2, padding=1)
    c4 = None
    for li, nblocks in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nblocks):
            p = f"{prefix}layer{li}.{bi}."
            stride = 2 if (li in (2, 3) and bi == 0) else 1
            dil = 2 if (li == 4 and bi > 0) else 1
            identity = y
            if (p + "downsample.0.weight") in sd:
                identity = _conv_bn(sd, y, p + "downsample.0", p + "downsample.1", stride, 0, 1, emulate, relu=False)
            t = _conv_bn(sd, y, p + "conv1", p + "bn1", 1, 0, 1, emulate, relu=True)
            t = _conv_bn(sd, t, p + "conv2", p + "bn2", stride, dil, dil, emulate, relu=True)
            y = _conv_bn(sd, t, p + "conv3", p + "bn3", 1, 0, 1, emulate, relu=True, residual=identity)
        if li == 3:
            c4 = y
    return c4, y


def tower(sd, x, prefix: str, emulate: bool):
    """4 x (conv3x3 + bias -> BN -> ReLU) then conv3x3 + bias (a2j/a2j.py:70-84)."""
    for i in range(1, 5):
        x = _conv_bn(sd, x, f"{prefix}conv{i}", f"{prefix}bn{i}", 1, 1, 1, emulate, relu=True)
    return F.conv2d(x, _q(sd[prefix + "output.weight"].float(), emulate), sd[prefix + "output.bias"].float(), padding=1)


def heads(sd, c4, c5, emulate: bool, num_classes: int = 21, num_anchors: int = 16):
    """a2j/a2j.py:85-89,131-135,178-181: permute(0,3,2,1) then view as
    [n, W*H*A, joints(,2)] -> anchor index (w*H + h)*A + a, channel a*J + j (*2 + xy)."""
    cls = tower(sd, c4, "classificationModel.", emulate).permute(0, 3, 2, 1).contiguous()
    reg = tower(sd, c5, "regressionModel.", emulate).permute(0, 3, 2, 1).contiguous()
    dep = tower(sd, c5, "DepthRegressionModel.", emulate).permute(0, 3, 2, 1).contiguous()
    n = cls.shape[0]
    return (cls.view(n, -1, num_classes), reg.view(n, -1, num_classes, 2), dep.view(n, -1, num_classes))


def all_anchors(shape=(11, 11), stride: int = 16, p=(2, 6, 10, 14)) -> torch.Tensor:
    """generate_anchors + shift (a2j/anchor.py:7-42) in closed form:
    anchors[(w*H + h)*16 + i*4 + j] = (stride*h + p[i], stride*w + p[j])."""
    hh, ww = shape
    out = np.zeros((ww, hh, len(p), len(p), 2), dtype=np.float64)
    for w in range(ww):
        for h in range(hh):
            for i, pi in enumerate(p):
                for j, pj in enumerate(p):
                    out[w, h, i, j, 0] = stride * h + pi
                    out[w, h, i, j, 1] = stride * w + pj
    return torch.from_numpy(out.reshape(-1, 2)).float()


def aggregate(cls: torch.Tensor, reg: torch.Tensor, dep: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """post_process.forward (a2j/anchor.py:57-82): softmax over anchors per joint, weighted
    sum of (anchor + offset) and of depth -> [n, joints, 3]."""
    w = F.softmax(cls, dim=1)                                   # [n, A, J]
    xy = (w.unsqueeze(-1) * (anchors[None, :, None, :] + reg)).sum(1)
    d = (w * dep).sum(1)
    return torch.cat((xy, d.unsqueeze(-1)), dim=-1)


def a2j_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, emulate_bf16: bool = False, channel_in: int = 1,
                return_taps: bool = False):
    """A2JModel.forward with gt=None (a2j/a2j.py:243-250)."""
    c4, c5 = backbone(sd, x.float(), emulate_bf16, channel_in)
    cls, reg, dep = heads(sd, c4, c5, emulate_bf16)
    anchors = sd["post_process.all_anchors"].float() if "post_process.all_anchors" in sd else all_anchors()
    out = aggregate(cls, reg, dep, anchors)
    if return_taps:
        return out, {"c4": c4, "c5": c5, "cls": cls, "reg": reg, "dep": dep}
    return out


# ------------------------------------------------------------------------------ J5
def uvd2xyz(uvd: np.ndarray, paras: np.ndarray) -> np.ndarray:
    """datasets3d/a2jdataset.py:31-38: pinhole back-projection, paras = (fx, fy, cx, cy)."""
    paras = np.asarray(paras)
    out = np.array(uvd, copy=True).reshape(-1, 3)
    out[:, :2] = (out[:, :2] - paras[2:]) * out[:, 2:] / paras[:2]
    return out.reshape(np.shape(uvd)).astype(np.float32)


def convert_joints(jt_uvd_pred, box, paras=None, crop_w: int = 176, crop_h: int = 176):
    """convert_joints with jt_uvd_gt=None (a2j/a2j.py:17-43): crop space -> image pixels,
    optional back-projection to millimetres."""
    p = np.asarray(jt_uvd_pred).reshape(-1, 3)
    box = np.asarray(box).reshape(4)
    x0, y0, x1, y1 = box[0], box[1], box[2], box[3]
    out = np.ones_like(p)
    out[:, 0] = p[:, 0] * (x1 - x0) / crop_w + x0
    out[:, 1] = p[:, 1] * (y1 - y0) / crop_h + y0
    out[:, 2] = p[:, 2]
    if paras is not None:
        out = uvd2xyz(out, np.asarray(paras).reshape(4)) * 1000.0
    return out

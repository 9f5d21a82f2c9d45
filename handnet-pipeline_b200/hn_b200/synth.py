"""Deterministic synthetic ("random-init") weights with the reference's state-dict keys.

There is no network for the pretrained checkpoints (scripts/download_models.sh:4-8 in the
reference), so benchmarks and parity tests run on seeded random weights of the reference
architectures.  The generators below write out the key/shape tables directly:

* FCOS: torchvision ``resnet_fpn_backbone('resnet34', returned_layers=[2,3,4])`` + FCOSHead
  (fcos_utils/fcos.py:476, 216-264, 343-371) -- 232 keys for ``ext=False``.
* A2J: a2j/resnet.py ResNet-50 under ``Backbone.model.`` + three towers + anchor buffers
  (a2j/a2j.py:212-224) -- 414 keys.

Initialisation follows the reference constructors in spirit (kaiming fan-out for the ResNets,
N(0, 0.01) for the FCOS heads, xavier for the A2J towers) but with non-trivial norm statistics
so that the folded scale/shift paths are exercised, and with the head biases raised so that
detections pass the hard-coded 0.7 score cut (at the reference's own init nothing passes and
HandNet returns early).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def _kaiming(g, cout, cin, k):
    std = math.sqrt(2.0 / (cout * k * k))
    return torch.randn((cout, cin, k, k), generator=g) * std


def _xavier(g, cout, cin, k):
    std = math.sqrt(2.0 / ((cin + cout) * k * k))
    return torch.randn((cout, cin, k, k), generator=g) * std


def _bn(sd, g, prefix, c, num_batches_tracked=False, gamma=1.0):
    sd[prefix + ".weight"] = (0.7 + 0.6 * torch.rand(c, generator=g)) * gamma
    sd[prefix + ".bias"] = 0.05 * torch.randn(c, generator=g)
    sd[prefix + ".running_mean"] = 0.05 * torch.randn(c, generator=g)
    sd[prefix + ".running_var"] = 0.7 + 0.6 * torch.rand(c, generator=g)
    if num_batches_tracked:
        sd[prefix + ".num_batches_tracked"] = torch.tensor(0, dtype=torch.int64)


def fcos_state_dict(num_classes: int = 3, ext: bool = False, seed: int = 0,
                    cls_bias=None, ctr_bias: float = 3.0, reg_bias: float = 2.0,
                    cls_weight_std: float = 0.05) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    p = "backbone.body."
    sd[p + "conv1.weight"] = _kaiming(g, 64, 3, 7)
    _bn(sd, g, p + "bn1", 64)
    inpl = 64
    for li, (planes, nblocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for bi in range(nblocks):
            q = f"{p}layer{li}.{bi}."
            sd[q + "conv1.weight"] = _kaiming(g, planes, inpl, 3)
            _bn(sd, g, q + "bn1", planes)
            sd[q + "conv2.weight"] = _kaiming(g, planes, planes, 3)
            _bn(sd, g, q + "bn2", planes, gamma=0.5)        # keeps the residual stream bounded
            if bi == 0 and li > 1:
                sd[q + "downsample.0.weight"] = _kaiming(g, planes, inpl, 1)
                _bn(sd, g, q + "downsample.1", planes)
            inpl = planes
    f = "backbone.fpn."
    for i, cin in enumerate((128, 256, 512)):
        sd[f"{f}inner_blocks.{i}.0.weight"] = _kaiming(g, 256, cin, 1) * 0.7
        sd[f"{f}inner_blocks.{i}.0.bias"] = 0.02 * torch.randn(256, generator=g)
    for i in range(3):
        sd[f"{f}layer_blocks.{i}.0.weight"] = _kaiming(g, 256, 256, 3) * 0.7
        sd[f"{f}layer_blocks.{i}.0.bias"] = 0.02 * torch.randn(256, generator=g)
    for hp in ("head.classification_head.", "head.regression_head."):
        for i in range(4):
            sd[f"{hp}conv.{3 * i}.weight"] = torch.randn((256, 256, 3, 3), generator=g) * 0.02
            sd[f"{hp}conv.{3 * i}.bias"] = 0.02 * torch.randn(256, generator=g)
            sd[f"{hp}conv.{3 * i + 1}.weight"] = 0.7 + 0.6 * torch.rand(256, generator=g)
            sd[f"{hp}conv.{3 * i + 1}.bias"] = 0.1 * torch.randn(256, generator=g)
        if hp.startswith("head.classification"):
            if cls_bias is None:
                cls_bias = [-4.0] * (num_classes - 1) + [3.0]
            sd[hp + "cls_logits.weight"] = torch.randn((num_classes, 256, 3, 3), generator=g) * cls_weight_std
            sd[hp + "cls_logits.bias"] = torch.tensor(cls_bias, dtype=torch.float32)
            sd[hp + "hand_lr_layer.weight"] = torch.randn((2, 256, 3, 3), generator=g) * 0.01
            sd[hp + "hand_lr_layer.bias"] = torch.zeros(2)
            if ext:
                sd[hp + "hand_contact_state_layer.weight"] = torch.randn((5, 256, 3, 3), generator=g) * 0.01
                sd[hp + "hand_contact_state_layer.bias"] = torch.zeros(5)
                sd[hp + "hand_dydx_layer.weight"] = torch.randn((3, 256, 3, 3), generator=g) * 0.01
                sd[hp + "hand_dydx_layer.bias"] = 0.05 * torch.ones(3)
        else:
            sd[hp + "bbox_reg.weight"] = torch.randn((4, 256, 3, 3), generator=g) * 0.01
            sd[hp + "bbox_reg.bias"] = torch.full((4,), float(reg_bias))
            sd[hp + "bbox_ctrness.weight"] = torch.randn((1, 256, 3, 3), generator=g) * 0.01
            sd[hp + "bbox_ctrness.bias"] = torch.full((1,), float(ctr_bias))
    return sd


def a2j_anchor_table(shape=(11, 11), stride: int = 16, p=(2, 6, 10, 14)) -> torch.Tensor:
    """Closed form of generate_anchors + shift (a2j/anchor.py:7-42):
    row (w*H + h)*16 + i*4 + j holds (stride*h + p[i], stride*w + p[j])."""
    hh, ww = shape
    pv = np.asarray(p, dtype=np.float32)
    w_i, h_i, a_i, a_j = np.meshgrid(np.arange(ww), np.arange(hh), np.arange(len(p)), np.arange(len(p)), indexing="ij")
    out = np.stack((stride * h_i + pv[a_i], stride * w_i + pv[a_j]), axis=-1).astype(np.float32)
    return torch.from_numpy(out.reshape(-1, 2))


def a2j_state_dict(num_joints: int = 21, seed: int = 1, channel_in: int = 1) -> "OrderedDict[str, torch.Tensor]":
    g = torch.Generator().manual_seed(seed)
    sd = OrderedDict()
    p = "Backbone.model."
    sd[p + "conv1.weight"] = _kaiming(g, 64, 4 if channel_in == 4 else 3, 7)
    _bn(sd, g, p + "bn1", 64, True)
    inpl = 64
    for li, (planes, nblocks) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
        for bi in range(nblocks):
            q = f"{p}layer{li}.{bi}."
            sd[q + "conv1.weight"] = _kaiming(g, planes, inpl, 1)
            _bn(sd, g, q + "bn1", planes, True)
            sd[q + "conv2.weight"] = _kaiming(g, planes, planes, 3)
            _bn(sd, g, q + "bn2", planes, True)
            sd[q + "conv3.weight"] = _kaiming(g, planes * 4, planes, 1)
            _bn(sd, g, q + "bn3", planes * 4, True, gamma=0.4)
            if bi == 0:
                sd[q + "downsample.0.weight"] = _kaiming(g, planes * 4, inpl, 1)
                _bn(sd, g, q + "downsample.1", planes * 4, True)
            inpl = planes * 4
    sd[p + "fc.weight"] = torch.randn((1000, 2048), generator=g) * 0.01     # present in the state dict, never run
    sd[p + "fc.bias"] = torch.zeros(1000)
    for name, cin, cout in (("regressionModel.", 2048, 16 * num_joints * 2),
                            ("classificationModel.", 1024, 16 * num_joints),
                            ("DepthRegressionModel.", 2048, 16 * num_joints)):
        c = cin
        for i in range(1, 5):
            sd[f"{name}conv{i}.weight"] = _xavier(g, 256, c, 3)
            sd[f"{name}conv{i}.bias"] = 0.02 * torch.randn(256, generator=g)
            _bn(sd, g, f"{name}bn{i}", 256, True)
            c = 256
        sd[name + "output.weight"] = _xavier(g, cout, 256, 3)
        sd[name + "output.bias"] = 0.02 * torch.randn(cout, generator=g)
    anchors = a2j_anchor_table()
    sd["criterion.all_anchors"] = anchors.clone()
    sd["criterion.thres"] = torch.tensor([16.0, 32.0])
    sd["post_process.all_anchors"] = anchors.clone()
    sd["post_process.thres"] = torch.tensor(8.0)
    return sd


def stress_head_tensors(seed: int, batch: int, locs: int, num_classes: int = 3, mu: float = -0.35):
    """Synthetic FCOS head outputs for BASELINE.json config 4: the 0.7 score cut is hard-coded in the
    reference (fcos_utils/fcos.py:600), so ~10 000 candidates per frame are reached by RAISING the logits: cls ~ N(mu, 1)
    with mu chosen so that ~56 % of sqrt(sigmoid(cls) * sigmoid(ctr)) exceed 0.7, ctr ~ N(2, 1), reg ~ U(0.5, 4)."""
    g = torch.Generator().manual_seed(seed)
    return {
        "cls_logits": torch.randn(batch, locs, num_classes, generator=g) + mu,
        "bbox_ctrness": torch.randn(batch, locs, 1, generator=g) + 2.0,
        "bbox_regression": 0.5 + 3.5 * torch.rand(batch, locs, 4, generator=g),
        "hand_lr": torch.randn(batch, locs, 2, generator=g),
    }


def fill_state_dict(shapes, seed: int = 0):
    """Deterministic synthetic weights for a list of (key, shape) pairs (pose2mesh: no checkpoint is available offline).
    Linear weights ~ N(0, 1 / fan_in), biases ~ N(0, 0.05); BatchNorm weight ~ U(0.6, 1.4), bias ~ N(0, 0.1), running_mean ~
    N(0, 0.2), running_var ~ U(0.5, 1.5).  Generated on the CPU in key order, so the same list gives the same tensors anywhere."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for key, shape in shapes:
        shape = tuple(shape)
        leaf = key.rsplit(".", 1)[-1]
        if leaf == "num_batches_tracked":
            sd[key] = torch.zeros((), dtype=torch.int64)
        elif leaf == "running_var":
            sd[key] = 0.5 + torch.rand(shape, generator=g)
        elif leaf == "running_mean":
            sd[key] = 0.2 * torch.randn(shape, generator=g)
        elif len(shape) == 2:
            sd[key] = torch.randn(shape, generator=g) * (1.0 / shape[1]) ** 0.5
        elif ".bn." in key or "batch_norm" in key:
            sd[key] = (0.6 + 0.8 * torch.rand(shape, generator=g)) if leaf == "weight" else 0.1 * torch.randn(shape, generator=g)
        else:
            sd[key] = 0.05 * torch.randn(shape, generator=g)
    return sd

"""Oracle for the FCOS hand detector.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

A functional torch-CPU restatement driven by a state dict with the reference's parameter
names (fcos_utils/fcos.py:455-514 builds the modules; torchvision supplies the backbone,
FPN, transform and NMS).  ``emulate_bf16=True`` rounds weights and inter-layer activations
to bfloat16 at exactly the points where the CUDA path stores bf16, so that kernels can be
gated at a tight tolerance; ``emulate_bf16=False`` is the fp32 reference arithmetic.

Pinned against the real reference by oracle/make_golden.py -> tests/golden/*.pt.
"""
from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from . import nms_oracle

IMAGE_MEAN = (0.485, 0.456, 0.406)   # fcos_utils/fcos.py:501-504
IMAGE_STD = (0.229, 0.224, 0.225)
SCORE_CUT = 0.7                      # fcos_utils/fcos.py:600 (hard-coded)
NMS_IOU = 0.3                        # fcos_utils/fcos.py:635 (hard-coded)
BN_EPS = 1e-5                        # torchvision/ops/misc.py FrozenBatchNorm2d default
GN_EPS = 1e-5                        # nn.GroupNorm default, fcos_utils/fcos.py:232


def _q(x: torch.Tensor, emulate: bool) -> torch.Tensor:
    return x.to(torch.bfloat16).to(torch.float32) if emulate else x


# ----------------------------------------------------------------------------- T1
def resized_size(h: int, w: int, min_size: int, max_size: int) -> Tuple[int, int]:
    """torchvision/models/detection/transform.py:25-72 (eager branch): python-double scale,
    output = floor(in * scale) as F.interpolate(recompute_scale_factor=True) does."""
    scale = min(min_size / min(h, w), max_size / max(h, w))
    return int(math.floor(float(h) * scale)), int(math.floor(float(w) * scale))


def transform(images: Sequence[torch.Tensor], min_size: int = 800, max_size: int = 1333,
              size_divisible: int = 32) -> Tuple[torch.Tensor, List[Tuple[int, int]]]:
    """normalize -> bilinear resize -> zero-padded batch (fcos_utils/fcos.py:709;
    torchvision transform.py:119-158, 160-169, 237-255).  Returns NCHW fp32 canvas."""
    mean = torch.tensor(IMAGE_MEAN, dtype=torch.float32)[:, None, None]
    std = torch.tensor(IMAGE_STD, dtype=torch.float32)[:, None, None]
    out, sizes = [], []
    for img in images:
        x = (img.to(torch.float32) - mean) / std
        h, w = x.shape[-2:]
        oh, ow = resized_size(h, w, min_size, max_size)
        x = F.interpolate(x[None], size=(oh, ow), mode="bilinear", align_corners=False)[0]
        out.append(x)
        sizes.append((oh, ow))
    ch = int(math.ceil(max(s[0] for s in sizes) / size_divisible) * size_divisible)
    cw = int(math.ceil(max(s[1] for s in sizes) / size_divisible) * size_divisible)
    canvas = torch.zeros((len(out), 3, ch, cw), dtype=torch.float32)
    for i, x in enumerate(out):
        canvas[i, :, : x.shape[1], : x.shape[2]] = x
    return canvas, sizes


# ------------------------------------------------------------------------- B1 / B2
def _bn_affine(sd: Dict[str, torch.Tensor], prefix: str) -> Tuple[torch.Tensor, torch.Tensor]:
    """FrozenBatchNorm2d as scale/shift (torchvision/ops/misc.py:54-63)."""
    w, b = sd[prefix + ".weight"].float(), sd[prefix + ".bias"].float()
    rm, rv = sd[prefix + ".running_mean"].float(), sd[prefix + ".running_var"].float()
    scale = w * (rv + BN_EPS).rsqrt()
    return scale, b - rm * scale


def _conv_affine(x, w, scale, shift, stride, pad, emulate, relu, residual=None, dilation=1):
    """conv (fp32 accumulate) -> *scale + shift (+residual) -> relu -> (bf16 round)."""
    if scale is None:     # plain conv + bias, evaluated the way nn.Conv2d does
        y = F.conv2d(x, _q(w.float(), emulate), shift, stride=stride, padding=pad, dilation=dilation)
    else:
        y = F.conv2d(x, _q(w.float(), emulate), None, stride=stride, padding=pad, dilation=dilation)
        y = y * scale[None, :, None, None] + shift[None, :, None, None]
    if residual is not None:
        y = y + residual
    if relu:
        y = F.relu(y)
    return _q(y, emulate)


def resnet34_body(sd, x, emulate: bool, prefix: str = "backbone.body."):
    """torchvision resnet34 body with FrozenBN, returns (C3, C4, C5)
    (fcos_utils/fcos.py:476; torchvision backbone_utils.py:62-118)."""
    s, b = _bn_affine(sd, prefix + "bn1")
    x = _conv_affine(x, sd[prefix + "conv1.weight"], s, b, 2, 3, emulate, relu=True)
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    feats = []
    for li, nblocks in enumerate((3, 4, 6, 3), start=1):
        for bi in range(nblocks):
            p = f"{prefix}layer{li}.{bi}."
            stride = 2 if (li > 1 and bi == 0) else 1
            identity = x
            if (p + "downsample.0.weight") in sd:
                ds, db = _bn_affine(sd, p + "downsample.1")
                identity = _conv_affine(x, sd[p + "downsample.0.weight"], ds, db, stride, 0, emulate, relu=False)
            s1, b1 = _bn_affine(sd, p + "bn1")
            y = _conv_affine(x, sd[p + "conv1.weight"], s1, b1, stride, 1, emulate, relu=True)
            s2, b2 = _bn_affine(sd, p + "bn2")
            x = _conv_affine(y, sd[p + "conv2.weight"], s2, b2, 1, 1, emulate, relu=True, residual=identity)
        if li >= 2:
            feats.append(x)
    return feats


def _fpn_key(sd, base: str) -> str:
    """torchvision >= 0.13 wraps FPN convs in Conv2dNormActivation ('.0'); 0.11.3 checkpoints
    have no '.0' (torchvision/ops/feature_pyramid_network.py:112-142 converts on load)."""
    return base + ".0" if (base + ".0.weight") in sd else base


def fpn(sd, feats, emulate: bool, prefix: str = "backbone.fpn."):
    """torchvision/ops/feature_pyramid_network.py:172-204; the 'pool' level is discarded
    by fcos_utils/fcos.py:741-742 so it is not computed."""
    ones = None
    n = len(feats)
    inner = [None] * n
    outs = [None] * n
    for i in range(n - 1, -1, -1):
        k = _fpn_key(sd, f"{prefix}inner_blocks.{i}")
        td = None
        if i < n - 1:
            td = F.interpolate(inner[i + 1], size=feats[i].shape[-2:], mode="nearest")
        inner[i] = _conv_affine(feats[i], sd[k + ".weight"], ones, sd[k + ".bias"].float(), 1, 0, emulate,
                                relu=False, residual=td)
    for i in range(n):
        k = _fpn_key(sd, f"{prefix}layer_blocks.{i}")
        outs[i] = _conv_affine(inner[i], sd[k + ".weight"], ones, sd[k + ".bias"].float(), 1, 1, emulate, relu=False)
    return outs


# ------------------------------------------------------------------------- H1 / H2
def _tower(sd, x, prefix: str, emulate: bool):
    """4 x (conv3x3 + bias -> GroupNorm(32) -> ReLU)  (fcos_utils/fcos.py:232-240, 352-360)."""
    ones = None
    for i in range(4):
        raw = _conv_affine(x, sd[f"{prefix}conv.{3 * i}.weight"], ones, sd[f"{prefix}conv.{3 * i}.bias"].float(),
                           1, 1, emulate, relu=False)
        y = F.group_norm(raw, 32, sd[f"{prefix}conv.{3 * i + 1}.weight"].float(),
                         sd[f"{prefix}conv.{3 * i + 1}.bias"].float(), eps=GN_EPS)
        x = _q(F.relu(y), emulate)
    return x


def _out_conv(sd, x, name: str, emulate: bool):
    return F.conv2d(x, _q(sd[name + ".weight"].float(), emulate), sd[name + ".bias"].float(), padding=1)


def _to_nhwa_k(t: torch.Tensor, k: int) -> torch.Tensor:
    """(N, A*K, H, W) -> (N, H*W*A, K) with A = 1 (fcos_utils/fcos.py:282-285)."""
    n, _, h, w = t.shape
    return t.view(n, -1, k, h, w).permute(0, 3, 4, 1, 2).reshape(n, -1, k)


def head(sd, feats, num_classes: int, ext: bool, emulate: bool, prefix: str = "head."):
    """FCOSHead.forward (fcos_utils/fcos.py:180-200, 267-329, 373-395)."""
    out = {k: [] for k in ("cls_logits", "hand_lr", "bbox_regression", "bbox_ctrness")}
    if ext:
        out["hand_contact_state"] = []
        out["hand_dxdy"] = []
    cp, rp = prefix + "classification_head.", prefix + "regression_head."
    for f in feats:
        ct = _tower(sd, f, cp, emulate)
        out["cls_logits"].append(_to_nhwa_k(_out_conv(sd, ct, cp + "cls_logits", emulate), num_classes))
        out["hand_lr"].append(_to_nhwa_k(_out_conv(sd, ct, cp + "hand_lr_layer", emulate), 2))
        if ext:
            d = F.relu(_out_conv(sd, ct, cp + "hand_dydx_layer", emulate))
            d = torch.cat([d[:, 0:1], 0.1 * F.normalize(d[:, 1:], p=2, dim=1)], dim=1)   # fcos.py:299-303
            out["hand_dxdy"].append(_to_nhwa_k(d, 3))
            out["hand_contact_state"].append(
                _to_nhwa_k(_out_conv(sd, ct, cp + "hand_contact_state_layer", emulate), 5))
        rt = _tower(sd, f, rp, emulate)
        out["bbox_regression"].append(_to_nhwa_k(F.relu(_out_conv(sd, rt, rp + "bbox_reg", emulate)), 4))
        out["bbox_ctrness"].append(_to_nhwa_k(_out_conv(sd, rt, rp + "bbox_ctrness", emulate), 1))
    return {k: torch.cat(v, dim=1) for k, v in out.items()}


# ------------------------------------------------------------------------------ A1
def anchors_for(canvas_hw: Tuple[int, int], grids: Sequence[Tuple[int, int]],
                sizes: Sequence[int] = (8, 16, 32)) -> torch.Tensor:
    """AnchorGenerator with one square anchor per cell (fcos_utils/anchor_utils.py:56-132;
    sizes from fcos_utils/fcos.py:489-491).  stride = canvas // grid; base anchor
    round([-s,-s,s,s]/2); level-major, then row-major over (y, x)."""
    per_level = []
    for (gh, gw), s in zip(grids, sizes):
        sh, sw = canvas_hw[0] // gh, canvas_hw[1] // gw
        base = torch.tensor([-s, -s, s, s], dtype=torch.float32).div(2).round()
        ys = torch.arange(gh, dtype=torch.int32) * sh
        xs = torch.arange(gw, dtype=torch.int32) * sw
        yy, xx = torch.meshgrid(ys, xs, indexing="ij")
        shifts = torch.stack((xx.reshape(-1), yy.reshape(-1), xx.reshape(-1), yy.reshape(-1)), dim=1)
        per_level.append((shifts.view(-1, 1, 4) + base.view(1, -1, 4)).reshape(-1, 4))
    return torch.cat(per_level)


# ------------------------------------------------------------------------- P1 .. P6
def decode_boxes(rel: torch.Tensor, anchors: torch.Tensor) -> torch.Tensor:
    """BoxLinearCoder.decode_single, normalize_by_size=True (fcos_utils/det_utils.py:266-294)."""
    a = anchors.to(rel.dtype)
    cx = 0.5 * (a[:, 0] + a[:, 2])
    cy = 0.5 * (a[:, 1] + a[:, 3])
    bw = a[:, 2] - a[:, 0]
    bh = a[:, 3] - a[:, 1]
    rel = rel * torch.stack((bw, bh, bw, bh), dim=1)
    return torch.stack((cx - rel[:, 0], cy - rel[:, 1], cx + rel[:, 2], cy + rel[:, 3]), dim=1)


def score_and_select(head_out: Dict[str, torch.Tensor]):
    """fcos_utils/fcos.py:598-603: sqrt(sigmoid(cls) * sigmoid(ctr)), max over classes
    (class 0 included), strict > 0.7 cut, side = argmax sigmoid(hand_lr)."""
    scores = torch.sqrt(torch.sigmoid(head_out["cls_logits"]) * torch.sigmoid(head_out["bbox_ctrness"]))
    scores_max, labels_max = torch.max(scores, dim=-1)
    masks = scores_max > SCORE_CUT
    _, sides_max = torch.max(torch.sigmoid(head_out["hand_lr"]), dim=-1)
    return scores_max, labels_max, masks, sides_max


def level_index(num_per_level: Sequence[int]) -> torch.Tensor:
    """fcos_utils/fcos.py:610-618: pyramid level of every location, as float32."""
    idx = torch.zeros(sum(num_per_level))
    start = 0
    for lvl, n in enumerate(num_per_level):
        idx[start:start + n] = lvl
        start += n
    return idx


def resize_boxes(boxes: torch.Tensor, from_hw, to_hw) -> torch.Tensor:
    """fcos_utils/fcos.py:770-783: float32 ratio tensors new/orig, per-axis multiply."""
    rh = torch.tensor(to_hw[0], dtype=torch.float32) / torch.tensor(from_hw[0], dtype=torch.float32)
    rw = torch.tensor(to_hw[1], dtype=torch.float32) / torch.tensor(from_hw[1], dtype=torch.float32)
    x1, y1, x2, y2 = boxes.unbind(1)
    return torch.stack((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=1)


def postprocess(head_out: Dict[str, torch.Tensor], anchors: torch.Tensor, num_per_level: Sequence[int],
                image_sizes, original_sizes, ext: bool = False) -> List[Dict[str, torch.Tensor]]:
    """postprocess_detections + postprocess (fcos_utils/fcos.py:572-669)."""
    scores_max, labels_max, masks, sides_max = score_and_select(head_out)
    fidx = level_index(num_per_level)
    if ext:
        _, contact_max = torch.max(torch.sigmoid(head_out["hand_contact_state"]), dim=-1)
    dets = []
    for i in range(scores_max.shape[0]):
        m = masks[i]
        boxes = decode_boxes(head_out["bbox_regression"][i], anchors)[m]
        scores, labels, sides = scores_max[i][m], labels_max[i][m], sides_max[i][m]
        keep = torch.from_numpy(nms_oracle.batched_nms(boxes.numpy(), scores.numpy(), labels.numpy(), NMS_IOU))
        d = {
            "boxes": resize_boxes(boxes[keep], image_sizes[i], original_sizes[i]),
            "scores": scores[keep],
            "labels": labels[keep],
            "sides": sides[keep].reshape(-1),
        }
        if ext:
            d["dxdymags"] = head_out["hand_dxdy"][i][m][keep]
            d["contacts"] = contact_max[i][m][keep].reshape(-1)
        else:
            d["feature_idx"] = fidx[m][keep].reshape(-1)
        d["_candidate_index"] = torch.nonzero(m).reshape(-1)[keep]      # oracle-only bookkeeping
        dets.append(d)
    return dets


# ---------------------------------------------------------------------- whole model
def fcos_forward(sd: Dict[str, torch.Tensor], images: Sequence[torch.Tensor], num_classes: int,
                 ext: bool = False, min_size: int = 800, max_size: int = 1333, emulate_bf16: bool = False,
                 return_taps: bool = False):
    """FCOS.forward in eval mode (fcos_utils/fcos.py:675-767)."""
    original_sizes = [tuple(int(v) for v in img.shape[-2:]) for img in images]
    canvas, image_sizes = transform(images, min_size, max_size)
    x = _q(canvas, emulate_bf16)
    c = resnet34_body(sd, x, emulate_bf16)
    p = fpn(sd, c, emulate_bf16)
    ho = head(sd, p, num_classes, ext, emulate_bf16)
    grids = [tuple(t.shape[-2:]) for t in p]
    anchors = anchors_for(tuple(canvas.shape[-2:]), grids)
    npl = [g[0] * g[1] for g in grids]
    dets = postprocess(ho, anchors, npl, image_sizes, original_sizes, ext)
    if return_taps:
        return dets, {"canvas": canvas, "c": c, "p": p, "head": ho, "anchors": anchors, "num_per_level": npl,
                      "image_sizes": image_sizes, "original_sizes": original_sizes}
    return dets

"""FCOS hand detector on B200 kernels -- drop-in for the reference's ``fcos_utils.fcos``.

Reference: fcos_utils/fcos.py (FCOS :399-767, FCOSHead :22-200, heads :203-395, resize_boxes :770-783,
psum :786-790).  Same constructor signature, attribute names, state-dict keys (232 for ``ext=False``) and
eval-mode output (a list of dicts per image), so ``HandNet``, ``ros_demo.py`` and
``trainval_net_fcos.py --test-only`` can use it unchanged.  What differs is how ``forward`` is computed:
nothing runs through torch/cuDNN -- the modules below only OWN the parameters, and ``hn_b200.runtime``
enqueues hand-written sm_100a kernels (tcgen05 implicit-GEMM convolutions in bf16, fused post-processing).

Training (``compute_loss``, matchers, GIoU loss) is out of scope of this build: calling the model in
training mode raises.  There is no CPU path: inputs must be CUDA tensors.
"""
from __future__ import annotations

import math
from functools import partial
from typing import Callable, Dict, List, Optional, Tuple

import torch
from torch import Tensor, nn

from hn_b200 import runtime

from . import det_utils
from .anchor_utils import AnchorGenerator


# --------------------------------------------------------------------------------------------------
# parameter containers (never called; the kernels read their tensors)
# --------------------------------------------------------------------------------------------------
class FrozenBatchNorm2d(nn.Module):
    """Buffers of torchvision.ops.misc.FrozenBatchNorm2d (weight, bias, running_mean, running_var)."""

    def __init__(self, num_features: int, eps: float = 1e-5):
        super().__init__()
        self.eps = eps
        self.register_buffer("weight", torch.ones(num_features))
        self.register_buffer("bias", torch.zeros(num_features))
        self.register_buffer("running_mean", torch.zeros(num_features))
        self.register_buffer("running_var", torch.ones(num_features))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        state_dict.pop(prefix + "num_batches_tracked", None)      # as torchvision does
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


def _conv(cin, cout, k, stride=1, bias=False):
    return nn.Conv2d(cin, cout, kernel_size=k, stride=stride, padding=k // 2, bias=bias)


class _BasicBlock(nn.Module):
    def __init__(self, inplanes, planes, stride):
        super().__init__()
        self.conv1 = _conv(inplanes, planes, 3, stride)
        self.bn1 = FrozenBatchNorm2d(planes)
        self.conv2 = _conv(planes, planes, 3)
        self.bn2 = FrozenBatchNorm2d(planes)
        self.downsample = None
        if stride != 1 or inplanes != planes:
            self.downsample = nn.Sequential(_conv(inplanes, planes, 1, stride), FrozenBatchNorm2d(planes))


class _ResNet34Body(nn.Module):
    """torchvision resnet34 trunk (conv1 .. layer4) as used by resnet_fpn_backbone (fcos.py:476)."""

    def __init__(self):
        super().__init__()
        self.conv1 = _conv(3, 64, 7, 2)
        self.bn1 = FrozenBatchNorm2d(64)
        inpl = 64
        for li, (planes, n) in enumerate(((64, 3), (128, 4), (256, 6), (512, 3)), start=1):
            blocks = []
            for bi in range(n):
                blocks.append(_BasicBlock(inpl, planes, 2 if (li > 1 and bi == 0) else 1))
                inpl = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")


class _FPN(nn.Module):
    """Parameters of torchvision's FeaturePyramidNetwork (keys ``inner_blocks.i.0.*`` / ``layer_blocks.i.0.*``)."""

    def __init__(self, in_channels: List[int], out_channels: int):
        super().__init__()
        self.inner_blocks = nn.ModuleList(nn.Sequential(_conv(c, out_channels, 1, bias=True)) for c in in_channels)
        self.layer_blocks = nn.ModuleList(nn.Sequential(_conv(out_channels, out_channels, 3, bias=True)) for _ in in_channels)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_uniform_(m.weight, a=1)
                nn.init.constant_(m.bias, 0)

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        # checkpoints written with torchvision 0.11.3 (the reference's pin) have no ".0"
        for block in ("inner_blocks", "layer_blocks"):
            for i in range(len(self.inner_blocks)):
                for t in ("weight", "bias"):
                    old, new = f"{prefix}{block}.{i}.{t}", f"{prefix}{block}.{i}.0.{t}"
                    if old in state_dict:
                        state_dict[new] = state_dict.pop(old)
        super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)


class _BackboneWithFPN(nn.Module):
    def __init__(self):
        super().__init__()
        self.body = _ResNet34Body()
        self.fpn = _FPN([128, 256, 512], 256)
        self.out_channels = 256


def _tower(in_channels, num_convs, norm_layer):
    layers = []
    for _ in range(num_convs):
        layers += [_conv(in_channels, in_channels, 3, bias=True), norm_layer(in_channels), nn.ReLU()]
    seq = nn.Sequential(*layers)
    for layer in seq.children():
        if isinstance(layer, nn.Conv2d):
            nn.init.normal_(layer.weight, std=0.01)
            nn.init.constant_(layer.bias, 0)
    return seq


class FCOSClassificationHead(nn.Module):
    """Parameters of the classification tower and its output convs (reference fcos.py:203-264)."""

    def __init__(self, in_channels: int, num_anchors: int, num_classes: int, num_convs: int = 4,
                 prior_probability: float = 0.01, norm_layer: Optional[Callable[..., nn.Module]] = None,
                 ext: bool = True) -> None:
        super().__init__()
        self.num_classes, self.num_anchors, self.ext = num_classes, num_anchors, ext
        norm_layer = norm_layer or partial(nn.GroupNorm, 32)
        self.conv = _tower(in_channels, num_convs, norm_layer)
        self.cls_logits = _conv(in_channels, num_anchors * num_classes, 3, bias=True)
        self.hand_lr_layer = _conv(in_channels, num_anchors * 2, 3, bias=True)
        outs = [self.cls_logits, self.hand_lr_layer]
        if ext:
            self.hand_contact_state_layer = _conv(in_channels, num_anchors * 5, 3, bias=True)
            self.hand_dydx_layer = _conv(in_channels, num_anchors * 3, 3, bias=True)
            outs += [self.hand_contact_state_layer, self.hand_dydx_layer]
        for m in outs:
            nn.init.normal_(m.weight, std=0.01)
            nn.init.zeros_(m.bias)
        nn.init.constant_(self.cls_logits.bias, -math.log((1 - prior_probability) / prior_probability))


class FCOSRegressionHead(nn.Module):
    """Parameters of the box tower, ``bbox_reg`` and ``bbox_ctrness`` (reference fcos.py:332-371)."""

    def __init__(self, in_channels: int, num_anchors: int, num_convs: int = 4,
                 norm_layer: Optional[Callable[..., nn.Module]] = None):
        super().__init__()
        norm_layer = norm_layer or partial(nn.GroupNorm, 32)
        self.conv = _tower(in_channels, num_convs, norm_layer)
        self.bbox_reg = _conv(in_channels, num_anchors * 4, 3, bias=True)
        self.bbox_ctrness = _conv(in_channels, num_anchors * 1, 3, bias=True)
        for m in (self.bbox_reg, self.bbox_ctrness):
            nn.init.normal_(m.weight, std=0.01)
            nn.init.zeros_(m.bias)


class FCOSHead(nn.Module):
    def __init__(self, in_channels: int, num_anchors: int, num_classes: int, num_convs: Optional[int] = 4,
                 ext: bool = True) -> None:
        super().__init__()
        self.ext = ext
        self.box_coder = det_utils.BoxLinearCoder(normalize_by_size=True)
        self.classification_head = FCOSClassificationHead(in_channels, num_anchors, num_classes, num_convs, ext=ext)
        self.regression_head = FCOSRegressionHead(in_channels, num_anchors, num_convs)


# --------------------------------------------------------------------------------------------------
class FCOS(runtime.WeightsEpochMixin, nn.Module):
    """FCOS(num_classes, ext=True, min_size=800, max_size=1333, ...) -- reference fcos.py:455-514.

    ``forward(images: List[Tensor[3,H,W]], targets=None) -> List[Dict[str, Tensor]]`` with keys ``boxes``,
    ``scores``, ``labels``, ``sides`` and ``feature_idx`` (``ext=False``) or ``dxdymags`` / ``contacts``
    (``ext=True``); detections are score-descending.  As in the reference the live post-processing constants
    are the hard-coded 0.7 score cut and 0.3 NMS IoU (fcos.py:600,635); ``score_thresh``, ``nms_thresh``,
    ``detections_per_img`` and ``topk_candidates`` are stored and unused."""

    def __init__(self, num_classes: int, ext: bool = True, min_size: int = 800, max_size: int = 1333,
                 image_mean: Optional[List[float]] = None, image_std: Optional[List[float]] = None,
                 anchor_generator: Optional[AnchorGenerator] = None, head: Optional[nn.Module] = None,
                 center_sampling_radius: float = 1.5, score_thresh: float = 0.2, nms_thresh: float = 0.6,
                 detections_per_img: int = 100, topk_candidates: int = 1000):
        super().__init__()
        self._install_weight_hooks()
        self.ext = ext
        self.backbone = _BackboneWithFPN()
        if anchor_generator is None:
            anchor_generator = AnchorGenerator(((8,), (16,), (32,)), ((1.0,),) * 3)
        assert isinstance(anchor_generator, AnchorGenerator)
        self.anchor_generator = anchor_generator
        assert self.anchor_generator.num_anchors_per_location()[0] == 1
        self.anchor_sizes = tuple(int(s[0]) for s in anchor_generator.sizes)
        if head is None:
            head = FCOSHead(self.backbone.out_channels, 1, num_classes, ext=ext)
        self.head = head
        self.box_coder = det_utils.BoxLinearCoder(normalize_by_size=True)
        self.image_mean = list(image_mean) if image_mean is not None else [0.485, 0.456, 0.406]
        self.image_std = list(image_std) if image_std is not None else [0.229, 0.224, 0.225]
        self.min_size, self.max_size = min_size, max_size
        self.center_sampling_radius = center_sampling_radius
        self.score_thresh, self.nms_thresh = score_thresh, nms_thresh            # stored, never read (as reference)
        self.detections_per_img, self.topk_candidates = detections_per_img, topk_candidates
        # the constants the reference actually uses
        self.score_cut = 0.7                  # fcos.py:600
        self.nms_iou = 0.3                    # fcos.py:635
        self.nms_coord_trick_numel = 4000     # torchvision/ops/boxes.py:80, CPU branch (the oracle's path)
        self._executor = runtime.FCOSExecutor(self)

    # -- device results ---------------------------------------------------------------------------
    def forward_device(self, images: List[Tensor]) -> Dict[str, Tensor]:
        """Dense detections on the device: boxes [B,L,4], scores/labels/sides/level [B,L], keep_count [B].
        No host synchronisation; used by HandNet to chain the crop and pose kernels."""
        if self.training:
            raise NotImplementedError("training mode (compute_loss) is outside the scope of this build")
        if len(images) == 0:
            raise ValueError("empty image list")
        for im in images:
            if not im.is_cuda:
                raise RuntimeError("FCOS (B200 build) needs CUDA tensors: there is no CPU fallback")
        return self._executor.forward_device(images)

    def head_outputs(self, images: List[Tensor]) -> Dict[str, Tensor]:
        """Raw head tensors [B, HWA, K] as FCOSHead.forward returns them (fcos.py:180-200)."""
        out = self.forward_device(images)
        v = self._executor.head_views(out["plan"])
        if "hand_dxdy_relu" in v:
            d = v.pop("hand_dxdy_relu")
            v["hand_dxdy"] = torch.cat([d[..., :1], 0.1 * torch.nn.functional.normalize(d[..., 1:], p=2, dim=-1)], -1)
        return v

    @staticmethod
    def split_detections(out: Dict[str, Tensor], ext: bool) -> List[Dict[str, Tensor]]:
        counts = out["keep_count"].tolist()                     # the one host sync of the detector
        dets = []
        for i, k in enumerate(counts):
            d = {"boxes": out["boxes"][i, :k], "scores": out["scores"][i, :k], "labels": out["labels"][i, :k]}
            if ext:
                d["dxdymags"] = out["dxdymags"][i, :k]
                d["contacts"] = out["contacts"][i, :k]
                d["sides"] = out["sides"][i, :k]
            else:
                d["sides"] = out["sides"][i, :k]
                d["feature_idx"] = out["level"][i, :k]
            dets.append(d)
        return dets

    def forward(self, images: List[Tensor], targets: Optional[List[Dict[str, Tensor]]] = None):
        out = self.forward_device(images)
        return self.split_detections(out, self.ext)


def resize_boxes(boxes: Tensor, original_size: List[int], new_size: List[int]) -> Tensor:
    """Scale boxes from ``original_size`` to ``new_size`` with float32 ratios (reference fcos.py:770-783)."""
    rh = torch.tensor(new_size[0], dtype=torch.float32, device=boxes.device) / \
        torch.tensor(original_size[0], dtype=torch.float32, device=boxes.device)
    rw = torch.tensor(new_size[1], dtype=torch.float32, device=boxes.device) / \
        torch.tensor(original_size[1], dtype=torch.float32, device=boxes.device)
    x1, y1, x2, y2 = boxes.unbind(1)
    return torch.stack((x1 * rw, y1 * rh, x2 * rw, y2 * rh), dim=1)


def psum(a):
    """Exclusive-then-inclusive prefix sums: [0, a0, a0+a1, ...] (reference fcos.py:786-790)."""
    out = [0]
    for v in a:
        out.append(out[-1] + v)
    return out

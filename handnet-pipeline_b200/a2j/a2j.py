"""A2J anchor-to-joint pose network on B200 kernels -- drop-in for the reference's ``a2j.a2j``.

Reference: a2j/a2j.py (convert_joints :17-43, towers :44-181, ResNetBackBone :184-210, A2JModel :212-250).
Constructor signatures, sub-module names and the 414 state-dict keys are the reference's; ``forward`` returns
``[n, 21, 3]`` float32 on the CPU exactly like ``A2JModel.eager_outputs`` (a2j.py:226-229).  The convolutions
run as tcgen05 implicit GEMMs in bf16 with BatchNorm + ReLU folded into the epilogue, and the anchor
post-process as one fused kernel (see hn_b200.runtime.A2JExecutor).

Training pieces of the reference file (A2J_loss use, A2JModelLightning steps, A2JDataModule) are out of scope;
thin stand-ins keep ``from a2j.a2j import A2JModelLightning`` importable.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn

from hn_b200 import runtime

from . import resnet
from .anchor import A2J_loss, post_process


def uvd2xyz(pts, paras, flipy=1):
    """Pinhole back-projection, paras = (fx, fy, cx, cy)  (reference datasets3d/a2jdataset.py:31-38)."""
    out = np.array(pts, copy=True).reshape(-1, 3)
    out[:, :2] = (out[:, :2] - paras[2:]) * out[:, 2:] / paras[:2]
    out[:, 1] *= flipy
    return out.reshape(np.shape(pts)).astype(np.float32)


def convert_joints(jt_uvd_pred, jt_uvd_gt, box, paras, cropWidth, cropHeight):
    """Crop-space (u, v, d) -> image pixels, optionally -> camera-space millimetres (reference a2j.py:17-43).
    Host-side numpy, as the callers pass numpy arrays (ros_demo.py:289,329-330)."""
    def to_image(j):
        j = j.reshape(-1, 3)
        out = np.ones_like(j)
        out[:, 0] = j[:, 0] * (X_max - X_min) / cropWidth + X_min
        out[:, 1] = j[:, 1] * (Y_max - Y_min) / cropHeight + Y_min
        out[:, 2] = j[:, 2]
        if paras_ is not None:
            out = uvd2xyz(out, paras_) * 1000.0
        return out

    box = box.reshape(4)
    paras_ = None if paras is None else paras.reshape(4)
    X_min, Y_min, X_max, Y_max = box[0], box[1], box[2], box[3]
    pred = to_image(jt_uvd_pred)
    if jt_uvd_gt is not None:
        return pred, to_image(jt_uvd_gt)
    return pred


class _Tower(nn.Module):
    """conv1..4 (3x3 + bias) with bn1..4 and ``output`` conv: the shared shape of the three A2J heads."""

    def __init__(self, num_features_in, out_channels, feature_size=256):
        super().__init__()
        c = num_features_in
        for i in range(1, 5):
            setattr(self, f"conv{i}", nn.Conv2d(c, feature_size, kernel_size=3, padding=1))
            setattr(self, f"bn{i}", nn.BatchNorm2d(feature_size))
            setattr(self, f"act{i}", nn.ReLU())
            c = feature_size
        self.output = nn.Conv2d(feature_size, out_channels, kernel_size=3, padding=1)
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_normal_(m.weight.data)


class DepthRegressionModel(_Tower):
    def __init__(self, num_features_in, num_anchors=16, num_classes=15, feature_size=256):
        super().__init__(num_features_in, num_anchors * num_classes, feature_size)
        self.num_anchors, self.num_classes = num_anchors, num_classes


class RegressionModel(_Tower):
    def __init__(self, num_features_in, num_anchors=16, num_classes=15, feature_size=256):
        super().__init__(num_features_in, num_anchors * num_classes * 2, feature_size)
        self.num_anchors, self.num_classes = num_anchors, num_classes


class ClassificationModel(_Tower):
    def __init__(self, num_features_in, num_anchors=16, num_classes=15, prior=0.01, feature_size=256):
        super().__init__(num_features_in, num_anchors * num_classes, feature_size)
        self.num_anchors, self.num_classes = num_anchors, num_classes


class ResNetBackBone(nn.Module):
    def __init__(self, channel_in):
        super().__init__()
        self.model = resnet.resnet50(pretrained=False)
        self.channel_in = channel_in
        if channel_in == 4:
            self.model.conv1 = nn.Conv2d(4, 64, kernel_size=7, stride=2, padding=3, bias=False)


class A2JModel(runtime.WeightsEpochMixin, nn.Module):
    """A2JModel(num_classes, crop_height, crop_width, is_3D=True, is_RGBD=False, spatial_factor=0.5)."""

    def __init__(self, num_classes, crop_height, crop_width, is_3D=True, is_RGBD=False, spatial_factor=0.5):
        super().__init__()
        self._install_weight_hooks()
        self.is_3D = is_3D
        self.num_joints = num_classes
        self.Backbone = ResNetBackBone(channel_in=4 if is_RGBD else 1)
        self.regressionModel = RegressionModel(2048, num_classes=num_classes)
        self.classificationModel = ClassificationModel(1024, num_classes=num_classes)
        if is_3D:
            self.DepthRegressionModel = DepthRegressionModel(2048, num_classes=num_classes)
        self.criterion = A2J_loss(shape=[crop_height // 16, crop_width // 16], thres=[16.0, 32.0], stride=16,
                                  spatialFactor=spatial_factor, img_shape=[crop_height, crop_width], P_h=None, P_w=None)
        self.post_process = post_process(shape=[crop_height // 16, crop_width // 16], stride=16, P_h=None, P_w=None)
        self.reg_loss_factor = 3
        self._executor = runtime.A2JExecutor(self)

    def forward_device(self, x: torch.Tensor) -> torch.Tensor:
        """[n, 1|4, H, W] float32 on the device -> [n, joints, 3] on the device, no host sync."""
        if not x.is_cuda:
            raise RuntimeError("A2JModel (B200 build) needs CUDA tensors: there is no CPU fallback")
        if not self.is_3D:
            raise NotImplementedError("is_3D=False is not used by the pipeline and not built")
        if self.training and not getattr(self, "_warned_training", False):
            # the reference's load_pretrained_a2j returns the model without .eval() (handnet_pipeline.py:26-41): a caller
            # that never calls .eval() gets batch-statistics BatchNorm there; only eval-mode BN is built here
            import warnings
            warnings.warn("A2JModel (B200 build) always applies eval-mode BatchNorm (running statistics); call .eval() "
                          "to get the same behaviour from the reference", stacklevel=2)
            self._warned_training = True
        return self._executor.forward_device(x.float().contiguous())

    def head_outputs(self, x: torch.Tensor):
        """(classification [n,A,J], regression [n,A,J,2], depth [n,A,J]) as the reference towers return them."""
        cls, reg, dep, _ = self._executor.heads_device(x.float().contiguous())
        return cls, reg, dep

    def forward(self, x, gt=None):
        if gt is not None:
            raise NotImplementedError("the A2J training loss is outside the scope of the B200 inference build")
        return self.forward_device(x).cpu()         # reference returns the key points on the CPU (a2j.py:229)


class A2JModelLightning(nn.Module):
    """Inference stand-in for the reference's LightningModule (a2j.py:252-366): wraps ``self.a2j`` and loads the
    ``a2j.*`` keys of a Lightning ``.ckpt``; the train/val/test steps are out of scope."""

    def __init__(self, num_classes: int = 21, crop_height: int = 176, crop_width: int = 176, is_3D: bool = True,
                 is_RGBD: bool = False, spatial_factor: float = 0.5, **_unused):
        super().__init__()
        self.a2j = A2JModel(num_classes, crop_height, crop_width, is_3D=is_3D, is_RGBD=is_RGBD,
                            spatial_factor=spatial_factor)

    @classmethod
    def load_from_checkpoint(cls, path, **kwargs):
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        hp = dict(ckpt.get("hyper_parameters", {}))
        hp.update(kwargs)
        net = cls(**{k: v for k, v in hp.items() if k in ("num_classes", "crop_height", "crop_width", "is_3D",
                                                          "is_RGBD", "spatial_factor")})
        net.load_state_dict(ckpt["state_dict"], strict=False)
        return net

    def forward(self, x):
        return self.a2j(x)


class A2JDataModule:
    """Placeholder for the reference's LightningDataModule (needs DexYCB; out of scope)."""

    def __init__(self, *a, **k):
        raise NotImplementedError("A2JDataModule needs the DexYCB dataset and is outside the scope of this build")

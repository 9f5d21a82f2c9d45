"""CPU restatement of pose2mesh's FlatPose2Mesh forward pass (eval mode) -- TEST INFRASTRUCTURE ONLY.

Imported by tests/ and by bench.py's CPU leg, never by the product.  Plain dense torch on a state dict; every function cites
the reference lines it follows.  Pinned against the unmodified reference by oracle/make_golden_pose2mesh.py ->
tests/golden/pose2mesh_case.pt (tests/test_oracle_golden.py)."""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

MANO_FEATURES = [(5, 32, 64, 64), (64, 128, 256), (256, 256, 256), (256, 256, 256), (256, 256, 256), (256, 128, 128), (128, 64, 3)]
EPS = 1e-5


def _bn(sd: Dict[str, torch.Tensor], p: str, x: torch.Tensor) -> torch.Tensor:
    """nn.BatchNorm1d in eval mode."""
    return (x - sd[p + "running_mean"]) / torch.sqrt(sd[p + "running_var"] + EPS) * sd[p + "weight"] + sd[p + "bias"]


def cheby_conv(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, L: torch.Tensor, K: int) -> torch.Tensor:
    """pose2mesh/lib/models/backbones/cheby_graph_conv.py:5-38 without the BatchNorm: x [B, V, Fin], L dense [V, V]."""
    terms = [x]
    if K > 1:
        terms.append(torch.einsum("vw,bwf->bvf", L, x))                               # :26-27
    for _ in range(2, K):
        terms.append(2 * torch.einsum("vw,bwf->bvf", L, terms[-1]) - terms[-2])       # :28-31
    feat = torch.stack(terms, dim=3).reshape(x.shape[0] * x.shape[1], -1)             # B*V x Fin*K, K fastest (:33-35)
    return (feat @ w.t() + b).view(x.shape[0], x.shape[1], -1)                        # :38


def mesh_forward(sd: Dict[str, torch.Tensor], graph_L: Sequence[torch.Tensor], x: torch.Tensor, prefix: str = "pose2mesh.",
                 table=MANO_FEATURES) -> torch.Tensor:
    """pose2mesh/lib/models/meshnet.py:78-117.  graph_L: dense Laplacians, finest first, joint graph last, WITH the 48 x 48
    level still in place (it is dropped here as :41 does)."""
    graph_L = list(graph_L)
    del graph_L[-2]
    x = x.reshape(-1, graph_L[-1].shape[0], table[0][0])
    last = len(table) - 1
    li = 0
    for i, blk in enumerate(table):
        skip = x
        L = graph_L[-(i + 1) + (1 if i == last else 0)]                                # :93-95
        for j in range(len(blk) - 1):
            x = cheby_conv(x, sd[f"{prefix}cl.{li}.weight"], sd[f"{prefix}cl.{li}.bias"], L, 3)
            if not (i == last and j == len(blk) - 2):
                b, v, f = x.shape
                x = torch.relu(_bn(sd, f"{prefix}bn.{li}.", x.reshape(b * v, f))).view(b, v, f)   # :40-41, meshnet.py:100-101
            li += 1
        if i == 0:                                                                     # :105-107
            x = (x.reshape(x.shape[0], -1) @ sd[prefix + "fc.weight"].t() + sd[prefix + "fc.bias"]).view(
                -1, graph_L[-2].shape[0], table[1][0])
        elif i < last:
            x = F.interpolate(skip, size=x.shape[2], mode="linear") + x                # :109-110, 114-115
            if i < last - 1:
                x = x.repeat_interleave(2, dim=1)                                      # nn.Upsample(scale_factor=2) on V (:69-76)
    return x


def posenet_forward(sd: Dict[str, torch.Tensor], x: torch.Tensor, prefix: str = "pose_lifter.", stages: int = 2) -> torch.Tensor:
    """pose2mesh/lib/models/posenet.py:75-87 (top) and :25-41 (stage), eval mode (dropout = identity)."""
    lin = lambda p, t: t @ sd[p + "weight"].t() + sd[p + "bias"]
    y = lin(prefix + "w1.", x)
    for s in range(stages):
        p = f"{prefix}linear_stages.{s}."
        h = lin(p + "w1.", torch.relu(_bn(sd, p + "batch_norm1.", y)))
        y = y + lin(p + "w2.", torch.relu(_bn(sd, p + "batch_norm2.", h)))
    return lin(prefix + "w2.", y)


def flat_pose2mesh(sd: Dict[str, torch.Tensor], graph_L: Sequence[torch.Tensor], pose2d: torch.Tensor):
    """pose2mesh/lib/models/pose2mesh_net.py:18-24: (cam_mesh [B, V, 3], pose3d [B, J, 3])."""
    nj = pose2d.shape[1]
    pose3d = posenet_forward(sd, pose2d.reshape(len(pose2d), -1)).reshape(-1, nj, 3)
    return mesh_forward(sd, graph_L, torch.cat((pose2d, pose3d / 1000), dim=2)), pose3d

"""Pose2Mesh graph network (joints -> mesh vertices, coarse to fine) on the hn_b200 kernels.

Call surface, layer tables and parameter names of /root/reference/pose2mesh/lib/models/meshnet.py:12-64 (`cl`, `bn`, `fc`;
get_model :119-122) so that its checkpoints load unchanged.  graph_L: the rescaled Laplacians of
graph_utils.build_coarse_graphs, finest first, joint graph last (scipy sparse, torch sparse or ops.CsrGraph); the 48 x 48 level
is dropped as upstream does (:41).  Forward = :78-117: per block a run of Chebyshev graph convolutions (K = 3) with ReLU, then
block 0 -> fully connected lift to the coarsest mesh level; later blocks -> + the block input linearly interpolated along the
feature axis, vertices doubled (nearest) except after the last two blocks."""
from __future__ import annotations

import torch
import torch.nn as nn

from hn_b200 import ops

from .backbones.cheby_graph_conv import as_graph, graph_conv_cheby

MANO_FEATURES = [(None, 32, 64, 64), (64, 128, 256), (256, 256, 256), (256, 256, 256), (256, 256, 256), (256, 128, 128),
                 (128, 64, None)]                                    # meshnet.py:25-29 (first / last entry: in / out channels)
BODY_FEATURES = [(None, 32, 64, 64), (64, 128, 256), (256, 256, 256), (256, 256, 256), (256, 256, 256), (256, 256, 256),
                 (256, 128, 128), (128, 128, 128), (128, 128, 128), (128, 64, None)]          # meshnet.py:32-36


def _joint_set() -> str:
    try:                                                            # the demo's own config module, when it is importable
        from core.config import cfg
        return str(cfg.DATASET.target_joint_set)
    except Exception:
        return "mano"


class Pose2Mesh(nn.Module):
    def __init__(self, num_joint_input_chan, num_mesh_output_chan, graph_L, joint_set=None):
        super().__init__()
        self.num_joint_input_chan, self.num_mesh_output_chan = num_joint_input_chan, num_mesh_output_chan
        table = MANO_FEATURES if (joint_set or _joint_set()) == "mano" else BODY_FEATURES
        self.CL_F = [tuple(num_joint_input_chan if f is None and i == 0 else num_mesh_output_chan if f is None else f
                           for f in blk) for i, blk in enumerate(table)]
        self.CL_K = [3] * len(self.CL_F)
        graph_L = list(graph_L)
        del graph_L[-2]                                             # meshnet.py:41
        self.graph_L = graph_L
        v_joint, v_coarse = graph_L[-1].shape[0], graph_L[-2].shape[0]
        self.fc = nn.Linear(v_joint * self.CL_F[0][-1], v_coarse * self.CL_F[1][0])
        cl, bn = [], []
        last = len(self.CL_F) - 1
        for i, blk in enumerate(self.CL_F):
            for j in range(len(blk) - 1):
                fin, fout = self.CL_K[i] * blk[j], blk[j + 1]
                lin = nn.Linear(fin, fout)
                bound = (2.0 / (fin + fout)) ** 0.5                 # meshnet.py:53-55
                nn.init.uniform_(lin.weight, -bound, bound)
                nn.init.zeros_(lin.bias)
                cl.append(lin)
                bn.append(None if (i == last and j == len(blk) - 2) else nn.BatchNorm1d(fout))
        self.cl, self.bn = nn.ModuleList(cl), nn.ModuleList(bn)

    def forward(self, x):
        if self.training:
            raise NotImplementedError("pose2mesh on the B200 build is inference only: call .eval()")
        dev = x.device
        graphs = [as_graph(L, dev) for L in self.graph_L]
        x = x.reshape(-1, graphs[-1].n, self.num_joint_input_chan).contiguous().float()
        last = len(self.CL_F) - 1
        li = 0
        for i, blk in enumerate(self.CL_F):
            skip = x
            g = graphs[-(i + 1) + (1 if i == last else 0)]                                   # meshnet.py:93-95
            for j in range(len(blk) - 1):
                final = i == last and j == len(blk) - 2
                x = graph_conv_cheby(x, self.cl[li], self.bn[li], g, blk[j + 1], self.CL_K[i], relu=not final)
                li += 1
            if i == 0:                                                                       # joints -> coarsest mesh level
                b = x.shape[0]
                x = ops.linear_f32([x.view(b, -1)], self.fc.weight.detach().float().contiguous(),
                                   self.fc.bias.detach().float().contiguous()).view(b, graphs[-2].n, self.CL_F[1][0])
            elif i < last:
                x = ops.mesh_residual_upsample(x, skip, 2 if i < last - 1 else 1)
        return x


def get_model(num_joint_input_chan, num_mesh_output_chan, graph_L):
    return Pose2Mesh(num_joint_input_chan, num_mesh_output_chan, graph_L)

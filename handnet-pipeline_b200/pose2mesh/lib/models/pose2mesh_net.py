"""FlatPose2Mesh: 2D joints -> (camera-space mesh, 3D joints), /root/reference/pose2mesh/lib/models/pose2mesh_net.py:9-30.

get_model(num_joint, graph_L) is what ros_demo.py:145 calls; forward(pose2d [B, J, 2]) -> (mesh [B, V, 3], pose3d [B, J, 3])
with mesh = Pose2Mesh(cat(pose2d, pose3d / 1000)) (:18-24).  Submodule names (`pose_lifter`, `pose2mesh`) as upstream, so
`load_state_dict(checkpoint['model_state_dict'])` (ros_demo.py:147) works unchanged."""
from __future__ import annotations

import torch
import torch.nn as nn

from . import meshnet, posenet


class FlatPose2Mesh(nn.Module):
    def __init__(self, num_joint, graph_L):
        super().__init__()
        self.num_joint = num_joint
        self.pose_lifter = posenet.get_model(num_joint, hid_dim=4096, num_layer=2, p_dropout=0.5, pretrained=False)
        self.pose2mesh = meshnet.get_model(num_joint_input_chan=2 + 3, num_mesh_output_chan=3, graph_L=graph_L)

    def forward(self, pose2d):
        if not pose2d.is_cuda:
            raise RuntimeError("pose2mesh needs a CUDA tensor (libhandnet_b200 has no CPU path)")
        pose2d = pose2d.contiguous().float()
        pose3d = self.pose_lifter(pose2d.view(len(pose2d), -1)).view(-1, self.num_joint, 3)
        cam_mesh = self.pose2mesh(torch.cat((pose2d, pose3d / 1000), dim=2))
        return cam_mesh, pose3d


def get_model(num_joint, graph_L):
    return FlatPose2Mesh(num_joint, graph_L)

"""B200 build of pose2mesh's network package (last row).

Mirrors the import surface of /root/reference/pose2mesh/lib/models/__init__.py:1-4 for the modules the demo uses
(ros_demo.py:30,145): put handnet-pipeline_b200/pose2mesh/lib on sys.path in front of the reference's pose2mesh/lib and
`models.pose2mesh_net.get_model(joint_num, graph_L)` builds the CUDA-kernel network; graph building, MANO and the camera
layer (training only) stay the caller's."""
from . import meshnet, pose2mesh_net, posenet  # noqa: F401

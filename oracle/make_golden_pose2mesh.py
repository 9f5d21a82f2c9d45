"""Golden vector for the pose2mesh row: run the UNMODIFIED reference network on the CPU -- THIS CONTAINER ONLY.

    python oracle/make_golden_pose2mesh.py        ->  tests/golden/pose2mesh_case.pt

The demo's checkpoint and the MANO model are not available offline, so the case is synthetic but structurally the demo's
(ros_demo.py:131-148): the reference's own graph_utils.build_coarse_graphs(levels=6) over an icosphere mesh (642 vertices ->
Laplacians of 1024 / 512 / 256 / 128 / 64 / 32 / 21 nodes) and models.pose2mesh_net.get_model(21, graph_L) loaded with
hn_b200.synth.fill_state_dict(seed) (74.6 M parameters, regenerated from the seed by the tests, not stored).  Stored: the
Laplacians (COO), the (key, shape) list of the reference's state dict, the input joints and the reference's outputs."""
import os
import sys
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
REF_LIB = "/root/reference/pose2mesh/lib"


def icosphere(subdivisions: int) -> np.ndarray:
    t = (1 + 5 ** 0.5) / 2
    v = [np.array(p, float) for p in [(-1, t, 0), (1, t, 0), (-1, -t, 0), (1, -t, 0), (0, -1, t), (0, 1, t), (0, -1, -t), (0, 1, -t),
                                      (t, 0, -1), (t, 0, 1), (-t, 0, -1), (-t, 0, 1)]]
    f = [(0, 11, 5), (0, 5, 1), (0, 1, 7), (0, 7, 10), (0, 10, 11), (1, 5, 9), (5, 11, 4), (11, 10, 2), (10, 7, 6), (7, 1, 8),
         (3, 9, 4), (3, 4, 2), (3, 2, 6), (3, 6, 8), (3, 8, 9), (4, 9, 5), (2, 4, 11), (6, 2, 10), (8, 6, 7), (9, 8, 1)]
    for _ in range(subdivisions):
        cache, nf = {}, []

        def mid(a, b):
            k = (min(a, b), max(a, b))
            if k not in cache:
                cache[k] = len(v)
                v.append((v[a] + v[b]) / 2)
            return cache[k]
        for a, b, c in f:
            ab, bc, ca = mid(a, b), mid(b, c), mid(c, a)
            nf += [(a, ab, ca), (b, bc, ab), (c, ca, bc), (ab, bc, ca)]
        f = nf
    return np.array(f)


def main():
    # the reference's network modules import `core.config` (which creates directories at import) and `funcs_utils`: stubs
    class D(dict):
        __getattr__ = dict.__getitem__
    core, cfgm, fu = types.ModuleType("core"), types.ModuleType("core.config"), types.ModuleType("funcs_utils")
    cfgm.cfg = D(DATASET=D(target_joint_set="mano"), MODEL=D(posenet_pretrained=False, posenet_path=""))
    core.config = cfgm
    fu.load_checkpoint = lambda **k: None
    sys.modules.update({"core": core, "core.config": cfgm, "funcs_utils": fu})
    sys.path.insert(0, REF_LIB)
    from graph_utils import build_coarse_graphs                    # the reference's own graph coarsening
    import models.pose2mesh_net as ref_net
    from hn_b200 import synth

    skeleton = ((0, 1), (0, 5), (0, 9), (0, 13), (0, 17), (1, 2), (2, 3), (3, 4), (5, 6), (6, 7), (7, 8), (9, 10), (10, 11), (11, 12),
                (13, 14), (14, 15), (15, 16), (17, 18), (18, 19), (19, 20))                      # ros_demo.py:134
    hori = ((1, 5), (5, 9), (9, 13), (13, 17), (2, 6), (6, 10), (10, 14), (14, 18), (3, 7), (7, 11), (11, 15), (15, 19), (4, 8),
            (8, 12), (12, 16), (16, 20))                                                           # ros_demo.py:135-137
    np.random.seed(0)
    _, graph_L, _, _ = build_coarse_graphs(icosphere(3), 21, skeleton, hori, levels=6)
    coo = []
    for L in graph_L:
        c = L.tocoo()
        coo.append({"shape": tuple(c.shape), "row": torch.from_numpy(c.row.astype(np.int64)),
                    "col": torch.from_numpy(c.col.astype(np.int64)), "val": torch.from_numpy(c.data.astype(np.float32))})
    model = ref_net.get_model(21, list(graph_L))
    shapes = [(k, tuple(v.shape)) for k, v in model.state_dict().items()]
    model.load_state_dict(synth.fill_state_dict(shapes, seed=7))
    model.eval()
    g = torch.Generator().manual_seed(11)
    pose2d = torch.randn(3, 21, 2, generator=g)                      # ros_demo.py:153-157 normalises the joints to zero mean / unit std
    torch.Tensor.cuda = lambda self, *a, **k: self                   # meshnet.py:80 moves the Laplacians with .cuda(); stay on the CPU
    with torch.no_grad():
        mesh, pose3d = model(pose2d)
    out = os.path.join(ROOT, "tests", "golden", "pose2mesh_case.pt")
    torch.save({"graph_L": coo, "shapes": shapes, "seed": 7, "pose2d": pose2d, "mesh": mesh, "pose3d": pose3d}, out)
    print("wrote", out, "mesh", tuple(mesh.shape), "pose3d", tuple(pose3d.shape), "|mesh| max", float(mesh.abs().max()))


if __name__ == "__main__":
    main()

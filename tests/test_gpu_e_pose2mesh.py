"""GPU parity of the pose2mesh row: the three fp32 kernels against plain torch, and the whole FlatPose2Mesh
through this package's `models` modules against the CPU oracle and the reference-generated golden case.  All calls go
through the C ABI (hn_b200.ops -> libhandnet_b200.so)."""
import pytest
import scipy.sparse as sp
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from hn_b200 import ops as _ops
    return _ops


def _random_laplacian(n, g, nnz_per_row=5):
    rows = torch.arange(n).repeat_interleave(nnz_per_row)
    cols = torch.randint(0, n, (n * nnz_per_row,), generator=g)
    vals = torch.randn(n * nnz_per_row, generator=g) * 0.4
    return sp.coo_matrix((vals.numpy(), (rows.numpy(), cols.numpy())), shape=(n, n)).tocsr()      # duplicates are summed


@pytest.mark.parametrize("b,v,f", [(1, 21, 5), (3, 64, 64), (2, 1024, 256), (2, 100, 33)])
def test_cheby_spmm_matches_dense(ops, b, v, f):
    """T1 = L x and T2 = 2 L T1 - T0 (cheby_graph_conv.py:26-31) from scipy, torch-sparse and dense Laplacians."""
    g = torch.Generator().manual_seed(v + f)
    Ls = _random_laplacian(v, g)
    Ld = torch.from_numpy(Ls.toarray()).float()
    x = torch.randn(b, v, f, generator=g)
    t1_ref = torch.einsum("vw,bwf->bvf", Ld, x)
    t2_ref = 2 * torch.einsum("vw,bwf->bvf", Ld, t1_ref) - x
    for lap in (Ls, Ld.to_sparse(), Ld):
        gr = ops.CsrGraph(lap)
        t1 = ops.cheby_spmm(gr, x.cuda())
        t2 = ops.cheby_spmm(gr, t1, z=x.cuda(), alpha=2.0, beta=-1.0)
        torch.cuda.synchronize()
        torch.testing.assert_close(t1.cpu(), t1_ref, rtol=1e-5, atol=1e-5)
        torch.testing.assert_close(t2.cpu(), t2_ref, rtol=1e-5, atol=2e-5)


@pytest.mark.parametrize("m,fin,n,planes", [(1, 42, 4096, 1), (3, 4096, 63, 1), (2 * 1024, 256, 256, 3), (63, 5, 32, 3), (17, 1344, 100, 1),
                                            (40, 33, 7, 2)])
def test_linear_f32_planes_affines_residual(ops, m, fin, n, planes):
    """hn_linear_f32: interleaved Chebyshev planes (kk = f * planes + p), BatchNorm + ReLU in front, BatchNorm (+ ReLU) behind,
    residual add -- every combination the pose2mesh modules use, ragged row / column / K tiles."""
    g = torch.Generator().manual_seed(m + n)
    a = [torch.randn(m, fin, generator=g) for _ in range(planes)]
    w = torch.randn(n, fin * planes, generator=g) * (fin * planes) ** -0.5
    bias = torch.randn(n, generator=g) * 0.1
    isc, ish = 0.5 + torch.rand(fin * planes, generator=g), torch.randn(fin * planes, generator=g) * 0.3
    osc, osh = 0.5 + torch.rand(n, generator=g), torch.randn(n, generator=g) * 0.3
    res = torch.randn(m, n, generator=g)
    feat = torch.stack(a, dim=2).reshape(m, fin * planes)
    dev = lambda t: t.cuda()
    cases = [
        (dict(), feat @ w.t() + bias),
        (dict(out_affine=(dev(osc), dev(osh)), relu=True), torch.relu((feat @ w.t() + bias) * osc + osh)),
        (dict(in_affine=(dev(isc), dev(ish)), res=dev(res)), torch.relu(feat * isc + ish) @ w.t() + bias + res),
    ]
    for kw, ref in cases:
        y = ops.linear_f32([dev(t) for t in a], dev(w), dev(bias), **kw)
        torch.cuda.synchronize()
        torch.testing.assert_close(y.cpu(), ref, rtol=2e-5, atol=2e-5 * max(1.0, float(ref.abs().max())))
    y = ops.linear_f32([dev(t) for t in a], dev(w), None)
    torch.cuda.synchronize()
    torch.testing.assert_close(y.cpu(), feat @ w.t(), rtol=2e-5, atol=2e-5)


@pytest.mark.parametrize("b,v,fout,fskip,up", [(2, 64, 256, 64, 2), (1, 128, 256, 256, 2), (3, 1024, 128, 256, 1), (2, 7, 10, 3, 2)])
def test_mesh_residual_upsample_matches_interpolate(ops, b, v, fout, fskip, up):
    """x + F.interpolate(skip, size=F, mode='linear') along the feature axis, vertices repeated (meshnet.py:107-114, 69-76)."""
    g = torch.Generator().manual_seed(fout + fskip)
    x, skip = torch.randn(b, v, fout, generator=g), torch.randn(b, v, fskip, generator=g)
    ref = (F.interpolate(skip, size=fout, mode="linear") + x).repeat_interleave(up, dim=1)
    out = ops.mesh_residual_upsample(x.cuda(), skip.cuda(), up)
    torch.cuda.synchronize()
    torch.testing.assert_close(out.cpu(), ref, rtol=1e-6, atol=1e-6)


def test_flat_pose2mesh_matches_reference_and_oracle(golden):
    """The demo's call sequence (ros_demo.py:145-148,160-162): models.pose2mesh_net.get_model(21, graph_L), load_state_dict,
    .cuda().eval(), model(joints) -> (mesh, pose3d).  Against the unmodified reference's outputs (golden file) and the CPU oracle;
    also one hand at a time == the batch (no cross-sample state)."""
    import models.pose2mesh_net as net
    from hn_b200 import ops, synth
    from oracle import pose2mesh_oracle
    case = golden("pose2mesh_case.pt")
    graph_L = [sp.coo_matrix((c["val"].numpy(), (c["row"].numpy(), c["col"].numpy())), shape=c["shape"]).tocsr()
               for c in case["graph_L"]]
    sd = synth.fill_state_dict(case["shapes"], seed=case["seed"])
    model = net.get_model(21, graph_L)
    model.load_state_dict(sd)
    model = model.cuda().eval()
    n0 = ops.launch_count()
    with torch.no_grad():
        mesh, pose3d = model(case["pose2d"].cuda())
        one = [model(case["pose2d"][i:i + 1].cuda()) for i in range(len(case["pose2d"]))]
    torch.cuda.synchronize()
    assert ops.launch_count() - n0 >= 4 * 50, "the CUDA kernels did not run"
    assert mesh.shape == case["mesh"].shape and pose3d.shape == case["pose3d"].shape
    torch.testing.assert_close(pose3d.cpu(), case["pose3d"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(mesh.cpu(), case["mesh"], rtol=1e-4, atol=1e-4)
    dense = []
    for c in case["graph_L"]:
        L = torch.zeros(c["shape"])
        L.index_put_((c["row"], c["col"]), c["val"], accumulate=True)
        dense.append(L)
    o_mesh, o_pose = pose2mesh_oracle.flat_pose2mesh(sd, dense, case["pose2d"])
    torch.testing.assert_close(mesh.cpu(), o_mesh, rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(pose3d.cpu(), o_pose, rtol=1e-4, atol=1e-4)
    for i, (m1, p1) in enumerate(one):
        assert torch.equal(m1[0], mesh[i]) and torch.equal(p1[0], pose3d[i]), "per-hand calls must equal the batched call"
    # the mesh vertices the demo keeps (ros_demo.py:162: graph_perm_reverse[:face.max() + 1]) are a gather of these rows
    with pytest.raises(NotImplementedError):
        model.train()(case["pose2d"].cuda())

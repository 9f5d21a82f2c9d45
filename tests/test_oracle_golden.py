"""Pin the oracle: replay oracle/ against the fixtures the real reference produced
(oracle/make_golden.py, run in the build container).  CPU only."""
import numpy as np
import pytest
import torch

from hn_b200 import synth
from oracle import a2j_oracle, fcos_oracle, handnet_oracle, nms_oracle
from oracle.golden_inputs import inputs_images, pad_crop_inputs, sample, stress_head_tensors


def _same(a, b, tol=0.0):
    assert a.shape == b.shape, (a.shape, b.shape)
    if tol == 0.0:
        assert torch.equal(a, b)
    else:
        torch.testing.assert_close(a, b, rtol=tol, atol=tol)


def canon_keep(keep, scores):
    """torchvision's per-class branch (numel > 4000) ends in a NON-stable descending sort
    (torchvision/ops/boxes.py:121), so the order inside a group of exactly equal scores is an
    artefact of torch.sort.  Canonical form: ties in ascending index order."""
    keep = np.asarray(keep)
    return keep[np.lexsort((keep, -scores[keep].astype(np.float64)))]


def canon_dets(d):
    """Same canonicalisation for detection dicts (ties broken by box coordinates)."""
    b, s = d["boxes"].numpy().astype(np.float64), d["scores"].numpy().astype(np.float64)
    order = torch.from_numpy(np.lexsort((b[:, 3], b[:, 2], b[:, 1], b[:, 0], -s)))
    return {k: v[order] for k, v in d.items() if torch.is_tensor(v) and v.shape[:1] == d["scores"].shape}


def test_nms_known_answers(golden):
    """oracle NMS == torchvision.ops.nms / batched_nms, bit-exact, incl. ties, the exact-threshold
    IoU case and the numel>4000 strategy switch."""
    for c in golden("nms_cases.pt")["cases"]:
        b, s, lab = c["boxes"].numpy(), c["scores"].numpy(), c["labels"].numpy()
        assert np.array_equal(nms_oracle.nms(b, s, c["thr"]), c["keep_nms"].numpy())
        mine, ref = nms_oracle.batched_nms(b, s, lab, c["thr"]), c["keep_batched"].numpy()
        if b.size > 4000:
            assert sorted(mine.tolist()) == sorted(ref.tolist())
            mine, ref = canon_keep(mine, s), canon_keep(ref, s)
        assert np.array_equal(mine, ref)


def test_nms_oracle_matches_installed_torchvision_live():
    tv = pytest.importorskip("torchvision")
    g = torch.Generator().manual_seed(123)
    for n in (5, 130, 1200):
        xy = torch.rand(n, 2, generator=g) * 100
        b = torch.cat((xy, xy + torch.rand(n, 2, generator=g) * 30 + 1), 1)
        s = torch.rand(n, generator=g)
        lab = torch.randint(0, 3, (n,), generator=g)
        ref = tv.ops.batched_nms(b, s, lab, 0.3).numpy()
        mine = nms_oracle.batched_nms(b.numpy(), s.numpy(), lab.numpy(), 0.3)
        assert np.array_equal(canon_keep(mine, s.numpy()), canon_keep(ref, s.numpy()))


def test_fcos_small_matches_reference(golden):
    fx = golden("fcos_small.pt")
    cfg = fx["cfg"]
    sd = synth.fcos_state_dict(cfg["num_classes"], cfg["ext"], seed=cfg["seed_w"])
    imgs = inputs_images(cfg["seed_x"], cfg["n"], cfg["h"], cfg["w"])
    with torch.inference_mode():
        dets, taps = fcos_oracle.fcos_forward(sd, imgs, cfg["num_classes"], cfg["ext"], cfg["min_size"],
                                              cfg["max_size"], return_taps=True)
    assert list(taps["canvas"].shape) == fx["canvas_shape"].tolist()
    assert [list(s) for s in taps["image_sizes"]] == fx["image_sizes"].tolist()
    _same(sample(taps["canvas"])[0], fx["canvas_s"])
    for i in range(3):
        _same(sample(taps["p"][i])[0], fx[f"p{i}_s"], 1e-5)
    for k in ("cls_logits", "bbox_regression", "bbox_ctrness", "hand_lr"):
        _same(sample(taps["head"][k])[0], fx[f"head_{k}_s"], 1e-5)
    for d, r in zip(dets, fx["dets"]):
        _same(d["labels"], r["labels"])
        _same(d["sides"], r["sides"])
        _same(d["feature_idx"], r["feature_idx"])
        _same(d["boxes"], r["boxes"], 1e-4)
        _same(d["scores"], r["scores"], 1e-6)


def test_fcos_ext_heads_match_reference(golden):
    fx = golden("fcos_ext_small.pt")
    cfg = fx["cfg"]
    sd = synth.fcos_state_dict(cfg["num_classes"], True, seed=cfg["seed_w"])
    imgs = inputs_images(cfg["seed_x"], cfg["n"], cfg["h"], cfg["w"])
    with torch.inference_mode():
        dets = fcos_oracle.fcos_forward(sd, imgs, cfg["num_classes"], True, cfg["min_size"], cfg["max_size"])
    for d, r in zip(dets, fx["dets"]):
        _same(d["labels"], r["labels"])
        _same(d["contacts"], r["contacts"])
        _same(d["sides"], r["sides"])
        _same(d["dxdymags"], r["dxdymags"], 1e-5)
        _same(d["boxes"], r["boxes"], 1e-4)


def test_a2j_matches_reference(golden):
    fx = golden("a2j_small.pt")
    cfg = fx["cfg"]
    sd = synth.a2j_state_dict(seed=cfg["seed_w"])
    g = torch.Generator().manual_seed(cfg["seed_x"])
    x = torch.rand(cfg["n"], 1, 176, 176, generator=g) * 1.5
    with torch.inference_mode():
        joints, taps = a2j_oracle.a2j_forward(sd, x, return_taps=True)
    for k in ("c4", "c5", "cls", "reg", "dep"):
        torch.testing.assert_close(sample(taps[k])[0], fx[k + "_s"], rtol=1e-4, atol=1e-4)
    # joints are softmax-weighted means over 1936 anchors: 1e-3 relative as BASELINE.json states
    torch.testing.assert_close(joints, fx["joints"], rtol=1e-3, atol=1e-3)
    assert torch.equal(a2j_oracle.all_anchors(), sd["post_process.all_anchors"])
    uv = a2j_oracle.convert_joints(fx["joints"][0].numpy(), fx["convert_box"].numpy())
    np.testing.assert_allclose(uv, fx["convert_uv"].numpy(), rtol=1e-6)
    xyz = a2j_oracle.convert_joints(fx["joints"][0].numpy(), fx["convert_box"].numpy(), fx["convert_paras"].numpy())
    np.testing.assert_allclose(xyz, fx["convert_xyz"].numpy(), rtol=1e-5)


def test_pad_and_crop_match_reference(golden):
    """S1 (int64 truncation + 0.4 pad in float32) and S2 (inclusive slice, legacy nearest)."""
    fx = golden("pad_crop_cases.pt")
    hh, ww = fx["hw"].tolist()
    boxes, depth = pad_crop_inputs(fx["seed"], fx["nb"], hh, ww)
    for i in range(fx["nb"]):
        pb = handnet_oracle.pad_box(boxes[i].numpy(), hh, ww)
        assert pb.tolist() == fx["crops"][i].tolist(), i
        crop = handnet_oracle.crop_resize(depth[i], pb)
        assert torch.equal(crop[:, ::4, ::4], fx["depth_batch_s4"][i])
        assert crop.double().sum().item() == fx["depth_batch_sum"][i].item()


def test_nearest_index_formula_against_interpolate():
    for n in list(range(1, 400)) + [640, 699, 1333]:
        ref = torch.nn.functional.interpolate(torch.arange(n, dtype=torch.float32)[None, None, None, :], size=(1, 176))
        mine = [handnet_oracle.nearest_src_index(d, n, 176) for d in range(176)]
        assert ref.reshape(-1).long().tolist() == mine, n


def test_postprocess_stress_matches_reference(golden):
    """BASELINE.json config 4: ~10k candidates/frame through decode, 0.7 cut and per-class NMS."""
    fx = golden("postprocess_stress.pt")
    cfg = fx["cfg"]
    npl = [gh * gw for gh, gw in cfg["grids"]]
    ho = stress_head_tensors(cfg["seed"], cfg["batch"], sum(npl), 3, cfg["mu"])
    anchors = fcos_oracle.anchors_for(tuple(cfg["canvas"]), cfg["grids"])
    _same(torch.stack((anchors[0], anchors[-1])), fx["anchors_first_last"])
    dets = fcos_oracle.postprocess(ho, anchors, npl, [(800, 1066)] * cfg["batch"], [(480, 640)] * cfg["batch"])
    for d, r in zip(dets, fx["dets"]):
        d, r = canon_dets(d), canon_dets(r)
        for k in ("boxes", "scores", "labels", "sides", "feature_idx"):
            _same(d[k], r[k])


@pytest.mark.timeout(600)
def test_handnet_vga_matches_reference(golden):
    fx = golden("handnet_vga.pt")
    cfg = fx["cfg"]
    fsd = synth.fcos_state_dict(3, False, seed=cfg["seed_fcos"])
    asd = synth.a2j_state_dict(seed=cfg["seed_a2j"])
    imgs = inputs_images(cfg["seed_x"], cfg["n"], cfg["h"], cfg["w"])
    g = torch.Generator().manual_seed(cfg["seed_x"] + 1)
    depth = torch.rand(cfg["n"], 1, cfg["h"], cfg["w"], generator=g) * 1.5
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.inference_mode():
        dets = fcos_oracle.fcos_forward(fsd, imgs, 3)
        final, depth_batch, crops, hit = handnet_oracle.handnet_forward(fsd, asd, imgs, depth, detections=dets)
    assert [len(d["boxes"]) for d in dets] == fx["n_kept"].tolist()
    _same(torch.stack([d["labels"][:8] for d in dets]), fx["top_labels"])
    _same(torch.stack([d["boxes"][:8] for d in dets]), fx["top_boxes"], 1e-4)
    _same(torch.stack([d["scores"][:8] for d in dets]), fx["top_scores"], 1e-6)
    _same(crops, fx["crops"])
    _same(sample(depth_batch)[0], fx["depth_batch_s"])
    torch.testing.assert_close(final, fx["final"], rtol=1e-3, atol=1e-3)
    assert hit.all()


def test_handnet_oracle_max_hands_is_the_per_box_path_of_the_reference():
    """max_hands = H (extension; the reference keeps boxes[:1], handnet_pipeline.py:84-85): slot (i, k) is exactly what the
    max_hands = 1 oracle -- pinned to the reference above -- returns for a frame whose FIRST hand box is box k; frames with
    fewer hand boxes leave the remaining slots empty."""
    asd = synth.a2j_state_dict(seed=3)
    g = torch.Generator().manual_seed(5)
    imgs = [torch.rand(3, 120, 160, generator=g) for _ in range(3)]
    depth = torch.rand(3, 1, 120, 160, generator=g) * 1.5
    mk = lambda boxes, labels: {"boxes": torch.tensor(boxes, dtype=torch.float32).reshape(-1, 4), "labels": torch.tensor(labels, dtype=torch.int64)}
    dets = [mk([[10.2, 20.7, 60.1, 90.9], [5, 5, 30, 30], [100.5, 40.5, 150.5, 110.5], [0, 0, 159, 119]], [2, 1, 2, 2]),
            mk([[30, 30, 80, 70]], [2]),
            mk([[1, 1, 9, 9]], [1])]
    with torch.inference_mode():
        final, db, crops, hit = handnet_oracle.handnet_forward(None, asd, imgs, depth, detections=dets, max_hands=3)
        assert final.shape == (3, 3, 21, 3) and hit.tolist() == [[True, True, True], [True, False, False], [False] * 3]
        assert db.shape == (4, 1, 176, 176) and crops.shape == (4, 4)
        row = 0
        for i, d in enumerate(dets):
            hand = d["boxes"][d["labels"] == 2]
            for k in range(min(3, len(hand))):
                one = [mk(hand[k].tolist(), [2]) if j == i else mk([], []) for j in range(3)]
                f1, db1, c1, h1 = handnet_oracle.handnet_forward(None, asd, imgs, depth, detections=one)
                assert h1.tolist() == [j == i for j in range(3)]
                _same(crops[row], c1[0])
                _same(db[row], db1[0])
                _same(final[i, k], f1[i])
                row += 1
        # and max_hands = 1 is unchanged: first hand box only
        f0, db0, c0, h0 = handnet_oracle.handnet_forward(None, asd, imgs, depth, detections=dets)
        assert f0.shape == (3, 21, 3) and h0.tolist() == [True, True, False]
        _same(f0[0], final[0, 0])
        _same(c0, torch.stack((crops[0], crops[3])))


# ------------------------------------------------------------------------------------------------
# pose2mesh (last row): oracle and the mirrored modules against the reference-generated golden case
# ------------------------------------------------------------------------------------------------
def _dense_laplacians(case):
    out = []
    for c in case["graph_L"]:
        L = torch.zeros(c["shape"])
        L.index_put_((c["row"], c["col"]), c["val"], accumulate=True)
        out.append(L)
    return out


def test_pose2mesh_oracle_matches_reference(golden):
    """oracle/pose2mesh_oracle.py == the unmodified FlatPose2Mesh (oracle/make_golden_pose2mesh.py) on the synthetic demo-shaped
    case: weights regenerated from the seed (hn_b200.synth.fill_state_dict), Laplacians from the golden file."""
    from hn_b200 import synth
    from oracle import pose2mesh_oracle
    case = golden("pose2mesh_case.pt")
    sd = synth.fill_state_dict(case["shapes"], seed=case["seed"])
    mesh, pose3d = pose2mesh_oracle.flat_pose2mesh(sd, _dense_laplacians(case), case["pose2d"])
    torch.testing.assert_close(pose3d, case["pose3d"], rtol=1e-4, atol=1e-4)
    torch.testing.assert_close(mesh, case["mesh"], rtol=1e-4, atol=1e-4)


def test_pose2mesh_modules_mirror_the_reference_state_dict(golden):
    """models.pose2mesh_net.get_model(21, graph_L) (ros_demo.py:145) of this package has exactly the reference's state-dict keys
    and shapes, so checkpoint['model_state_dict'] loads unchanged (ros_demo.py:146-147)."""
    import scipy.sparse as sp
    import models.pose2mesh_net as net
    case = golden("pose2mesh_case.pt")
    graph_L = [sp.coo_matrix((c["val"].numpy(), (c["row"].numpy(), c["col"].numpy())), shape=c["shape"]).tocsr()
               for c in case["graph_L"]]
    m = net.get_model(21, graph_L)
    assert [(k, tuple(v.shape)) for k, v in m.state_dict().items()] == [(k, tuple(s)) for k, s in case["shapes"]]
    assert len(graph_L) == 7, "the caller's list is not modified (the reference deletes the 48 x 48 level in place)"
    m.eval()
    with pytest.raises(RuntimeError):
        m(case["pose2d"])                       # CPU tensors: no CPU path

"""CPU-only checks of the reference-facing Python surface: module paths, constructor
signatures, state-dict keys and shapes, old-checkpoint key conversion, loud failure without a GPU."""
import inspect

import numpy as np
import pytest
import torch

from hn_b200 import synth


def test_fcos_state_dict_keys_match_reference_tables():
    from fcos_utils.fcos import FCOS
    for ext, n in ((False, 232), (True, 236)):
        m = FCOS(num_classes=3, ext=ext)
        sd = m.state_dict()
        ref = synth.fcos_state_dict(3, ext)
        assert list(sd.keys()) == list(ref.keys())
        assert len(sd) == n
        for k in sd:
            assert sd[k].shape == ref[k].shape, k
        assert m.load_state_dict(ref, strict=True)


def test_fcos_accepts_torchvision_0_11_fpn_keys():
    from fcos_utils.fcos import FCOS
    ref = synth.fcos_state_dict(3, False)
    old = {k.replace("_blocks.0.0.", "_blocks.0.").replace("_blocks.1.0.", "_blocks.1.").replace("_blocks.2.0.", "_blocks.2."): v
           for k, v in ref.items()}
    assert "backbone.fpn.inner_blocks.1.weight" in old
    m = FCOS(num_classes=3, ext=False)
    res = m.load_state_dict(old, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    assert torch.equal(m.state_dict()["backbone.fpn.inner_blocks.1.0.weight"], ref["backbone.fpn.inner_blocks.1.0.weight"])


def test_fcos_signature_and_constants():
    from fcos_utils.fcos import FCOS, FCOSHead, psum, resize_boxes
    params = list(inspect.signature(FCOS.__init__).parameters)
    assert params[1:] == ["num_classes", "ext", "min_size", "max_size", "image_mean", "image_std", "anchor_generator",
                          "head", "center_sampling_radius", "score_thresh", "nms_thresh", "detections_per_img",
                          "topk_candidates"]
    m = FCOS(num_classes=3, ext=False, nms_thresh=0.5)
    assert (m.score_cut, m.nms_iou) == (0.7, 0.3) and m.nms_thresh == 0.5
    assert abs(float(m.head.classification_head.cls_logits.bias[0]) + np.log(99)) < 1e-6
    assert psum([3, 4]) == [0, 3, 7]
    b = resize_boxes(torch.tensor([[0.0, 0.0, 1066.0, 800.0]]), [800, 1066], [480, 640])
    assert torch.allclose(b, torch.tensor([[0.0, 0.0, 640.0, 480.0]]), atol=1e-3)
    assert isinstance(m.head, FCOSHead)


def test_a2j_state_dict_keys_match_reference_tables():
    from a2j.a2j import A2JModel
    m = A2JModel(21, 176, 176)
    sd = m.state_dict()
    ref = synth.a2j_state_dict()
    assert len(sd) == 414
    assert set(sd.keys()) == set(ref.keys())
    for k in sd:
        assert sd[k].shape == ref[k].shape, k
    assert torch.equal(sd["post_process.all_anchors"], ref["post_process.all_anchors"])
    assert m.load_state_dict(ref, strict=True)


def test_anchor_helpers_match_oracle():
    from a2j.anchor import generate_anchors, shift
    from fcos_utils.anchor_utils import AnchorGenerator
    from oracle import a2j_oracle, fcos_oracle
    a = torch.from_numpy(shift([11, 11], 16, generate_anchors())).float()
    assert torch.equal(a, a2j_oracle.all_anchors())

    class IL:
        tensors = torch.zeros(1, 3, 800, 1088)
        image_sizes = [(800, 1066)]
    grids = [(100, 136), (50, 68), (25, 34)]
    gen = AnchorGenerator(((8,), (16,), (32,)), ((1.0,),) * 3)
    got = gen(IL(), [torch.zeros(1, 1, h, w) for h, w in grids])[0]
    assert torch.equal(got, fcos_oracle.anchors_for((800, 1088), grids))


def test_box_coder_and_convert_joints_match_oracle():
    from a2j.a2j import convert_joints
    from fcos_utils.det_utils import BoxLinearCoder
    from oracle import a2j_oracle, fcos_oracle
    g = torch.Generator().manual_seed(0)
    anchors = fcos_oracle.anchors_for((64, 96), [(8, 12)], sizes=(8,))
    rel = torch.rand(96, 4, generator=g) * 3
    assert torch.equal(BoxLinearCoder(True).decode_single(rel, anchors), fcos_oracle.decode_boxes(rel, anchors))
    j = torch.rand(21, 3, generator=g).numpy() * 176
    box = np.array([10.0, 20.0, 200.0, 180.0], dtype=np.float32)
    par = np.array([600.0, 601.0, 320.0, 240.0], dtype=np.float32)
    np.testing.assert_allclose(convert_joints(j, None, box, par, 176, 176), a2j_oracle.convert_joints(j, box, par), rtol=1e-6)
    np.testing.assert_allclose(convert_joints(j, None, box, None, 176, 176), a2j_oracle.convert_joints(j, box), rtol=1e-6)


def test_handnet_surface():
    from handnet_pipeline.handnet_pipeline import HandNet, load_pretrained_a2j, load_pretrained_fcos  # noqa: F401

    class Args:
        pretrained_fcos = ""
        pretrained_a2j = ""
    net = HandNet(Args(), reload_detector=False, num_classes=3, reload_a2j=False, RGBD=False).eval()
    assert net.num_classes == 3 and net.RGBD is False
    assert not any(p.requires_grad for p in net.parameters())
    assert list(inspect.signature(net.forward).parameters) == ["images", "depth_images", "is_3D", "is_detect"]
    assert net([torch.zeros(3, 8, 8)], None, is_detect=True) is None
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):          # no silent CPU fallback
            net([torch.zeros(3, 32, 32)], torch.zeros(1, 1, 32, 32))


def test_resized_size_matches_oracle():
    from hn_b200.runtime import resized_size
    from oracle import fcos_oracle
    for h, w in ((480, 640), (1080, 1920), (120, 160), (97, 131), (333, 1000)):
        assert resized_size(h, w, 800, 1333) == fcos_oracle.resized_size(h, w, 800, 1333)
    assert resized_size(480, 640, 800, 1333) == (800, 1066)
    assert resized_size(1080, 1920, 800, 1333) == (749, 1333)


def test_weight_tiling_round_trip_and_layout():
    """ops.tile_k: [cout_pad, K] -> k-block-major [K/64][cout_pad][64] (what hn_conv2d_bf16 streams), and back."""
    from hn_b200 import ops
    g = torch.Generator().manual_seed(0)
    w = torch.randn(48, 32 * 6, generator=g).to(torch.bfloat16)             # K = 192 = 3 k-blocks
    t = ops.tile_k(w)
    assert t.shape == w.shape and t.is_contiguous()
    assert torch.equal(ops.untile_k(t), w)
    blocks = t.view(3, 48, 64)
    for kb in range(3):
        assert torch.equal(blocks[kb], w[:, kb * 64:(kb + 1) * 64])
    # pack_conv_weight: tap-major K (k = (r*kw + s)*cin + c), zero rows beyond cout, then tiled
    wt = torch.randn(5, 64, 3, 3, generator=g)
    p = ops.pack_conv_weight(wt)
    assert p.shape == (16, 9 * 64) and p.dtype == torch.bfloat16
    flat = ops.untile_k(p)
    assert torch.equal(flat[:5], wt.permute(0, 2, 3, 1).reshape(5, -1).to(torch.bfloat16)) and flat[5:].abs().sum() == 0
    # stem weights: K = 4 kernel-row pairs x 8 pixels x 2 rows x 4 channels (the order of the row-pair frame);
    # px = 0 / r = 7 / ch = 3 are zero
    ws = torch.randn(64, 3, 7, 7, generator=g)
    s = ops.untile_k(ops.pack_stem_weight(ws, 256)).view(64, 4, 8, 2, 4).permute(0, 1, 3, 2, 4).reshape(64, 8, 8, 4)
    assert s[:, 7].abs().sum() == 0 and s[:, :, 0].abs().sum() == 0 and s[..., 3].abs().sum() == 0
    assert torch.equal(s[:, :7, 1:, :3], ws.permute(0, 2, 3, 1).to(torch.bfloat16))
    assert ops.stem_frame_hw((800, 1088)) == (806, 1096)
    # window stem: k = r*32 + px*4 + ch (a kernel row's 8 pixels x 4 channels are 64 contiguous bytes of the canvas row)
    sw = ops.untile_k(ops.pack_stem_weight(ws, 256, order="window")).view(64, 8, 8, 4)
    assert sw[:, 7].abs().sum() == 0 and sw[:, :, 0].abs().sum() == 0 and sw[..., 3].abs().sum() == 0
    assert torch.equal(sw[:, :7, 1:, :3], ws.permute(0, 2, 3, 1).to(torch.bfloat16))


def test_window_stem_k_order_is_a_sliding_window_over_canvas_rows():
    """The window stem's GEMM (hn_conv_desc.stem_window): for output pixel (oy, ox) kernel row r reads the 64 contiguous bytes
    of canvas row 2*oy - 3 + r that start at pixel 2*ox - 4 (zero outside the canvas), against pack_stem_weight(order="window").
    Emulated on the CPU with unfold: equals F.conv2d(stride 2, padding 3) (backbone.body.conv1, fcos_utils/fcos.py:476)."""
    import torch.nn.functional as F
    from hn_b200 import ops
    g = torch.Generator().manual_seed(3)
    n, h, w = 2, 10, 14
    canvas = torch.zeros(n, h, w, 4)
    canvas[..., :3] = torch.randn(n, h, w, 3, generator=g)
    wt = torch.randn(8, 3, 7, 7, generator=g)
    ref = F.conv2d(canvas[..., :3].permute(0, 3, 1, 2), wt, stride=2, padding=3)
    wk = ops.untile_k(ops.pack_stem_weight(wt.to(torch.bfloat16).float(), 256, order="window"))[:8].float()      # [cout, 256]
    padded = torch.zeros(n, h + 8, w + 8, 4)                        # 3 rows above (+ the zero-weight 8th row below), 4 px left
    padded[:, 3:3 + h, 4:4 + w] = canvas
    oh, ow = (h + 1) // 2, (w + 1) // 2
    out = torch.zeros(n, 8, oh, ow)
    for oy in range(oh):
        for ox in range(ow):
            win = padded[:, 2 * oy:2 * oy + 8, 2 * ox:2 * ox + 8, :].reshape(n, 256)        # rows 2oy-3.., pixels 2ox-4..
            out[:, :, oy, ox] = win @ wk.t()
    ref_bf = F.conv2d(canvas[..., :3].permute(0, 3, 1, 2), wt.to(torch.bfloat16).float(), stride=2, padding=3)
    torch.testing.assert_close(out, ref_bf, rtol=1e-4, atol=1e-4)
    assert (ref - ref_bf).abs().max() < 0.2


# ------------------------------------------------------------------------------------------------
# weight-change detection and checkpoint loading (handnet_pipeline.py:14-52)
# ------------------------------------------------------------------------------------------------
class _Args:
    pretrained_fcos = ""
    pretrained_a2j = ""


def _quiet_handnet(*a, **k):
    import contextlib
    import io
    from handnet_pipeline.handnet_pipeline import HandNet
    with contextlib.redirect_stdout(io.StringIO()):
        return HandNet(*a, **k)


def test_weights_epoch_changes_on_nested_load_inplace_edit_and_apply():
    """nn.Module.load_state_dict on a PARENT recurses through _load_from_state_dict, so the children's own
    load_state_dict override never runs: the epoch must still change (post hook), and so must it for in-place edits."""
    from hn_b200.runtime import weights_token
    net = _quiet_handnet(_Args(), num_classes=3).eval()
    t0 = weights_token(net)
    assert weights_token(net) == t0                       # stable while nothing changes
    net.load_state_dict(net.state_dict())                 # nested load through the parent
    t1 = weights_token(net)
    assert t1[0] != t0[0] and t1[1] != t0[1]
    with torch.no_grad():
        net.detector.head.classification_head.cls_logits.bias.mul_(1.0)      # in-place edit of one parameter
    t2 = weights_token(net)
    assert t2[0] != t1[0] and t2[1] == t1[1]
    with torch.inference_mode():
        net.a2j.regressionModel.output.bias.add_(0.0)
    t3 = weights_token(net)
    assert t3[1] != t2[1] and t3[0] == t2[0]
    net.double()                                          # _apply reaches the children through the parent
    t4 = weights_token(net)
    assert t4[0] != t3[0] and t4[1] != t3[1]
    net.a2j.load_state_dict(net.a2j.state_dict())         # direct load on the child
    assert weights_token(net)[1] != t4[1]


def test_weights_epoch_of_a_model_built_under_inference_mode():
    """Parameters created inside torch.inference_mode() are inference tensors: they have no version counter (reading
    ``_version`` raises).  bench.py --config hd1080 builds its network that way; the epoch must still work and still change
    on a reload."""
    from hn_b200.runtime import weights_token
    with torch.inference_mode():
        net = _quiet_handnet(_Args(), num_classes=3).eval()
        assert all(p.is_inference() for p in net.parameters())
        t0 = weights_token(net)
        assert weights_token(net) == t0
        net.load_state_dict(net.state_dict())
        assert weights_token(net) != t0


def test_handnet_max_hands_surface():
    """max_hands is an extension of the constructor (default 1 = the reference, handnet_pipeline.py:84-85)."""
    import pytest
    assert _quiet_handnet(_Args(), num_classes=3).max_hands == 1
    assert _quiet_handnet(_Args(), num_classes=3, max_hands=4).max_hands == 4
    with pytest.raises(ValueError):
        _quiet_handnet(_Args(), num_classes=3, max_hands=0)


def test_handnet_reloads_fcos_and_a2j_checkpoints(tmp_path):
    """HandNet(args, reload_detector=True, reload_a2j=True): {"model": state_dict} files, strict=False
    (handnet_pipeline.py:17-19, 36-38)."""
    fsd = synth.fcos_state_dict(3, False, seed=3)
    asd = synth.a2j_state_dict(seed=4)
    args = _Args()
    args.pretrained_fcos = str(tmp_path / "fcos_10.pth")
    args.pretrained_a2j = str(tmp_path / "a2j_25.pth")
    extra = dict(fsd)
    extra["not.in.the.model"] = torch.zeros(1)             # strict=False tolerates foreign keys
    torch.save({"model": extra, "epoch": 10}, args.pretrained_fcos)
    torch.save({"model": asd}, args.pretrained_a2j)
    net = _quiet_handnet(args, reload_detector=True, num_classes=3, reload_a2j=True)
    got = net.detector.state_dict()
    for k, v in fsd.items():
        assert torch.equal(got[k], v), k
    got = net.a2j.state_dict()
    for k, v in asd.items():
        assert torch.equal(got[k], v), k
    assert all(not p.requires_grad for p in net.parameters())
    assert not net.detector.training


@pytest.mark.parametrize("rgbd", [False, True])
def test_handnet_loads_lightning_ckpt(tmp_path, rgbd):
    """A Lightning-style .ckpt ({"state_dict": {"a2j.<key>": ...}, "hyper_parameters": {...}}) through
    load_pretrained_a2j -> A2JModelLightning.load_from_checkpoint (handnet_pipeline.py:28-29, a2j/a2j.py:252-283),
    for the depth-only and the RGBD (4-channel stem) variants."""
    from a2j.a2j import A2JModel, A2JModelLightning
    src = A2JModel(21, 176, 176, is_RGBD=rgbd)
    g = torch.Generator().manual_seed(11)
    sd = {k: (torch.randn(v.shape, generator=g) * 0.05 if v.dtype.is_floating_point else v.clone())
          for k, v in src.state_dict().items()}
    ckpt = {"state_dict": {"a2j." + k: v for k, v in sd.items()},
            "hyper_parameters": {"num_classes": 21, "crop_height": 176, "crop_width": 176, "is_3D": True, "is_RGBD": rgbd,
                                 "spatial_factor": 0.5, "display_freq": 5000, "output_dir": "models/a2j"},
            "epoch": 3, "global_step": 1234, "pytorch-lightning_version": "1.5.10"}
    args = _Args()
    args.pretrained_a2j = str(tmp_path / ("rgbd.ckpt" if rgbd else "depth.ckpt"))
    torch.save(ckpt, args.pretrained_a2j)
    net = _quiet_handnet(args, num_classes=3, reload_a2j=True, RGBD=rgbd)
    assert isinstance(net.a2j, A2JModelLightning) and net.RGBD == rgbd
    pose = net._pose_net()
    assert pose.Backbone.channel_in == (4 if rgbd else 1)
    assert pose.Backbone.model.conv1.weight.shape[1] == (4 if rgbd else 3)
    got = pose.state_dict()
    for k, v in sd.items():
        assert torch.equal(got[k], v), k
    assert not net.a2j.training


REFERENCE = "/root/reference"


@pytest.mark.skipif(not __import__("os").path.isdir(REFERENCE), reason="the reference checkout is only in the build container")
def test_names_the_reference_entry_scripts_import_resolve_in_the_package():
    """Drop-in at import level: every name that ros_demo.py, a2j_infer.py, a2j_mesh.py, trainval_net_fcos.py and
    trainval_net_a2j.py import from handnet_pipeline.*, fcos_utils.{fcos,det_utils,anchor_utils} and a2j.* exists in this
    package (the scripts' other imports -- rospy, cv_bridge, datasets, pose2mesh -- are outside the path).  The call
    sequences themselves run on the GPU in tests/test_gpu_c_models.py."""
    import ast
    import importlib
    import os
    ours = {"handnet_pipeline.handnet_pipeline", "fcos_utils.fcos", "fcos_utils.det_utils", "fcos_utils.anchor_utils",
            "a2j.a2j", "a2j.anchor", "a2j.resnet"}
    seen = []
    for script in ("ros_demo.py", "a2j_infer.py", "a2j_mesh.py", "trainval_net_fcos.py", "trainval_net_a2j.py"):
        path = os.path.join(REFERENCE, script)
        if not os.path.exists(path):
            continue
        tree = ast.parse(open(path).read())
        for node in ast.walk(tree):
            if isinstance(node, ast.ImportFrom) and node.module in ours:
                mod = importlib.import_module(node.module)
                for alias in node.names:
                    assert hasattr(mod, alias.name), f"{script}: from {node.module} import {alias.name}"
                    seen.append((script, node.module, alias.name))
    assert ("ros_demo.py", "handnet_pipeline.handnet_pipeline", "HandNet") in seen
    assert any(m == "fcos_utils.fcos" and n == "FCOS" for _, m, n in seen)

"""GPU parity of the steps either side of the path: frame ingest, RGBD A2J stem, batched convert_joints.

Ingest and the channel pack are byte / exact-rounding work: bit-exact against numpy / torch.  convert_joints is pinned
to the reference's golden vectors (tests/golden/a2j_small.pt) through the oracle and compared bit for bit with it."""
import numpy as np
import pytest
import torch

from oracle import a2j_oracle
from hn_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from hn_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize("n,h,w", [(2, 480, 640), (1, 7, 13), (3, 33, 250)])
def test_ingest_frames_bit_exact(ops, n, h, w):
    """ros_demo.py:266: cv2 BGR uint8 -> RGB.transpose(2,0,1).astype(float32) / 255.0; :230-231 uint16 mm / 1000.0."""
    g = np.random.default_rng(5)
    bgr = g.integers(0, 256, size=(n, h, w, 3), dtype=np.uint8)
    mm = g.integers(0, 65536, size=(n, h, w), dtype=np.uint16)
    mm[0, 0, :4] = (0, 1, 999, 65535)
    rgb_ref = np.stack([im[:, :, ::-1].transpose(2, 0, 1).astype(np.float32) / 255.0 for im in bgr])
    d = mm.astype(np.float32)
    d /= 1000.0
    rgb, depth = ops.ingest_frames(torch.from_numpy(bgr).cuda(), torch.from_numpy(mm.view(np.int16)).cuda())
    assert depth.shape == (n, 1, h, w)
    assert np.array_equal(rgb.cpu().numpy(), rgb_ref)
    assert np.array_equal(depth.cpu().numpy()[:, 0], d)
    # one input only
    rgb2, none = ops.ingest_frames(torch.from_numpy(bgr).cuda(), None)
    assert none is None and torch.equal(rgb2, rgb)


def test_ingest_empty_batch_is_an_error(ops):
    with pytest.raises((RuntimeError, AssertionError)):
        ops.ingest_frames(torch.zeros((0, 4, 4, 3), dtype=torch.uint8, device="cuda"), None)


def test_pack_nhwc4_frame_exact(ops):
    g = torch.Generator().manual_seed(3)
    x = torch.randn(3, 5, 20, 36, generator=g).cuda()
    fr = ops.StemFrame(3, (20, 36), "cuda")
    ops.pack_nhwc4_frame(x, (2, 1, 0, 3), fr)
    ref = x[:, [2, 1, 0, 3]].permute(0, 2, 3, 1).to(torch.bfloat16)
    assert torch.equal(fr.canvas(), ref)
    # the frame around the canvas stays zero
    t = fr.rows().clone()
    t[:, 3:23, 4:40] = 0
    assert t.abs().sum() == 0
    ops.pack_nhwc4_frame(x, (0, -1, 4, -1), fr)          # -1: zero channel
    assert torch.equal(fr.canvas()[..., 0], x[:, 0].to(torch.bfloat16)) and fr.canvas()[..., 1].abs().sum() == 0
    assert torch.equal(fr.canvas()[..., 2], x[:, 4].to(torch.bfloat16))


def test_convert_joints_device_matches_oracle_and_golden(ops, golden):
    fx = golden("a2j_small.pt")
    j = fx["joints"][:1].float().contiguous()
    box = fx["convert_box"].to(torch.int64).reshape(1, 4)
    paras = fx["convert_paras"] if "convert_paras" in fx else None       # float32 intrinsics in the golden fixture
    uv = ops.convert_joints(j.cuda(), box.cuda())
    assert np.array_equal(uv.cpu().numpy()[0], a2j_oracle.convert_joints(j[0].numpy(), box[0].numpy()))
    # the golden vectors were made with a float32 box (float32 arithmetic); the pipeline's crops are int64, for which
    # numpy evaluates in float64: same values to 1 float32 ulp
    torch.testing.assert_close(uv.cpu()[0], fx["convert_uv"].float(), rtol=1e-6, atol=0)
    if paras is not None:
        xyz = ops.convert_joints(j.cuda(), box.cuda(), paras)
        assert np.array_equal(xyz.cpu().numpy()[0], a2j_oracle.convert_joints(j[0].numpy(), box[0].numpy(), paras.numpy()))
        torch.testing.assert_close(xyz.cpu()[0], fx["convert_xyz"].float(), rtol=1e-5, atol=1e-3)
    # a batch with different boxes, against the oracle joint by joint
    g = torch.Generator().manual_seed(9)
    jj = (torch.rand(5, 21, 3, generator=g) * 176).contiguous()
    bb = torch.tensor([[0, 0, 175, 175], [10, 20, 300, 250], [639, 479, 639, 479], [5, 7, 6, 9], [100, 50, 420, 400]])
    pp = torch.tensor([615.0, 614.5, 320.25, 241.75], dtype=torch.float64)
    got = ops.convert_joints(jj.cuda(), bb.cuda(), pp).cpu().numpy()
    for i in range(5):
        assert np.array_equal(got[i], a2j_oracle.convert_joints(jj[i].numpy(), bb[i].numpy(), pp.numpy()))


def test_a2j_rgbd_variant_vs_oracle():
    """is_RGBD=True: 4-channel 7x7 stem (a2j/a2j.py:191-192) through the direct-stem path."""
    from a2j.a2j import A2JModel
    sd = synth.a2j_state_dict(seed=4, channel_in=4)
    m = A2JModel(21, 176, 176, is_RGBD=True).eval()
    m.load_state_dict(sd)
    m.cuda()
    g = torch.Generator().manual_seed(11)
    x = torch.rand(2, 4, 176, 176, generator=g)
    x[:, 3] *= 1.5
    with torch.inference_mode():
        joints = m(x.cuda())
        j_emu = a2j_oracle.a2j_forward(sd, x, emulate_bf16=True, channel_in=4)
        j_ref = a2j_oracle.a2j_forward(sd, x, emulate_bf16=False, channel_in=4)
    assert joints.shape == (2, 21, 3) and joints.device.type == "cpu"
    assert ((joints - j_emu).abs() / j_emu.abs().clamp(min=1.0)).max() < 1e-3      # bf16 path vs bf16-emulating oracle
    assert (joints - j_ref).abs().max() < 5e-2                                        # vs fp32 arithmetic


def test_handnet_forward_frames_equals_forward_on_converted_inputs():
    from handnet_pipeline.handnet_pipeline import HandNet

    class Args:
        pretrained_fcos = ""
        pretrained_a2j = ""
    net = HandNet(Args(), reload_detector=False, num_classes=3, reload_a2j=False).eval()
    net.detector.load_state_dict(synth.fcos_state_dict(3, False, seed=0))
    net.a2j.load_state_dict(synth.a2j_state_dict(seed=1))
    net.cuda()
    g = np.random.default_rng(2)
    bgr = g.integers(0, 256, size=(2, 120, 160, 3), dtype=np.uint8)
    mm = g.integers(300, 1500, size=(2, 120, 160), dtype=np.uint16)
    rgb = [torch.from_numpy(im[:, :, ::-1].transpose(2, 0, 1).astype(np.float32) / 255.0).cuda() for im in bgr]
    depth = torch.from_numpy(mm.astype(np.float32) / np.float32(1000.0))[:, None].cuda()
    with torch.inference_mode():
        a = net(rgb, depth_images=depth)
        b = net.forward_frames(torch.from_numpy(bgr), torch.from_numpy(mm.view(np.int16)))
    for x, y in zip(a, b):
        assert torch.equal(x.cpu(), y.cpu())

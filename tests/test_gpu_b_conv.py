"""GPU parity: tcgen05 shifted-GEMM convolution and the glue kernels around it.

The comparator for a single layer is torch's fp32 convolution (TF32 off) on the same bf16-rounded inputs and
weights: both sides accumulate bf16 x bf16 products in fp32, so they agree to fp32 summation-order noise; the
output is then rounded to bf16 by the kernel (tolerance: 1 bf16 ulp = 2^-8 relative)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import fcos_oracle

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from hn_b200 import ops as _ops
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _ops


def q(x):
    return x.to(torch.bfloat16).float()


def close_bf16(got, ref, what=""):
    """|got - ref| <= 2^-7 * |ref| + small: one bf16 rounding plus fp32 reorder noise."""
    err = (got - ref).abs()
    tol = ref.abs() * 2 ** -7 + 2e-3 * ref.abs().max().clamp(min=1e-3)
    bad = err > tol
    assert not bad.any(), f"{what}: {int(bad.sum())} of {bad.numel()} off, max err {err.max().item():.4e}"


def rand(g, *shape, scale=1.0):
    return q(torch.randn(*shape, generator=g) * scale)


CASES = [
    # name, n, h, w, cin, cout, k, stride, dil, halo, block_n
    ("1x1_64_64_bn64", 2, 9, 13, 64, 64, 1, 1, 1, 1, 64),
    ("3x3_64_64_bn64", 2, 20, 28, 64, 64, 3, 1, 1, 1, 64),
    ("3x3_128_128_bn128", 2, 20, 28, 128, 128, 3, 1, 1, 1, 128),
    ("3x3_256_256_bn256", 2, 25, 34, 256, 256, 3, 1, 1, 1, 256),
    ("3x3_256_256_auto", 3, 50, 68, 256, 256, 3, 1, 1, 1, 0),
    ("1x1_512_2048_auto", 3, 11, 11, 512, 2048, 1, 1, 1, 1, 0),
    ("3x3_2048_256_auto", 2, 11, 11, 2048, 256, 3, 1, 1, 1, 0),
    ("3x3_64_32_bn32", 1, 12, 20, 64, 32, 3, 1, 1, 1, 32),
    ("3x3_256_5_bn16", 2, 25, 34, 256, 5, 3, 1, 1, 1, 16),
    ("3x3_256_336_pad", 2, 11, 11, 256, 336, 3, 1, 1, 1, 0),
    # cout a multiple of 32 but cout_pad > cout: a wholly padded N tile must not be stored (found by tools/conv_fuzz.py)
    ("3x3_256_96_bn32_padded_tile", 2, 30, 43, 256, 96, 3, 1, 1, 1, 32),
    ("3x3_256_96_bn16_padded_tile", 2, 30, 43, 256, 96, 3, 1, 1, 1, 16),
    ("3x3_256_160_auto_padded_tile", 2, 30, 43, 256, 160, 3, 1, 1, 1, 0),
    ("3x3_s2_64_128", 2, 20, 28, 64, 128, 3, 2, 1, 1, 0),
    ("3x3_s2_odd_128_256", 2, 25, 33, 128, 256, 3, 2, 1, 1, 0),
    ("1x1_s2_256_512", 2, 22, 22, 256, 512, 1, 2, 1, 1, 0),
    ("3x3_dil2_512_512", 3, 11, 11, 512, 512, 3, 1, 2, 2, 0),
    ("3x3_halo2_in_halo1_conv", 2, 11, 11, 64, 64, 3, 1, 1, 2, 64),
    ("3x3_256_256_bn256_odd_tiles", 3, 25, 34, 256, 256, 3, 1, 1, 1, 256),
    ("1x1_512_2048_bn256", 3, 11, 11, 512, 2048, 1, 1, 1, 1, 256),
    ("3x3_s2_128_256_bn256", 2, 25, 33, 128, 256, 3, 2, 1, 1, 256),
    ("3x3_256_256_auto_big", 4, 100, 136, 256, 256, 3, 1, 1, 1, 0),
    # resident weights (one narrow N tile, >= 2 tiles per SM): layer1-like 3x3 and stem-like 1x1 GEMM
    ("rb_3x3_64_64_layer1", 2, 200, 272, 64, 64, 3, 1, 1, 1, 0),
    ("rb_1x1_256_64_stem", 1, 400, 544, 256, 64, 1, 1, 1, 0, 0),
    ("rb_3x3_256_5_headout", 4, 100, 136, 256, 5, 3, 1, 1, 1, 0),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_scale_shift_relu(ops, case):
    name, n, h, w, cin, cout, k, stride, dil, halo, bn = case
    g = torch.Generator().manual_seed(sum(map(ord, name)))
    x = rand(g, n, cin, h, w).to(DEV)
    wt = rand(g, cout, cin, k, k, scale=(cin * k * k) ** -0.5).to(DEV)
    scale = (0.5 + torch.rand(cout, generator=g)).to(DEV)
    shift = (0.3 * torch.randn(cout, generator=g)).to(DEV)
    ref = F.conv2d(x, wt, None, stride=stride, padding=(k // 2) * dil, dilation=dil)
    ref = F.relu(ref * scale[None, :, None, None] + shift[None, :, None, None])
    xin = ops.Act.from_nchw(x, halo) if stride == 1 else ops.PhaseAct.from_nchw(x, halo)
    out = ops.Act(n, ref.shape[2], ref.shape[3], cout, 1, DEV)
    ops.conv2d(xin, ops.pack_conv_weight(wt), cout=cout, ksize=k, stride=stride, dilation=dil, scale=scale,
               shift=shift, relu=True, out=out, block_n=bn)
    torch.cuda.synchronize()
    close_bf16(out.to_nchw(), ref, name)
    full = out.t.float().abs().sum()
    assert torch.isclose(full, out.interior().float().abs().sum()), "halo of the output must stay zero"


@pytest.mark.parametrize("shape", [(2, 60, 70, 64, 64, 0), (2, 50, 68, 256, 256, 0), (3, 40, 50, 128, 128, 0), (4, 100, 136, 256, 5, 1)])
def test_conv_epilogue_alternate_tiles(ops, shape):
    """The two epilogue warps of a TMEM lane quarter either split the column chunks of one tile or take alternate tiles
    (default for single-N-tile layers up to 64 columns; debug bit 16 forces it, bit 15 forbids it): same values."""
    n, h, w, cin, cout, rows_out = shape
    g = torch.Generator().manual_seed(h + w + cout)
    x = rand(g, n, cin, h, w).to(DEV)
    wt = rand(g, cout, cin, 3, 3, scale=(cin * 9) ** -0.5).to(DEV)
    scale = (0.5 + torch.rand(cout, generator=g)).to(DEV)
    shift = (0.3 * torch.randn(cout, generator=g)).to(DEV)
    ref = F.conv2d(x, wt, None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None]
    outs = []
    for debug in (65536, 32768):
        if rows_out:
            buf = torch.zeros(n, h * w, 8, device=DEV)
            ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=cout, ksize=3, scale=scale, shift=shift,
                       out_f32=buf, out_rows_per_image=h * w, out_row_offset=0, debug=debug)
            torch.cuda.synchronize()
            got = buf[..., :cout].reshape(n, h, w, cout).permute(0, 3, 1, 2)
            torch.testing.assert_close(got, ref, rtol=2e-3, atol=2e-3)
            outs.append(buf.clone())
        else:
            idn = rand(g, n, cout, h, w).to(DEV) if debug == 65536 else idn
            out = ops.Act(n, h, w, cout, 1, DEV)
            ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=cout, ksize=3, scale=scale, shift=shift, relu=True,
                       res=ops.Act.from_nchw(idn, 1), res_mode=1, out=out, debug=debug)
            torch.cuda.synchronize()
            close_bf16(out.to_nchw(), F.relu(ref + idn), f"alternate tiles debug={debug}")
            outs.append(out.t.clone())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize("shape", [(3, 100, 136, 128, 128), (3, 100, 137, 128, 128), (4, 90, 120, 64, 128), (8, 100, 136, 128, 128)],
                         ids=["even_tiles", "odd_tiles", "one_chunk_odd", "layer2_vga_b8"])
def test_conv_duo_items_share_weight_boxes(ops, shape):
    """3x3 stride-1 layers of one <= 128-column N tile with >= 2 tiles per SM take work items of TWO M tiles behind one weight
    box per k-step (debug bit 18 forbids it): same accumulation order per tile, so the outputs are bit-identical."""
    n, h, w, cin, cout = shape
    g = torch.Generator().manual_seed(h * w + cin)
    x = rand(g, n, cin, h, w).to(DEV)
    wt = rand(g, cout, cin, 3, 3, scale=(cin * 9) ** -0.5).to(DEV)
    scale = (0.5 + torch.rand(cout, generator=g)).to(DEV)
    shift = (0.3 * torch.randn(cout, generator=g)).to(DEV)
    idn = rand(g, n, cout, h, w).to(DEV)
    ref = F.relu(F.conv2d(x, wt, None, padding=1) * scale[None, :, None, None] + shift[None, :, None, None] + idn)
    outs = []
    for debug in (0, 262144):
        out = ops.Act(n, h, w, cout, 1, DEV)
        ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=cout, ksize=3, scale=scale, shift=shift, relu=True,
                   res=ops.Act.from_nchw(idn, 1), res_mode=1, out=out, debug=debug)
        torch.cuda.synchronize()
        close_bf16(out.to_nchw(), ref, f"duo debug={debug}")
        assert torch.isclose(out.t.float().abs().sum(), out.interior().float().abs().sum()), "halo of the output must stay zero"
        outs.append(out.t.clone())
    assert torch.equal(outs[0], outs[1])


def test_conv_residual_phase_copy_and_geometry_remap(ops):
    """BasicBlock tail: conv + scale/shift + identity + ReLU, written twice (plain halo-2 and phase-split)."""
    g = torch.Generator().manual_seed(7)
    n, h, w, c = 2, 26, 18, 128
    x = rand(g, n, c, h, w).to(DEV)
    idn = rand(g, n, c, h, w).to(DEV)
    wt = rand(g, c, c, 3, 3, scale=(c * 9) ** -0.5).to(DEV)
    scale = (0.5 + torch.rand(c, generator=g)).to(DEV)
    shift = (0.3 * torch.randn(c, generator=g)).to(DEV)
    ref = F.relu(F.conv2d(x, wt, padding=1) * scale[None, :, None, None] + shift[None, :, None, None] + idn)
    out = ops.Act(n, h, w, c, 2, DEV)
    ph = ops.PhaseAct(n, h, w, c, 1, DEV)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=c, ksize=3, scale=scale, shift=shift,
               relu=True, res=ops.Act.from_nchw(idn, 3), res_mode=1, out=out, out_phase=ph)
    torch.cuda.synchronize()
    close_bf16(out.to_nchw(), ref, "plain")
    want = ops.PhaseAct.from_nchw(out.to_nchw(), 1)
    assert torch.equal(ph.t, want.t), "phase-split copy must hold the same bf16 values"


def test_conv_fpn_lateral_upsample_add(ops):
    g = torch.Generator().manual_seed(8)
    n, h, w = 2, 50, 68
    x = rand(g, n, 256, h, w).to(DEV)
    top = rand(g, n, 256, h // 2, w // 2).to(DEV)
    wt = rand(g, 256, 256, 1, 1, scale=1 / 16).to(DEV)
    bias = (0.1 * torch.randn(256, generator=g)).to(DEV)
    ref = F.conv2d(x, wt, bias) + F.interpolate(top, size=(h, w), mode="nearest")
    out = ops.Act(n, h, w, 256, 1, DEV)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=256, ksize=1, shift=bias,
               res=ops.Act.from_nchw(top, 1), res_mode=2, out=out)
    torch.cuda.synchronize()
    close_bf16(out.to_nchw(), ref, "fpn lateral")


def test_conv_fp32_rows_output_fused_heads(ops):
    """FCOS output convs: [cls(3) | lr(2)] fused, ReLU on a channel sub-range, fp32 rows at a level offset."""
    g = torch.Generator().manual_seed(9)
    n, h, w = 2, 25, 34
    x = rand(g, n, 256, h, w).to(DEV)
    wt = rand(g, 5, 256, 3, 3, scale=0.02).to(DEV)
    bias = torch.randn(5, generator=g).to(DEV)
    ref = F.conv2d(x, wt, bias, padding=1)
    ref[:, 3:5] = F.relu(ref[:, 3:5])
    ref_rows = ref.permute(0, 2, 3, 1).reshape(n, h * w, 5)
    buf = torch.full((n, 1000 + h * w, 8), -7.0, device=DEV)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=5, ksize=3, shift=bias, relu=(3, 5),
               out_f32=buf, out_rows_per_image=1000 + h * w, out_row_offset=1000)
    torch.cuda.synchronize()
    torch.testing.assert_close(buf[:, 1000:, :5], ref_rows, rtol=1e-4, atol=1e-4)
    assert (buf[:, :1000] == -7.0).all() and (buf[:, 1000:, 5:] == -7.0).all(), "nothing else may be written"


def test_conv_fp32_channel_planes_output(ops):
    """out_kind 2: the fused FCOS output convolutions write fp32 channel planes [n][planes][rows] at a level's row offset
    (what hn_fcos_decode_select streams); same values as the row layout, nothing else written."""
    g = torch.Generator().manual_seed(19)
    n, h, w = 2, 25, 34
    x = rand(g, n, 256, h, w).to(DEV)
    wt = rand(g, 5, 256, 3, 3, scale=0.02).to(DEV)
    bias = torch.randn(5, generator=g).to(DEV)
    ref = F.conv2d(x, wt, bias, padding=1)
    ref[:, 0:4] = F.relu(ref[:, 0:4])
    rows_total = 300 + h * w + 7
    planes = torch.full((n, 8, rows_total), -7.0, device=DEV)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=5, ksize=3, shift=bias, relu=(0, 4),
               out_f32=planes, out_rows_per_image=rows_total, out_row_offset=300, out_planar=True)
    rowbuf = torch.zeros((n, h * w, 8), device=DEV)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=5, ksize=3, shift=bias, relu=(0, 4), out_f32=rowbuf)
    torch.cuda.synchronize()
    got = planes[:, :5, 300:300 + h * w]
    torch.testing.assert_close(got, ref.reshape(n, 5, h * w), rtol=1e-4, atol=1e-4)
    assert torch.equal(got, rowbuf[..., :5].permute(0, 2, 1)), "planes and rows hold the same fp32 values"
    assert (planes[:, 5:] == -7.0).all() and (planes[:, :, :300] == -7.0).all() and (planes[:, :, 300 + h * w:] == -7.0).all()


def test_conv_fp32_rows_transposed_a2j_layout(ops):
    """A2J output convs: rows ordered w-major (permute(0,3,2,1), a2j/a2j.py:85-89)."""
    g = torch.Generator().manual_seed(10)
    n, h, w, cout = 3, 11, 11, 336
    x = rand(g, n, 256, h, w).to(DEV)
    wt = rand(g, cout, 256, 3, 3, scale=0.02).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    ref = F.conv2d(x, wt, bias, padding=1).permute(0, 3, 2, 1).reshape(n, w * h, cout)
    buf = torch.zeros((n, w * h, cout), device=DEV)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=cout, ksize=3, shift=bias, out_f32=buf,
               out_transpose_hw=True)
    torch.cuda.synchronize()
    torch.testing.assert_close(buf, ref, rtol=1e-4, atol=1e-4)


def test_conv_groupnorm_tower_layer(ops):
    """conv3x3 + bias -> GroupNorm(32) -> ReLU (fcos_utils/fcos.py:232-240): stats from the conv epilogue."""
    g = torch.Generator().manual_seed(11)
    n, h, w, c = 3, 25, 34, 256
    x = rand(g, n, c, h, w).to(DEV)
    wt = rand(g, c, c, 3, 3, scale=0.02).to(DEV)
    bias = (0.05 * torch.randn(c, generator=g)).to(DEV)
    gamma = (0.7 + 0.6 * torch.rand(c, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(c, generator=g)).to(DEV)
    raw = q(F.conv2d(x, wt, bias, padding=1))
    ref = F.relu(F.group_norm(raw, 32, gamma, beta, eps=1e-5))
    out = ops.Act(n, h, w, c, 1, DEV)
    stats = torch.zeros(n, 32, 2, dtype=torch.int64, device=DEV)          # 40.24 fixed-point sums
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=c, ksize=3, shift=bias, out=out,
               gn_stats=stats, gn_groups=32)
    torch.cuda.synchronize()
    raw_got = out.to_nchw()
    close_bf16(raw_got, raw, "raw conv")
    want = torch.stack((raw_got.double().reshape(n, 32, -1).sum(-1), (raw_got.double() ** 2).reshape(n, 32, -1).sum(-1)), -1)
    torch.testing.assert_close(stats.double() / ops.GN_FIX_SCALE, want, rtol=1e-5, atol=1e-3)
    # integer accumulation: a second launch produces exactly the same sums (fp atomics did not)
    stats2 = torch.zeros_like(stats)
    ops.conv2d(ops.Act.from_nchw(x, 1), ops.pack_conv_weight(wt), cout=c, ksize=3, shift=bias, out=ops.Act(n, h, w, c, 1, DEV),
               gn_stats=stats2, gn_groups=32)
    torch.cuda.synchronize()
    assert torch.equal(stats, stats2), "GroupNorm statistics must be bit-reproducible"
    ops.groupnorm_relu(out, stats, 32, gamma, beta, 1e-5)
    torch.cuda.synchronize()
    close_bf16(out.to_nchw(), F.relu(F.group_norm(raw_got, 32, gamma, beta, eps=1e-5)), "gn apply")
    close_bf16(out.to_nchw(), ref, "gn vs reference chain")
    assert torch.isclose(out.t.float().abs().sum(), out.interior().float().abs().sum()), "halo must stay zero"


def test_preprocess_matches_transform(ops):
    g = torch.Generator().manual_seed(12)
    imgs = [torch.rand(3, 120, 160, generator=g), torch.rand(3, 97, 131, generator=g)]
    canvas, sizes = fcos_oracle.transform(imgs, 256, 448)
    got = ops.preprocess([i.cuda() for i in imgs], sizes, tuple(canvas.shape[-2:]), fcos_oracle.IMAGE_MEAN,
                         fcos_oracle.IMAGE_STD)
    torch.cuda.synchronize()
    got = got.float().cpu()
    assert (got[..., 3] == 0).all()
    ref = canvas.permute(0, 2, 3, 1)
    err = (got[..., :3] - ref).abs()
    assert (err <= ref.abs() * 2 ** -8 + 1e-5).all(), err.max()
    for i, (oh, ow) in enumerate(sizes):
        assert got[i, oh:].abs().sum() == 0 and got[i, :, ow:].abs().sum() == 0, "padding must be zero"


def test_stem_im2col_gemm_and_maxpool(ops):
    """backbone.body.conv1 + FrozenBN + ReLU + maxpool as im2col -> tcgen05 GEMM -> pool."""
    g = torch.Generator().manual_seed(13)
    n, h, w = 2, 64, 96
    canvas = torch.zeros(n, h, w, 4)
    canvas[..., :3] = rand(g, n, h, w, 3)
    wt = rand(g, 64, 3, 7, 7, scale=0.08)
    scale = 0.5 + torch.rand(64, generator=g)
    shift = 0.2 * torch.randn(64, generator=g)
    x_nchw = canvas[..., :3].permute(0, 3, 1, 2).contiguous()
    conv = q(F.relu(F.conv2d(x_nchw, wt, stride=2, padding=3) * scale[None, :, None, None] + shift[None, :, None, None]))
    ref = F.max_pool2d(conv, 3, 2, 1)
    a, oh, ow = ops.im2col_7x7s2(canvas.to(torch.bfloat16).cuda(), 256)
    stem = ops.Act(n, oh, ow, 64, 0, DEV)
    ops.conv2d(ops.Act(n, oh, ow, 256, 0, DEV, t=a.view(n, oh, ow, 256)), ops.pack_stem_weight(wt.cuda(), 256),
               cout=64, ksize=1, scale=scale.cuda(), shift=shift.cuda(), relu=True, out=stem)
    pooled = ops.Act(n, (oh + 1) // 2, (ow + 1) // 2, 64, 1, DEV)
    ops.maxpool3x3s2(stem.t, pooled)
    torch.cuda.synchronize()
    close_bf16(stem.to_nchw().cpu(), conv, "stem conv")
    close_bf16(pooled.to_nchw().cpu(), ref, "maxpool")


@pytest.mark.parametrize("n,h,w", [(2, 64, 96), (1, 96, 320), (3, 32, 544)])
def test_stem_direct_from_framed_canvas(ops, n, h, w):
    """backbone.body.conv1 + FrozenBN + ReLU straight from the zero-framed canvas (overlapping-stride TMA patches, no
    im2col buffer) == F.conv2d(stride 2, padding 3); also == the im2col + GEMM path bit for bit (same K order)."""
    g = torch.Generator().manual_seed(15)
    canvas = torch.zeros(n, h, w, 4)
    canvas[..., :3] = rand(g, n, h, w, 3)
    wt = rand(g, 64, 3, 7, 7, scale=0.08)
    scale = 0.5 + torch.rand(64, generator=g)
    shift = 0.2 * torch.randn(64, generator=g)
    x_nchw = canvas[..., :3].permute(0, 3, 1, 2).contiguous()
    conv = q(F.relu(F.conv2d(x_nchw, wt, stride=2, padding=3) * scale[None, :, None, None] + shift[None, :, None, None]))
    frame = ops.StemFrame(n, (h, w), DEV)
    frame.set_canvas(canvas.to(torch.bfloat16))
    wp = ops.pack_stem_weight(wt.cuda(), 256)
    stem = ops.Act(n, h // 2, w // 2, 64, 0, DEV)
    ops.conv2d(frame, wp, cout=64, ksize=1, scale=scale.cuda(), shift=shift.cuda(), relu=True, out=stem)
    a, oh, ow = ops.im2col_7x7s2(canvas.to(torch.bfloat16).cuda(), 256)
    stem2 = ops.Act(n, oh, ow, 64, 0, DEV)
    ops.conv2d(ops.Act(n, oh, ow, 256, 0, DEV, t=a.view(n, oh, ow, 256)), wp, cout=64, ksize=1, scale=scale.cuda(),
               shift=shift.cuda(), relu=True, out=stem2)
    torch.cuda.synchronize()
    close_bf16(stem.to_nchw().cpu(), conv, "direct stem conv")
    assert torch.equal(stem.t, stem2.t), "direct stem and im2col + GEMM must agree exactly"


@pytest.mark.parametrize("n,h,w", [(2, 64, 96), (1, 96, 320), (3, 32, 544), (2, 50, 250), (1, 800, 1088)])
def test_stem_window_from_plain_canvas(ops, n, h, w):
    """backbone.body.conv1 + FrozenBN + ReLU straight from the PLAIN row-major canvas: the im2col matrix of a kernel row is
    the canvas row itself, read through an un-swizzled UMMA descriptor whose rows start 16 bytes apart (stem_window); one
    TMA box of 8 rows x 256 pixels per tile of 125 output pixels, borders zero-filled by TMA.  == F.conv2d(stride 2,
    padding 3) and == the row-pair frame stem / im2col + GEMM path bit for bit (same products, same fp32 accumulation
    order inside the tensor core is not guaranteed across K orders: compared to bf16 rounding)."""
    g = torch.Generator().manual_seed(16)
    canvas = torch.zeros(n, h, w, 4)
    canvas[..., :3] = rand(g, n, h, w, 3)
    wt = rand(g, 64, 3, 7, 7, scale=0.08)
    scale = 0.5 + torch.rand(64, generator=g)
    shift = 0.2 * torch.randn(64, generator=g)
    x_nchw = canvas[..., :3].permute(0, 3, 1, 2).contiguous()
    conv = q(F.relu(F.conv2d(x_nchw, wt, stride=2, padding=3) * scale[None, :, None, None] + shift[None, :, None, None]))
    cv = ops.StemCanvas(n, (h, w), DEV).set_canvas(canvas.to(torch.bfloat16).cuda())
    stem = ops.Act(n, cv.oh, cv.ow, 64, 0, DEV)
    stem.t.fill_(-7.0)
    ops.conv2d(cv, ops.pack_stem_weight(wt.cuda(), 256, order="window"), cout=64, ksize=1, scale=scale.cuda(), shift=shift.cuda(),
               relu=True, out=stem)
    torch.cuda.synchronize()
    close_bf16(stem.to_nchw().cpu(), conv, "window stem conv")
    if h % 2 == 0:
        frame = ops.StemFrame(n, (h, w), DEV)
        frame.set_canvas(canvas.to(torch.bfloat16))
        stem2 = ops.Act(n, h // 2, w // 2, 64, 0, DEV)
        ops.conv2d(frame, ops.pack_stem_weight(wt.cuda(), 256), cout=64, ksize=1, scale=scale.cuda(), shift=shift.cuda(), relu=True,
                   out=stem2)
        torch.cuda.synchronize()
        close_bf16(stem.to_nchw().cpu(), stem2.to_nchw().cpu(), "window stem vs row-pair stem")


def test_stem_fp32_depth_one_channel(ops):
    """A2J stem: one depth channel expanded to three == weights summed over Cin (a2j/a2j.py:197-199)."""
    g = torch.Generator().manual_seed(14)
    n = 3
    depth = torch.rand(n, 176, 176, generator=g) * 1.5
    wt = torch.randn(64, 3, 7, 7, generator=g) * 0.05
    wsum = q(wt.sum(1, keepdim=True))
    ref = F.conv2d(q(depth)[:, None], wsum, stride=2, padding=3)
    a, oh, ow = ops.im2col_7x7s2(depth.cuda(), 64)
    out = ops.Act(n, oh, ow, 64, 0, DEV)
    ops.conv2d(ops.Act(n, oh, ow, 64, 0, DEV, t=a.view(n, oh, ow, 64)), ops.pack_stem_weight(wsum.cuda(), 64),
               cout=64, ksize=1, out=out)
    torch.cuda.synchronize()
    close_bf16(out.to_nchw().cpu(), ref, "a2j stem")


@pytest.mark.parametrize("shape", [(8, 11, 11, 2048, 256, 3, 0), (8, 11, 11, 1024, 256, 1, 4), (3, 22, 22, 128, 128, 3, 3),
                                   (8, 11, 11, 256, 336, 3, 0)])
def test_conv_split_k(ops, shape):
    """Split-K: every K split stores its partial tile into its own fp32 slice, the last-arriving work item adds the
    slices in split order (deterministic) and runs the epilogue (scale/shift, residual, ReLU); counters are left zero,
    the scratch needs no initialisation."""
    n, h, w, cin, cout, k, splits = shape
    g = torch.Generator().manual_seed(cin + cout)
    x = rand(g, n, cin, h, w).to(DEV)
    idn = rand(g, n, cout, h, w).to(DEV)
    wt = rand(g, cout, cin, k, k, scale=(cin * k * k) ** -0.5).to(DEV)
    scale = (0.5 + torch.rand(cout, generator=g)).to(DEV)
    shift = (0.3 * torch.randn(cout, generator=g)).to(DEV)
    ref = F.relu(F.conv2d(x, wt, padding=k // 2) * scale[None, :, None, None] + shift[None, :, None, None] + idn)
    wp = ops.pack_conv_weight(wt)
    rows = n * (h + 2) * (w + 2)
    m_tiles = (rows + 127) // 128
    ws = torch.full((8 * m_tiles * 128 * wp.shape[0],), float("nan"), dtype=torch.float32, device=DEV)   # 8 slices of junk
    cnt = torch.zeros(m_tiles * (wp.shape[0] // 16), dtype=torch.int32, device=DEV)
    xin, res = ops.Act.from_nchw(x, 1), ops.Act.from_nchw(idn, 1)
    outs = []
    for rep in range(3):                       # later launches: counters must have been left clean
        out = ops.Act(n, h, w, cout, 1, DEV)
        ops.conv2d(xin, wp, cout=cout, ksize=k, scale=scale, shift=shift, relu=True, res=res, res_mode=1, out=out,
                   splitk=(ws, cnt), splits=splits)
        torch.cuda.synchronize()
        close_bf16(out.to_nchw(), ref, f"split-k rep {rep}")
        assert cnt.abs().max() == 0
        outs.append(out.t.clone())
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2]), "split-K must be deterministic"


LEVEL_SETS = [
    # batch, levels (h, w)
    (2, [(20, 28), (10, 14), (5, 7)]),
    (3, [(100, 136), (50, 68), (25, 34)]),
    (1, [(96, 168), (48, 84)]),
]


@pytest.mark.parametrize("n,levels", LEVEL_SETS, ids=["small", "vga_b3", "two_levels"])
def test_conv_levels_tower_layer_equals_per_level_launches(ops, n, levels):
    """hn_conv2d_bf16_levels: ONE launch over P3+P4+P5 with shared weights (the FCOS tower loop over levels,
    fcos_utils/fcos.py:278-289) == one launch per level, bit for bit: bf16 outputs, zero halos and the fixed-point
    GroupNorm sums; the fused GroupNorm-apply launch likewise."""
    g = torch.Generator().manual_seed(21 + n)
    c = 256
    wt = rand(g, c, c, 3, 3, scale=0.02).to(DEV)
    bias = (0.05 * torch.randn(c, generator=g)).to(DEV)
    gamma = (0.7 + 0.6 * torch.rand(c, generator=g)).to(DEV)
    beta = (0.1 * torch.randn(c, generator=g)).to(DEV)
    wp = ops.pack_conv_weight(wt)
    xs = [rand(g, n, c, h, w).to(DEV) for h, w in levels]
    acts = [ops.Act.from_nchw(x, 1) for x in xs]
    # per level
    outs_a = [ops.Act(n, h, w, c, 1, DEV) for h, w in levels]
    st_a = [torch.zeros(n, 32, 2, dtype=torch.int64, device=DEV) for _ in levels]
    for a, o, st in zip(acts, outs_a, st_a):
        ops.conv2d(a, wp, cout=c, ksize=3, shift=bias, out=o, gn_stats=st, gn_groups=32)
    # fused
    outs_b = [ops.Act(n, h, w, c, 1, DEV) for h, w in levels]
    st_b = [torch.zeros(n, 32, 2, dtype=torch.int64, device=DEV) for _ in levels]
    ops.conv2d_levels(acts, wp, cout=c, ksize=3, shift=bias, outs=outs_b, gn_stats=st_b, gn_groups=32)
    torch.cuda.synchronize()
    for lvl, (x, oa, ob) in enumerate(zip(xs, outs_a, outs_b)):
        close_bf16(ob.to_nchw(), q(F.conv2d(x, wt, bias, padding=1)), f"level {lvl} vs F.conv2d")
        assert torch.equal(oa.t, ob.t), f"level {lvl}: fused launch differs from the per-level launch"
        assert torch.equal(st_a[lvl], st_b[lvl]), f"level {lvl}: GroupNorm sums differ"
    for o, st in zip(outs_a, st_a):
        ops.groupnorm_relu(o, st, 32, gamma, beta, 1e-5)
    ops.groupnorm_relu_levels(outs_b, st_b, 32, gamma, beta, 1e-5)
    torch.cuda.synchronize()
    for lvl, (oa, ob) in enumerate(zip(outs_a, outs_b)):
        assert torch.equal(oa.t, ob.t), f"level {lvl}: fused GroupNorm apply differs"
        assert torch.isclose(ob.t.float().abs().sum(), ob.interior().float().abs().sum()), "halo must stay zero"


@pytest.mark.parametrize("n,levels", LEVEL_SETS, ids=["small", "vga_b3", "two_levels"])
@pytest.mark.parametrize("cout,relu", [(5, (3, 5)), (13, (10, 13))], ids=["cls_lr", "ext_heads"])
def test_conv_levels_head_outputs_equal_per_level_launches(ops, n, levels, cout, relu):
    """The fused FCOS output convolutions ([cls | lr | ...], ReLU on a channel sub-range) over all levels in one launch:
    fp32 rows of every level at its row offset of the shared [n, locs, ld] buffer, nothing else written."""
    g = torch.Generator().manual_seed(33 + n + cout)
    wt = rand(g, cout, 256, 3, 3, scale=0.02).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    wp = ops.pack_conv_weight(wt)
    xs = [rand(g, n, 256, h, w).to(DEV) for h, w in levels]
    acts = [ops.Act.from_nchw(x, 1) for x in xs]
    starts = [0]
    for h, w in levels:
        starts.append(starts[-1] + h * w)
    locs, ld = starts[-1], 16
    buf_a = torch.full((n, locs + 3, ld), -7.0, device=DEV)
    buf_b = torch.full((n, locs + 3, ld), -7.0, device=DEV)
    for a, off in zip(acts, starts):
        ops.conv2d(a, wp, cout=cout, ksize=3, shift=bias, relu=relu, out_f32=buf_a, out_rows_per_image=locs + 3, out_row_offset=off)
    ops.conv2d_levels(acts, wp, cout=cout, ksize=3, shift=bias, relu=relu, out_f32=buf_b, out_rows_per_image=locs + 3,
                      out_row_offsets=starts[:-1])
    torch.cuda.synchronize()
    # (per-level launches of the small levels stream their weights and accumulate kernel row by kernel row, the fused launch
    # keeps them resident and accumulates chunk by chunk: fp32 rows agree to summation-order noise, not bit for bit)
    torch.testing.assert_close(buf_a, buf_b, rtol=1e-5, atol=1e-5)
    assert (buf_b[:, locs:] == -7.0).all() and (buf_b[..., cout:] == -7.0).all(), "nothing else may be written"
    # the product's layout: channel planes [n][planes][locs] in one launch over all levels
    planes = torch.full((n, ld, locs + 3), -7.0, device=DEV)
    ops.conv2d_levels(acts, wp, cout=cout, ksize=3, shift=bias, relu=relu, out_f32=planes, out_rows_per_image=locs + 3,
                      out_row_offsets=starts[:-1], out_planar=True)
    torch.cuda.synchronize()
    assert torch.equal(planes[:, :cout, :locs], buf_b[:, :locs, :cout].permute(0, 2, 1))
    assert (planes[:, cout:] == -7.0).all() and (planes[:, :, locs:] == -7.0).all()
    for lvl, ((h, w), x) in enumerate(zip(levels, xs)):
        ref = F.conv2d(x, wt, bias, padding=1)
        ref[:, relu[0]:relu[1]] = F.relu(ref[:, relu[0]:relu[1]])
        torch.testing.assert_close(buf_b[:, starts[lvl]:starts[lvl + 1], :cout], ref.permute(0, 2, 3, 1).reshape(n, h * w, cout),
                                   rtol=1e-4, atol=1e-4)


def test_conv_levels_rejects_mismatched_levels(ops):
    a = ops.Act(1, 8, 8, 256, 1, DEV)
    b = ops.Act(1, 4, 4, 256, 1, DEV)
    w1 = ops.pack_conv_weight(torch.randn(256, 256, 3, 3, device=DEV) * 0.02)
    with pytest.raises(RuntimeError, match="levels"):
        # four levels: more than the kernel's three segments
        ops.conv2d_levels([a, b, b, b], w1, cout=256, ksize=3, outs=[ops.Act(1, 8, 8, 256, 1, DEV)] + [ops.Act(1, 4, 4, 256, 1, DEV)] * 3)
    # a 64-wide layer has no multi-level instantiation: it runs level by level inside the call
    w64 = ops.pack_conv_weight(torch.randn(64, 256, 3, 3, device=DEV) * 0.02)
    o = [ops.Act(1, 8, 8, 64, 1, DEV), ops.Act(1, 4, 4, 64, 1, DEV)]
    ops.conv2d_levels([a, b], w64, cout=64, ksize=3, outs=o)
    o2 = [ops.Act(1, 8, 8, 64, 1, DEV), ops.Act(1, 4, 4, 64, 1, DEV)]
    for x, y in zip([a, b], o2):
        ops.conv2d(x, w64, cout=64, ksize=3, out=y)
    torch.cuda.synchronize()
    assert torch.equal(o[0].t, o2[0].t) and torch.equal(o[1].t, o2[1].t)


"""Torch-tensor front end of the C ABI: one thin function per entry point of include/handnet_b200.h.

PyTorch is used for device memory and streams only; all arithmetic happens in libhandnet_b200.so.
Every function raises RuntimeError when the library reports an error -- there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import ConvDesc, check, ptr, stream_ptr

BF16 = torch.bfloat16
GN_FIX_SCALE = float(1 << 24)      # hn_conv_desc.gn_stats holds (sum, sum of squares) as value * 2^24 int64
_ENV_DEBUG = int(__import__('os').environ.get('HN_CONV_DEBUG', '0'))   # bring-up experiments only
PROFILE = None      # set to a list by runtime.conv_profile: (start event, end event, algorithmic FLOPs) per conv launch


def _require_cuda(t: torch.Tensor, name: str):
    if not t.is_cuda:
        raise RuntimeError(f"{name} must live on a CUDA device (libhandnet_b200 has no CPU path)")


# ------------------------------------------------------------------------------------------------
# activations
# ------------------------------------------------------------------------------------------------
class Act:
    """Haloed NHWC bf16 activation: t[n, h + 2*halo, w + 2*halo, c]; the halo is zero and stays zero."""

    __slots__ = ("t", "n", "h", "w", "c", "halo")

    def __init__(self, n: int, h: int, w: int, c: int, halo: int, device="cuda", t: Optional[torch.Tensor] = None):
        self.n, self.h, self.w, self.c, self.halo = n, h, w, c, halo
        shape = (n, h + 2 * halo, w + 2 * halo, c)
        self.t = torch.zeros(shape, dtype=BF16, device=device) if t is None else t
        assert tuple(self.t.shape) == shape and self.t.dtype == BF16 and self.t.is_contiguous()

    def interior(self) -> torch.Tensor:
        p = self.halo
        return self.t[:, p:p + self.h, p:p + self.w, :]

    @staticmethod
    def from_nchw(x: torch.Tensor, halo: int) -> "Act":
        n, c, h, w = x.shape
        a = Act(n, h, w, c, halo, x.device)
        a.interior().copy_(x.permute(0, 2, 3, 1))
        return a

    def to_nchw(self) -> torch.Tensor:
        return self.interior().permute(0, 3, 1, 2).float().contiguous()


class StemFrame:
    """Zero-framed 4-channel bf16 canvas of the direct stem: logically [n, hc+6, wc+8, 4] with the canvas at row 3,
    column 4; stored with its rows in PAIRS, t[n, (hc+6)/2, wc+8, 2, 4], so that two kernel rows of the 7x7 window over
    8 pixels are one contiguous 128-byte run (hn_conv2d_bf16, stem_pitch_*)."""

    __slots__ = ("t", "n", "hc", "wc", "fh", "fw", "oh", "ow")

    def __init__(self, n: int, canvas_hw: Tuple[int, int], device="cuda"):
        self.n, (self.hc, self.wc) = n, canvas_hw
        self.fh, self.fw = canvas_hw[0] + 6, canvas_hw[1] + 8
        assert self.fh % 2 == 0, "the stem frame stores its rows in pairs: canvas height must be even"
        self.oh, self.ow = canvas_hw[0] // 2, canvas_hw[1] // 2
        self.t = torch.zeros((n, self.fh // 2, self.fw, 2, 4), dtype=BF16, device=device)

    def rows(self) -> torch.Tensor:
        """Row-major COPY of the whole frame [n, fh, fw, 4]."""
        return self.t.permute(0, 1, 3, 2, 4).reshape(self.n, self.fh, self.fw, 4)

    def canvas(self) -> torch.Tensor:
        """Row-major COPY of the canvas rectangle [n, hc, wc, 4] (the storage is row-pair interleaved)."""
        return self.rows()[:, 3:3 + self.hc, 4:4 + self.wc, :]

    def set_canvas(self, x: torch.Tensor) -> "StemFrame":
        """Write a row-major [n, hc, wc, 4] canvas into the frame (tests / tools; the kernels write it directly)."""
        full = torch.zeros((self.n, self.fh, self.fw, 4), dtype=BF16, device=self.t.device)
        full[:, 3:3 + self.hc, 4:4 + self.wc, :] = x.to(BF16)
        self.t.copy_(full.view(self.n, self.fh // 2, 2, self.fw, 4).permute(0, 1, 3, 2, 4))
        return self


class StemCanvas:
    """Plain row-major 4-channel bf16 canvas [n, hc, wc, 4] of the WINDOW stem (hn_conv_desc.stem_window): no frame, the
    convolution's TMA loads zero-fill the borders."""

    __slots__ = ("t", "n", "hc", "wc", "oh", "ow")

    def __init__(self, n: int, canvas_hw: Tuple[int, int], device="cuda"):
        self.n, (self.hc, self.wc) = n, canvas_hw
        assert self.wc % 2 == 0, "the window stem needs an even canvas width (16-byte row pitch)"
        self.oh, self.ow = (canvas_hw[0] + 1) // 2, (canvas_hw[1] + 1) // 2
        self.t = torch.zeros((n, self.hc, self.wc, 4), dtype=BF16, device=device)

    def canvas(self) -> torch.Tensor:
        return self.t

    def set_canvas(self, x: torch.Tensor) -> "StemCanvas":
        self.t.copy_(x.to(BF16))
        return self


class PhaseAct:
    """Phase-split haloed NHWC bf16: t[4, n, h2 + 2*halo, w2 + 2*halo, c], phase = (y & 1) * 2 + (x & 1),
    (h2, w2) = ceil((h, w) / 2).  Input format of the stride-2 convolutions."""

    __slots__ = ("t", "n", "h", "w", "h2", "w2", "c", "halo")

    def __init__(self, n: int, h: int, w: int, c: int, halo: int, device="cuda"):
        self.n, self.h, self.w, self.c, self.halo = n, h, w, c, halo
        self.h2, self.w2 = (h + 1) // 2, (w + 1) // 2
        self.t = torch.zeros((4, n, self.h2 + 2 * halo, self.w2 + 2 * halo, c), dtype=BF16, device=device)

    @staticmethod
    def from_nchw(x: torch.Tensor, halo: int) -> "PhaseAct":
        n, c, h, w = x.shape
        a = PhaseAct(n, h, w, c, halo, x.device)
        xh = x.permute(0, 2, 3, 1)
        for py in range(2):
            for px in range(2):
                sub = xh[:, py::2, px::2, :]
                a.t[py * 2 + px, :, halo:halo + sub.shape[1], halo:halo + sub.shape[2], :] = sub
        return a


def pad_cout(cout: int) -> int:
    if cout <= 16:
        return 16
    if cout <= 32:
        return 32
    return (cout + 63) // 64 * 64


def tile_k(m: torch.Tensor) -> torch.Tensor:
    """[cout_pad][K] -> the kernel's weight layout [K/64][cout_pad][64] (k-block major), returned with the logical
    shape [cout_pad, K]: the tile of one k-block and BN output channels is then ONE contiguous run of BN*128 bytes,
    which is what the TMA box of hn_conv2d_bf16 fetches (DRAM- and L2-friendly; a row-major [cout][K] matrix would
    scatter the box over BN rows K*2 bytes apart)."""
    cp, k = m.shape
    assert k % 64 == 0
    return m.view(cp, k // 64, 64).permute(1, 0, 2).contiguous().view(cp, k)


def untile_k(m: torch.Tensor) -> torch.Tensor:
    """Inverse of tile_k (tests, debugging)."""
    cp, k = m.shape
    return m.view(k // 64, cp, 64).permute(1, 0, 2).contiguous().view(cp, k)


def pack_conv_weight(w: torch.Tensor, scale: Optional[torch.Tensor] = None) -> torch.Tensor:
    """OIHW fp32 -> bf16 weights for hn_conv2d_bf16: K = kh*kw*cin in tap-major order (channels fastest), output
    channels padded with zero rows to cout_pad, stored k-block major (tile_k).  Logical shape [cout_pad, K]."""
    cout, cin, kh, kw = w.shape
    m = w.detach().float().permute(0, 2, 3, 1).reshape(cout, kh * kw * cin)
    out = torch.zeros((pad_cout(cout), m.shape[1]), dtype=BF16, device=w.device)
    out[:cout] = m.to(BF16)
    return tile_k(out)


def pack_stem_weight(w: torch.Tensor, k_pad: int, order: str = "pairs") -> torch.Tensor:
    """7x7 stem OIHW (cin = 3, 4 or 1) -> bf16 [cout_pad][k_pad] in the K order of hn_im2col_7x7s2 / the direct stem:
    C = 1 (depth): k = r*8 + px;  C = 4 (RGB or RGBD canvas): order "pairs" (row-pair frame, StemFrame): k = j*64 + px*8 +
    rr*4 + ch with kernel row r = 2j + rr; order "window" (plain canvas, StemCanvas): k = r*32 + px*4 + ch.
    px = 0, r = 7 and (for RGB) ch = 3 carry zero weights."""
    cout, cin, kh, kw = w.shape
    assert (kh, kw) == (7, 7) and (cin, k_pad) in ((3, 256), (4, 256), (1, 64)) and order in ("pairs", "window")
    c = 4 if cin >= 3 else 1
    m = torch.zeros((cout, 8, 8, c), dtype=torch.float32, device=w.device)
    m[:, :7, 1:, :cin] = w.detach().float().permute(0, 2, 3, 1)
    if c == 4 and order == "pairs":
        # 4-channel canvas: a k-block is two kernel rows interleaved per pixel, k = j*64 + px*8 + rr*4 + ch (r = 2j + rr),
        # the order in which the row-pair frame holds them contiguously
        m = m.view(cout, 4, 2, 8, c).permute(0, 1, 3, 2, 4)
    out = torch.zeros((pad_cout(cout), k_pad), dtype=BF16, device=w.device)
    out[:cout] = m.reshape(cout, -1).to(BF16)
    return tile_k(out)


# ------------------------------------------------------------------------------------------------
# entry points
# ------------------------------------------------------------------------------------------------
def device_info() -> Tuple[int, int, int]:
    sm, ma, mi = C.c_int(), C.c_int(), C.c_int()
    check(_lib.load().hn_device_info(C.byref(sm), C.byref(ma), C.byref(mi)), "hn_device_info")
    return sm.value, ma.value, mi.value


def conv_cta_cap(max_ctas: int):
    """Upper bound on the CTAs of the following convolution launches (0 = all SMs)."""
    check(_lib.load().hn_conv_set_cta_cap(int(max_ctas)), "hn_conv_set_cta_cap")


def conv_pdl(enabled: bool):
    """Programmatic dependent launch for the following convolution launches (default on)."""
    check(_lib.load().hn_conv_set_pdl(int(bool(enabled))), "hn_conv_set_pdl")


def launch_count() -> int:
    return int(_lib.load().hn_launch_count())


STEM_PAD_TOP, STEM_PAD_LEFT = 3, 4       # position of the canvas inside the zero frame the direct stem reads


def stem_frame_hw(canvas_hw: Tuple[int, int]) -> Tuple[int, int]:
    """Frame (pitch) of the zero-padded canvas for hn_conv2d_bf16's direct 7x7/2 stem: 3 rows above / below (the
    eighth, zero-weight kernel row included) and 4 columns left / right."""
    return canvas_hw[0] + 6, canvas_hw[1] + 8


def preprocess(images: Sequence[torch.Tensor], out_sizes: Sequence[Tuple[int, int]], canvas_hw: Tuple[int, int],
               mean: Sequence[float], std: Sequence[float], canvas: Optional[torch.Tensor] = None,
               frame: Optional[torch.Tensor] = None) -> torch.Tensor:
    """T1.  images: list of fp32 [3,H,W] CUDA tensors -> bf16 canvas [B, Hc, Wc, 4]; or, with `frame` (bf16
    StemFrame.t: rows in pairs, [B, (Hc+6)/2, Wc+8, 2, 4], zero outside the canvas), into the canvas rectangle at (3, 4)
    of the frame."""
    b = len(images)
    imgs = []
    for im in images:
        _require_cuda(im, "image")
        if im.dtype != torch.float32 or im.dim() != 3 or im.shape[0] != 3:
            raise RuntimeError("images must be float32 [3, H, W]")
        imgs.append(im.contiguous())
    dev = imgs[0].device
    if canvas is None and frame is None:
        canvas = torch.empty((b, canvas_hw[0], canvas_hw[1], 4), dtype=BF16, device=dev)
    ptrs = (C.c_void_p * b)(*[t.data_ptr() for t in imgs])
    ih = (C.c_int * b)(*[int(t.shape[1]) for t in imgs])
    iw = (C.c_int * b)(*[int(t.shape[2]) for t in imgs])
    oh = (C.c_int * b)(*[int(s[0]) for s in out_sizes])
    ow = (C.c_int * b)(*[int(s[1]) for s in out_sizes])
    m3 = (C.c_float * 3)(*[float(v) for v in mean])
    s3 = (C.c_float * 3)(*[float(v) for v in std])
    if frame is not None:
        fh, fw = stem_frame_hw(canvas_hw)
        assert frame.dtype == BF16 and tuple(frame.shape) == (b, fh // 2, fw, 2, 4) and frame.is_contiguous()
        check(_lib.load().hn_preprocess_resize_pad_framed(ptrs, ih, iw, oh, ow, b, m3, s3, frame.data_ptr(), canvas_hw[0],
                                                          canvas_hw[1], STEM_PAD_TOP, STEM_PAD_LEFT, fh, fw, stream_ptr()),
              "hn_preprocess_resize_pad_framed")
        return frame
    check(_lib.load().hn_preprocess_resize_pad(ptrs, ih, iw, oh, ow, b, m3, s3, canvas.data_ptr(), canvas_hw[0],
                                               canvas_hw[1], stream_ptr()), "hn_preprocess_resize_pad")
    return canvas


def ingest_frames(bgr_u8: Optional[torch.Tensor], depth_u16: Optional[torch.Tensor], rgb_out: Optional[torch.Tensor] = None,
                  depth_out: Optional[torch.Tensor] = None):
    """Camera frames -> network inputs.  bgr_u8: uint8 [n,h,w,3] (cv2 BGR) -> fp32 [n,3,h,w] RGB in 0..1;
    depth_u16: uint16 [n,h,w] millimetres (or int16 storage of the same bits) -> fp32 [n,1,h,w] metres."""
    ref = bgr_u8 if bgr_u8 is not None else depth_u16
    _require_cuda(ref, "frames")
    n, h, w = int(ref.shape[0]), int(ref.shape[1]), int(ref.shape[2])
    if bgr_u8 is not None:
        assert bgr_u8.dtype == torch.uint8 and tuple(bgr_u8.shape) == (n, h, w, 3) and bgr_u8.is_contiguous()
        if rgb_out is None:
            rgb_out = torch.empty((n, 3, h, w), dtype=torch.float32, device=ref.device)
        assert rgb_out.dtype == torch.float32 and tuple(rgb_out.shape) == (n, 3, h, w) and rgb_out.is_contiguous()
    if depth_u16 is not None:
        assert depth_u16.dtype in (torch.uint16, torch.int16) and tuple(depth_u16.shape) == (n, h, w) and depth_u16.is_contiguous()
        if depth_out is None:
            depth_out = torch.empty((n, 1, h, w), dtype=torch.float32, device=ref.device)
        assert depth_out.dtype == torch.float32 and depth_out.numel() == n * h * w and depth_out.is_contiguous()
    check(_lib.load().hn_ingest_frames(ptr(bgr_u8), ptr(depth_u16), n, h, w, ptr(rgb_out if bgr_u8 is not None else None),
                                       ptr(depth_out if depth_u16 is not None else None), stream_ptr()), "hn_ingest_frames")
    return rgb_out, depth_out


def pack_nhwc4_frame(src: torch.Tensor, chan_map: Sequence[int], frame: "StemFrame") -> "StemFrame":
    """fp32 [n,c,h,w] -> the canvas rectangle of a StemFrame, frame channel j = src channel chan_map[j] (-1: zero)."""
    _require_cuda(src, "src")
    n, c, h, w = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous() and (frame.n, frame.hc, frame.wc) == (n, h, w)
    cm = (C.c_int * 4)(*[int(v) for v in chan_map])
    check(_lib.load().hn_pack_nhwc4_frame(src.data_ptr(), n, c, h, w, cm, frame.t.data_ptr(), STEM_PAD_TOP, STEM_PAD_LEFT,
                                          frame.fh, frame.fw, stream_ptr()), "hn_pack_nhwc4_frame")
    return frame


def convert_joints(uvd: torch.Tensor, crops: torch.Tensor, paras: Optional[torch.Tensor] = None, crop_w: int = 176,
                   crop_h: int = 176, has_hand: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Batched a2j.convert_joints on the device.  uvd fp32 [n,J,3], crops int64 [n,4], paras (fx,fy,cx,cy) as a
    float64 or float32 tensor (numpy evaluates the back-projection in the intrinsics' precision)."""
    _require_cuda(uvd, "uvd")
    n, j, _ = uvd.shape
    assert uvd.dtype == torch.float32 and uvd.is_contiguous() and crops.dtype == torch.int64 and tuple(crops.shape) == (n, 4)
    crops = crops.contiguous()
    f64 = 1
    if paras is not None:
        f64 = 0 if paras.dtype == torch.float32 else 1
        paras = paras.to(device=uvd.device, dtype=torch.float64 if f64 else torch.float32).contiguous()
        assert paras.numel() == 4
    out = torch.empty_like(uvd)
    check(_lib.load().hn_convert_joints(uvd.data_ptr(), crops.data_ptr(), ptr(has_hand), ptr(paras), f64, n, j, crop_w,
                                        crop_h, out.data_ptr(), stream_ptr()), "hn_convert_joints")
    return out


def im2col_7x7s2(x: torch.Tensor, k_pad: int, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, int, int]:
    """x: bf16 [n,h,w,4] canvas (k_pad 256) or fp32 [n,h,w] depth (k_pad 64) -> bf16 [n*oh*ow, k_pad]."""
    _require_cuda(x, "x")
    is_f32 = x.dtype == torch.float32
    if is_f32:
        n, h, w = x.shape
        c = 1
    else:
        n, h, w, c4 = x.shape
        assert c4 == 4 and x.dtype == BF16
        c = 4
    oh, ow = (h + 1) // 2, (w + 1) // 2
    if out is None:
        out = torch.empty((n * oh * ow, k_pad), dtype=BF16, device=x.device)
    check(_lib.load().hn_im2col_7x7s2(x.data_ptr(), int(is_f32), n, h, w, c, out.data_ptr(), k_pad, stream_ptr()),
          "hn_im2col_7x7s2")
    return out, oh, ow


def _conv_desc(x, weight: torch.Tensor, *, cout: int, ksize: int, stride: int = 1, dilation: int = 1,
               scale: Optional[torch.Tensor] = None, shift: Optional[torch.Tensor] = None, relu=False,
               res: Optional[Act] = None, res_mode: int = 0, out: Optional[Act] = None,
               out_f32: Optional[torch.Tensor] = None, out_rows_per_image: int = 0, out_row_offset: int = 0,
               out_transpose_hw: bool = False, out_planar: bool = False, out_phase: Optional[PhaseAct] = None,
               gn_stats: Optional[torch.Tensor] = None, gn_groups: int = 0, block_n: int = 0, cluster: int = 0,
               debug: int = 0, splitk=None, splits: int = 0, trace: Optional[torch.Tensor] = None) -> ConvDesc:
    """Fill a struct hn_conv_desc.  x: Act (stride 1), PhaseAct (stride 2) or StemFrame (direct 7x7/2 stem; ksize = 1, the
    weight from pack_stem_weight).  relu: bool or (lo, hi) channel range."""
    d = ConvDesc()
    if isinstance(x, StemFrame):
        assert stride == 1 and ksize == 1
        d.in_, d.n, d.h, d.w, d.cin, d.halo_in, d.in_phases = x.t.data_ptr(), x.n, x.oh, x.ow, 256, 0, 1
        d.stem_pitch_h, d.stem_pitch_w = x.fh, x.fw
    elif isinstance(x, StemCanvas):
        assert stride == 1 and ksize == 1
        d.in_, d.n, d.h, d.w, d.cin, d.halo_in, d.in_phases = x.t.data_ptr(), x.n, x.oh, x.ow, 256, 0, 1
        d.stem_pitch_h, d.stem_pitch_w, d.stem_window = x.hc, x.wc, 1
    elif isinstance(x, PhaseAct):
        assert stride == 2
        d.in_, d.n, d.h, d.w, d.cin, d.halo_in, d.in_phases = x.t.data_ptr(), x.n, x.h2, x.w2, x.c, x.halo, 4
    else:
        assert stride == 1
        d.in_, d.n, d.h, d.w, d.cin, d.halo_in, d.in_phases = x.t.data_ptr(), x.n, x.h, x.w, x.c, x.halo, 1
    assert weight.dtype == BF16 and weight.is_contiguous() and weight.shape[1] == ksize * ksize * d.cin, \
        (weight.shape, ksize, d.cin)
    d.weight, d.cout, d.cout_pad = weight.data_ptr(), cout, weight.shape[0]
    d.kh = d.kw = ksize
    d.stride, d.dilation = stride, dilation
    d.scale, d.shift = ptr(scale), ptr(shift)
    if relu is True:
        d.relu_lo, d.relu_hi = 0, cout
    elif relu:
        d.relu_lo, d.relu_hi = relu
    if res is not None:
        d.res, d.res_mode, d.res_h, d.res_w, d.res_halo = res.t.data_ptr(), res_mode or 1, res.h, res.w, res.halo
        assert res.c == cout
    if out_f32 is not None:
        assert out_f32.dtype == torch.float32 and out_f32.is_contiguous()
        d.out, d.out_kind = out_f32.data_ptr(), 2 if out_planar else 1
        d.out_rows_per_image = out_rows_per_image or d.h * d.w
        # rows [.., rows, ld] (kind 1) or channel planes [.., planes, rows] (kind 2: ld = planes per image)
        d.out_row_offset, d.out_ld, d.out_transpose_hw = out_row_offset, out_f32.shape[-2 if out_planar else -1], int(out_transpose_hw)
    else:
        assert out is not None and out.c == cout and (out.n, out.h, out.w) == (d.n, d.h, d.w), "output geometry"
        d.out, d.out_kind, d.out_halo = out.t.data_ptr(), 0, out.halo
    if out_phase is not None:
        assert (out_phase.n, out_phase.h, out_phase.w, out_phase.c) == (d.n, d.h, d.w, cout)
        d.out_phase, d.out_phase_halo = out_phase.t.data_ptr(), out_phase.halo
    if gn_stats is not None:
        assert gn_stats.dtype == torch.int64 and gn_stats.numel() == d.n * gn_groups * 2
        d.gn_stats, d.gn_groups = gn_stats.data_ptr(), gn_groups
    d.block_n = block_n
    d.cluster = cluster
    d.debug = debug or _ENV_DEBUG
    if trace is not None:
        assert trace.dtype == torch.int64 and trace.numel() >= 3 * 2048 * 2
        d.trace = trace.data_ptr()
    if splitk is not None:            # (fp32 scratch, uint32 counters): both zero, exclusive to this conv
        ws_t, cnt_t = splitk
        d.splitk_ws, d.splitk_ws_bytes = ws_t.data_ptr(), ws_t.numel() * ws_t.element_size()
        d.splitk_counters, d.splitk_counters_len = cnt_t.data_ptr(), cnt_t.numel()
        d.splits = splits
    return d


def _conv_flops(d: ConvDesc, cout: int, k_elems: int) -> float:
    return 2.0 * d.n * d.h * d.w * cout * k_elems


def conv2d(x, weight: torch.Tensor, *, cout: int, ksize: int, algo_k: int = 0, **kw):
    """hn_conv2d_bf16 (see _conv_desc for the arguments).  Returns the output Act (or the fp32 row buffer)."""
    d = _conv_desc(x, weight, cout=cout, ksize=ksize, **kw)
    out, out_f32 = kw.get("out"), kw.get("out_f32")
    if PROFILE is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        check(_lib.load().hn_conv2d_bf16(C.byref(d), stream_ptr()), "hn_conv2d_bf16")
        ev1.record()
        PROFILE.append((ev0, ev1, _conv_flops(d, cout, algo_k or weight.shape[1]),
                        dict(n=d.n, h=d.h, w=d.w, cin=d.cin, cout=cout, k=ksize, stride=d.stride, dil=d.dilation,
                             halo=d.halo_in, f32=out_f32 is not None, gn=kw.get("gn_stats") is not None, levels=1)))
        return out if out_f32 is None else out_f32
    check(_lib.load().hn_conv2d_bf16(C.byref(d), stream_ptr()), "hn_conv2d_bf16")
    return out if out_f32 is None else out_f32


def conv2d_levels(xs: Sequence[Act], weight: torch.Tensor, *, cout: int, ksize: int, outs: Optional[Sequence[Act]] = None,
                  out_f32: Optional[torch.Tensor] = None, out_rows_per_image: int = 0,
                  out_row_offsets: Optional[Sequence[int]] = None, gn_stats: Optional[Sequence[torch.Tensor]] = None,
                  gn_groups: int = 0, **kw):
    """hn_conv2d_bf16_levels: the same 3x3 convolution over several pyramid levels (shared weight / scale / shift) in ONE
    launch.  xs / outs / gn_stats / out_row_offsets are per level."""
    n = len(xs)
    arr = (ConvDesc * n)()
    for i, x in enumerate(xs):
        d = _conv_desc(x, weight, cout=cout, ksize=ksize, out=None if outs is None else outs[i], out_f32=out_f32,
                       out_rows_per_image=out_rows_per_image, out_row_offset=0 if out_row_offsets is None else out_row_offsets[i],
                       gn_stats=None if gn_stats is None else gn_stats[i], gn_groups=gn_groups, **kw)
        C.memmove(C.addressof(arr[i]), C.addressof(d), C.sizeof(ConvDesc))
    prof = PROFILE
    if prof is not None:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
    check(_lib.load().hn_conv2d_bf16_levels(arr, n, stream_ptr()), "hn_conv2d_bf16_levels")
    if prof is not None:
        ev1.record()
        d0 = arr[0]
        prof.append((ev0, ev1, sum(_conv_flops(arr[i], cout, weight.shape[1]) for i in range(n)),
                     dict(n=d0.n, h=d0.h, w=d0.w, cin=d0.cin, cout=cout, k=ksize, stride=1, dil=d0.dilation, halo=d0.halo_in,
                          f32=out_f32 is not None, gn=gn_stats is not None, levels=n)))
    return outs if out_f32 is None else out_f32


def maxpool3x3s2(x: torch.Tensor, out: Act) -> Act:
    """x: bf16 [n,h,w,c] (no halo) -> out Act [n, ceil(h/2), ceil(w/2), c]."""
    n, h, w, c = x.shape
    assert x.dtype == BF16 and x.is_contiguous() and (out.h, out.w, out.c) == ((h + 1) // 2, (w + 1) // 2, c)
    check(_lib.load().hn_maxpool3x3s2(x.data_ptr(), n, h, w, c, out.t.data_ptr(), out.halo, stream_ptr()),
          "hn_maxpool3x3s2")
    return out


def groupnorm_relu(x: Act, stats: torch.Tensor, groups: int, gamma: torch.Tensor, beta: torch.Tensor,
                   eps: float = 1e-5) -> Act:
    check(_lib.load().hn_groupnorm_relu(x.t.data_ptr(), x.n, x.h, x.w, x.c, x.halo, stats.data_ptr(), groups,
                                        gamma.data_ptr(), beta.data_ptr(), eps, stream_ptr()), "hn_groupnorm_relu")
    return x


def groupnorm_relu_levels(xs: Sequence[Act], stats: Sequence[torch.Tensor], groups: int, gamma: torch.Tensor,
                          beta: torch.Tensor, eps: float = 1e-5):
    """GroupNorm + ReLU in place over several pyramid levels in one launch."""
    n = len(xs)
    xp = (C.c_void_p * n)(*[x.t.data_ptr() for x in xs])
    sp = (C.c_void_p * n)(*[t.data_ptr() for t in stats])
    ia = C.c_int * n
    x0 = xs[0]
    assert all((x.c, x.halo) == (x0.c, x0.halo) for x in xs)
    check(_lib.load().hn_groupnorm_relu_levels(xp, ia(*[x.n for x in xs]), ia(*[x.h for x in xs]), ia(*[x.w for x in xs]), n,
                                               x0.c, x0.halo, sp, groups, gamma.data_ptr(), beta.data_ptr(), eps, stream_ptr()),
          "hn_groupnorm_relu_levels")
    return xs


class Levels:
    """Pyramid description shared by decode and gather."""

    def __init__(self, grids: Sequence[Tuple[int, int]], canvas_hw: Tuple[int, int], anchor_sizes: Sequence[int]):
        self.grids = [tuple(int(v) for v in g) for g in grids]
        self.n = len(self.grids)
        self.locs = sum(h * w for h, w in self.grids)
        starts = [0]
        for h, w in self.grids:
            starts.append(starts[-1] + h * w)
        self.starts = starts
        arr = C.c_int * self.n
        self.h = arr(*[g[0] for g in self.grids])
        self.w = arr(*[g[1] for g in self.grids])
        self.sh = arr(*[canvas_hw[0] // g[0] for g in self.grids])      # anchor_utils.py:118-124
        self.sw = arr(*[canvas_hw[1] // g[1] for g in self.grids])
        self.anchor = arr(*[int(s) for s in anchor_sizes])
        self.start_arr = (C.c_int * (self.n + 1))(*starts)


def head_planes(t: torch.Tensor) -> torch.Tensor:
    """A head tensor [B, locs, k] re-laid as fp32 channel planes [B][k][pitch] (pitch = locs rounded up to 32, so that every
    plane starts 128-byte aligned), returned as the strided view [B, locs, k]: the layout the detector's output convolutions
    write (hn_conv_desc.out_kind 2) and hn_fcos_decode_select streams with 16-byte loads."""
    b, locs, k = t.shape
    buf = torch.zeros((b, k, (locs + 31) // 32 * 32), dtype=torch.float32, device=t.device)
    buf[:, :, :locs] = t.permute(0, 2, 1)
    return buf[:, :, :locs].permute(0, 2, 1)


def _head_strides(t: Optional[torch.Tensor]):
    """(pointer, image stride, location stride, channel stride) of a head tensor view [B, locs, k] in any layout."""
    if t is None:
        return 0, 0, 0, 0
    assert t.dtype == torch.float32 and t.dim() == 3
    return t.data_ptr(), t.stride(0), t.stride(1), t.stride(2) if t.shape[2] > 1 else 1


def fcos_decode_select(cls: torch.Tensor, ctr: torch.Tensor, reg: torch.Tensor, num_classes: int, levels: Levels,
                       score_thresh: float, ws: Optional[torch.Tensor] = None):
    """P1-P4.  cls / ctr / reg are fp32 views [B, locs, k] of any layout (channel planes or rows): strides are passed on."""
    b, locs = cls.shape[0], cls.shape[1]
    assert locs == levels.locs
    dev = cls.device
    need = int(_lib.load().hn_fcos_select_workspace_bytes(b, locs))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=dev)
    cand = {
        "count": torch.empty(b, dtype=torch.int32, device=dev),
        "loc": torch.empty((b, locs), dtype=torch.int32, device=dev),
        "score": torch.empty((b, locs), dtype=torch.float32, device=dev),
        "label": torch.empty((b, locs), dtype=torch.int32, device=dev),
        "box": torch.empty((b, locs, 4), dtype=torch.float32, device=dev),
    }
    pc, ic, lc, cc = _head_strides(cls)
    pt, it, lt, _ = _head_strides(ctr)
    pr, ir, lr, cr = _head_strides(reg)
    check(_lib.load().hn_fcos_decode_select(
        pc, ic, lc, cc, pt, it, lt, pr, ir, lr, cr, b, locs,
        num_classes, levels.n, levels.h, levels.w, levels.sh, levels.sw, levels.anchor, float(score_thresh),
        cand["count"].data_ptr(), cand["loc"].data_ptr(), cand["score"].data_ptr(), cand["label"].data_ptr(),
        cand["box"].data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()), "hn_fcos_decode_select")
    return cand


def nms_workspace(batch: int, cap: int, device) -> torch.Tensor:
    return torch.empty(int(_lib.load().hn_nms_workspace_bytes(batch, cap)), dtype=torch.uint8, device=device)


def nms_batched(box: torch.Tensor, score: torch.Tensor, label: torch.Tensor, count: torch.Tensor,
                iou_thresh: float, coord_trick_max_numel: int = 4000, ws: Optional[torch.Tensor] = None):
    """P5.  Returns (keep [B, cap] int32 indices into the candidate list, keep_count [B] int32)."""
    b, cap = score.shape
    dev = score.device
    if ws is None:
        ws = nms_workspace(b, cap, dev)
    keep = torch.empty((b, cap), dtype=torch.int32, device=dev)
    keep_count = torch.empty(b, dtype=torch.int32, device=dev)
    check(_lib.load().hn_nms_batched(box.data_ptr(), score.data_ptr(), label.data_ptr(), count.data_ptr(), b, cap,
                                     float(iou_thresh), int(coord_trick_max_numel), keep.data_ptr(),
                                     keep_count.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()), "hn_nms_batched")
    return keep, keep_count


def fcos_gather(keep, keep_count, cand, hand_lr: torch.Tensor, levels: Levels, ratios_h: Sequence[float],
                ratios_w: Sequence[float], contact: Optional[torch.Tensor] = None,
                dxdy: Optional[torch.Tensor] = None):
    """P6.  Dense per-image outputs with capacity cap = locs."""
    b, cap = keep.shape
    dev = keep.device
    out = {
        "boxes": torch.empty((b, cap, 4), dtype=torch.float32, device=dev),
        "scores": torch.empty((b, cap), dtype=torch.float32, device=dev),
        "labels": torch.empty((b, cap), dtype=torch.int64, device=dev),
        "sides": torch.empty((b, cap), dtype=torch.int64, device=dev),
        "level": torch.empty((b, cap), dtype=torch.float32, device=dev),
    }
    if contact is not None:
        out["contacts"] = torch.empty((b, cap), dtype=torch.int64, device=dev)
        out["dxdymags"] = torch.empty((b, cap, 3), dtype=torch.float32, device=dev)
    rh = (C.c_float * b)(*[float(v) for v in ratios_h])
    rw = (C.c_float * b)(*[float(v) for v in ratios_w])
    pl, il, ll, cl = _head_strides(hand_lr)
    pk, ik, lk, ck = _head_strides(contact)
    pd, id_, ld_, cd = _head_strides(dxdy)
    check(_lib.load().hn_fcos_gather(
        keep.data_ptr(), keep_count.data_ptr(), cand["loc"].data_ptr(), cand["score"].data_ptr(),
        cand["label"].data_ptr(), cand["box"].data_ptr(), pl, il, ll, cl, pk, ik, lk, ck, pd, id_, ld_, cd,
        b, cap, levels.locs, levels.n, levels.start_arr, rh, rw,
        out["boxes"].data_ptr(), out["scores"].data_ptr(), out["labels"].data_ptr(), out["sides"].data_ptr(),
        out["level"].data_ptr(), ptr(out.get("contacts")), ptr(out.get("dxdymags")), stream_ptr()), "hn_fcos_gather")
    return out


def select_crop_resize(boxes: torch.Tensor, labels: torch.Tensor, keep_count: torch.Tensor, hand_label: int,
                       depth: torch.Tensor, out_size: int = 176, out=None, hands: int = 1):
    """S1 + S2.  depth: fp32 [B, C, H, W].  Returns (crops [B*hands,4] int64, has_hand [B*hands] int32, depth_batch
    [B*hands,C,out,out]); slot i*hands + h is the h-th kept hand detection of frame i (hands = 1: the reference's
    boxes[:1]).  `out` supplies those three tensors (the pipeline's hand-off buffers) instead of allocating them."""
    b, cap = labels.shape
    _, dc, ih, iw = depth.shape
    assert depth.dtype == torch.float32 and depth.is_contiguous() and depth.shape[0] == b and hands >= 1
    dev = depth.device
    n = b * hands
    if out is None:
        crops = torch.empty((n, 4), dtype=torch.int64, device=dev)
        has = torch.empty(n, dtype=torch.int32, device=dev)
        db = torch.empty((n, dc, out_size, out_size), dtype=torch.float32, device=dev)
    else:
        crops, has, db = out
        assert crops.dtype == torch.int64 and tuple(crops.shape) == (n, 4) and crops.is_contiguous()
        assert has.dtype == torch.int32 and has.numel() == n and has.is_contiguous()
        assert db.dtype == torch.float32 and tuple(db.shape) == (n, dc, out_size, out_size) and db.is_contiguous()
    check(_lib.load().hn_select_crop_resize_multi(boxes.data_ptr(), labels.data_ptr(), keep_count.data_ptr(), b, cap,
                                                  hand_label, hands, depth.data_ptr(), dc, ih, iw, out_size,
                                                  crops.data_ptr(), has.data_ptr(), db.data_ptr(), stream_ptr()),
          "hn_select_crop_resize_multi")
    return crops, has, db


def a2j_aggregate(cls: torch.Tensor, reg: torch.Tensor, dep: torch.Tensor, anchors: torch.Tensor,
                  ws: Optional[torch.Tensor] = None) -> torch.Tensor:
    """J4.  cls [n,A,J], reg [n,A,J,2], dep [n,A,J] fp32 contiguous; anchors [A,2] fp32 -> [n,J,3]."""
    n, a, j = cls.shape
    for t in (cls, reg, dep, anchors):
        _require_cuda(t, "a2j head")
        assert t.dtype == torch.float32 and t.is_contiguous()
    need = int(_lib.load().hn_a2j_workspace_bytes(n, j))
    if ws is None or ws.numel() < need:
        ws = torch.empty(need, dtype=torch.uint8, device=cls.device)
    out = torch.empty((n, j, 3), dtype=torch.float32, device=cls.device)
    check(_lib.load().hn_a2j_aggregate(cls.data_ptr(), reg.data_ptr(), dep.data_ptr(), anchors.data_ptr(), n, a, j,
                                       out.data_ptr(), ws.data_ptr(), ws.numel(), stream_ptr()), "hn_a2j_aggregate")
    return out


# ------------------------------------------------------------------------------------------------
# pose2mesh lifting: Chebyshev graph convolution, fp32 linear layers, mesh block residual
# ------------------------------------------------------------------------------------------------
class CsrGraph:
    """A (rescaled) graph Laplacian on the device in CSR form.  Accepts a scipy sparse matrix, a torch sparse tensor or a
    dense tensor (what pose2mesh's graph_utils.build_coarse_graphs / sparse_python_to_torch hand over)."""

    __slots__ = ("n", "row_ptr", "col", "val")

    def __init__(self, lap, device="cuda"):
        if hasattr(lap, "tocsr"):                              # scipy
            m = lap.tocsr()
            m.sort_indices()
            rp, ci, vv = torch.from_numpy(m.indptr.copy()), torch.from_numpy(m.indices.copy()), torch.from_numpy(m.data.copy())
            self.n = int(m.shape[0])
        else:
            t = lap.detach().cpu()
            t = (t if t.is_sparse else t.to_sparse()).coalesce()              # COO, sorted by (row, col)
            self.n = int(t.shape[0])
            ri, ci, vv = t.indices()[0], t.indices()[1], t.values()
            rp = torch.zeros(self.n + 1, dtype=torch.int64)
            rp[1:] = torch.cumsum(torch.bincount(ri, minlength=self.n), 0)
        self.row_ptr = rp.to(torch.int32).to(device).contiguous()
        self.col = ci.to(torch.int32).to(device).contiguous()
        self.val = vv.to(torch.float32).to(device).contiguous()


def cheby_spmm(g: CsrGraph, x: torch.Tensor, z: Optional[torch.Tensor] = None, alpha: float = 1.0, beta: float = 0.0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out = alpha * L x + beta * z for x, z [B, V, F] fp32 (cheby_graph_conv.py:26-31)."""
    _require_cuda(x, "x")
    b, v, f = x.shape
    assert v == g.n and x.dtype == torch.float32 and x.is_contiguous()
    if z is not None:
        assert z.shape == x.shape and z.dtype == torch.float32 and z.is_contiguous()
    if out is None:
        out = torch.empty_like(x)
    check(_lib.load().hn_cheby_spmm(g.row_ptr.data_ptr(), g.col.data_ptr(), g.val.data_ptr(), x.data_ptr(), ptr(z), float(alpha),
                                    float(beta), b, v, f, out.data_ptr(), stream_ptr()), "hn_cheby_spmm")
    return out


def linear_f32(planes: Sequence[torch.Tensor], weight: torch.Tensor, bias: Optional[torch.Tensor] = None, *,
               in_affine=None, out_affine=None, relu: bool = False, res: Optional[torch.Tensor] = None,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """y = post(pre([planes interleaved]) @ weight.T + bias) (+ res); planes: 1..3 tensors [M, fin] whose features are
    interleaved plane-fastest (kk = f * len(planes) + p), see hn_linear_f32 in the header."""
    a0 = planes[0]
    _require_cuda(a0, "input")
    m, fin = a0.shape
    n = weight.shape[0]
    assert weight.shape[1] == fin * len(planes) and weight.dtype == torch.float32 and weight.is_contiguous()
    for t in planes:
        assert t.shape == a0.shape and t.dtype == torch.float32 and t.is_contiguous()
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32, device=a0.device)
    isc, ish = in_affine if in_affine is not None else (None, None)
    osc, osh = out_affine if out_affine is not None else (None, None)
    pp = [t.data_ptr() for t in planes] + [0, 0]
    check(_lib.load().hn_linear_f32(pp[0], pp[1], pp[2], len(planes), m, fin, weight.data_ptr(), ptr(bias), n, ptr(isc), ptr(ish),
                                    ptr(osc), ptr(osh), int(relu), ptr(res), out.data_ptr(), stream_ptr()), "hn_linear_f32")
    return out


def mesh_residual_upsample(x: torch.Tensor, skip: torch.Tensor, up: int) -> torch.Tensor:
    """[B, V, F] + linear interpolation of skip [B, V, Fs] along the feature axis, vertices repeated `up` times
    (meshnet.py:107-114, 69-76)."""
    _require_cuda(x, "x")
    b, v, f = x.shape
    assert skip.shape[:2] == (b, v) and x.is_contiguous() and skip.is_contiguous() and x.dtype == skip.dtype == torch.float32
    out = torch.empty((b, v * up, f), dtype=torch.float32, device=x.device)
    check(_lib.load().hn_mesh_residual_upsample(x.data_ptr(), skip.data_ptr(), b * v, f, skip.shape[2], up, out.data_ptr(),
                                                stream_ptr()), "hn_mesh_residual_upsample")
    return out

"""ctypes binding of libhandnet_b200.so (include/handnet_b200.h).

The library is built in-tree by hn_b200.build.  There is NO fallback: if it is missing, cannot be
loaded, or a call fails, a RuntimeError is raised.
"""
from __future__ import annotations

import ctypes as C
import os

import torch  # noqa: F401  (loads libcudart before our library so both share one runtime)

from . import build as _build

_c_p = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_d = C.c_double
_ip = C.POINTER(C.c_int)
_fp = C.POINTER(C.c_float)


class ConvDesc(C.Structure):
    """struct hn_conv_desc (include/handnet_b200.h)."""
    _fields_ = [
        ("in_", _c_p), ("n", _i), ("h", _i), ("w", _i), ("cin", _i), ("halo_in", _i), ("in_phases", _i),
        ("weight", _c_p), ("cout", _i), ("cout_pad", _i), ("kh", _i), ("kw", _i), ("stride", _i), ("dilation", _i),
        ("scale", _c_p), ("shift", _c_p), ("relu_lo", _i), ("relu_hi", _i),
        ("res", _c_p), ("res_mode", _i), ("res_h", _i), ("res_w", _i), ("res_halo", _i),
        ("out", _c_p), ("out_kind", _i), ("out_halo", _i),
        ("out_rows_per_image", _i), ("out_row_offset", _i), ("out_ld", _i), ("out_transpose_hw", _i),
        ("out_phase", _c_p), ("out_phase_halo", _i),
        ("gn_stats", _c_p), ("gn_groups", _i),
        ("block_n", _i), ("cluster", _i), ("debug", _i),
        ("splitk_ws", _c_p), ("splitk_ws_bytes", _i64), ("splitk_counters", _c_p), ("splitk_counters_len", _i), ("splits", _i),
        ("trace", _c_p), ("stem_pitch_h", _i), ("stem_pitch_w", _i), ("stem_window", _i),
    ]


# name -> (restype, argtypes); kept in one table so tests can check that every symbol declared in
# include/handnet_b200.h is exported and bound.
SIGNATURES = {
    "hn_last_error": (C.c_char_p, []),
    "hn_version": (_i, []),
    "hn_device_info": (_i, [_ip, _ip, _ip]),
    "hn_launch_count": (_i64, []),
    "hn_preprocess_resize_pad": (_i, [C.POINTER(_c_p), _ip, _ip, _ip, _ip, _i, _fp, _fp, _c_p, _i, _i, _c_p]),
    "hn_preprocess_resize_pad_framed": (_i, [C.POINTER(_c_p), _ip, _ip, _ip, _ip, _i, _fp, _fp, _c_p, _i, _i, _i, _i, _i, _i,
                                             _c_p]),
    "hn_im2col_7x7s2": (_i, [_c_p, _i, _i, _i, _i, _i, _c_p, _i, _c_p]),
    "hn_conv2d_bf16": (_i, [C.POINTER(ConvDesc), _c_p]),
    "hn_conv2d_bf16_levels": (_i, [C.POINTER(ConvDesc), _i, _c_p]),
    "hn_conv_set_cta_cap": (_i, [_i]),
    "hn_conv_set_pdl": (_i, [_i]),
    "hn_groupnorm_relu_levels": (_i, [C.POINTER(_c_p), _ip, _ip, _ip, _i, _i, _i, C.POINTER(_c_p), _i, _c_p, _c_p, _f, _c_p]),
    "hn_ingest_frames": (_i, [_c_p, _c_p, _i, _i, _i, _c_p, _c_p, _c_p]),
    "hn_pack_nhwc4_frame": (_i, [_c_p, _i, _i, _i, _i, _ip, _c_p, _i, _i, _i, _i, _c_p]),
    "hn_convert_joints": (_i, [_c_p, _c_p, _c_p, _c_p, _i, _i, _i, _i, _i, _c_p, _c_p]),
    "hn_maxpool3x3s2": (_i, [_c_p, _i, _i, _i, _i, _c_p, _i, _c_p]),
    "hn_groupnorm_relu": (_i, [_c_p, _i, _i, _i, _i, _i, _c_p, _i, _c_p, _c_p, _f, _c_p]),
    "hn_fcos_select_workspace_bytes": (_i64, [_i, _i]),
    "hn_fcos_decode_select": (_i, [_c_p, _i64, _i, _i, _c_p, _i64, _i, _c_p, _i64, _i, _i, _i, _i, _i, _i, _ip, _ip, _ip, _ip, _ip, _d,
                                    _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _i64, _c_p]),
    "hn_nms_workspace_bytes": (_i64, [_i, _i]),
    "hn_nms_batched": (_i, [_c_p, _c_p, _c_p, _c_p, _i, _i, _d, _i, _c_p, _c_p, _c_p, _i64, _c_p]),
    "hn_fcos_gather": (_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _i64, _i, _i, _c_p, _i64, _i, _i, _c_p, _i64, _i, _i,
                            _i, _i, _i, _i, _ip, _fp, _fp,
                            _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_p]),
    "hn_select_crop_resize": (_i, [_c_p, _c_p, _c_p, _i, _i, _i, _c_p, _i, _i, _i, _i, _c_p, _c_p, _c_p, _c_p]),
    "hn_select_crop_resize_multi": (_i, [_c_p, _c_p, _c_p, _i, _i, _i, _i, _c_p, _i, _i, _i, _i, _c_p, _c_p, _c_p, _c_p]),
    "hn_cheby_spmm": (_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _f, _f, _i, _i, _i, _c_p, _c_p]),
    "hn_linear_f32": (_i, [_c_p, _c_p, _c_p, _i, _i, _i, _c_p, _c_p, _i, _c_p, _c_p, _c_p, _c_p, _i, _c_p, _c_p, _c_p]),
    "hn_mesh_residual_upsample": (_i, [_c_p, _c_p, _i, _i, _i, _i, _c_p, _c_p]),
    "hn_a2j_workspace_bytes": (_i64, [_i, _i]),
    "hn_a2j_aggregate": (_i, [_c_p, _c_p, _c_p, _c_p, _i, _i, _i, _c_p, _c_p, _i64, _c_p]),
}

_lib = None


def lib_path() -> str:
    return _build.LIB


def load(build_if_missing: bool = True):
    """Load (building first if the sources are newer) and type every entry point."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("HN_LIB_AB")          # A/B timing of two builds of the library in one GPU call (tools only)
    if not path:
        if build_if_missing and _build.needs_build():
            _build.build()
        path = _build.LIB
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: run `python -m hn_b200.build` (there is no CPU fallback)")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError here means the header and the library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().hn_last_error()
        raise RuntimeError(f"{what} failed (status {rc}): {msg.decode() if msg else '?'}")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def ptr(t) -> int:
    return 0 if t is None else t.data_ptr()

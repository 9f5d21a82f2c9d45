"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol that
include/handnet_b200.h declares, and compute entry points fail loudly (no CPU fallback) without a GPU."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "handnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(hn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    from hn_b200 import _lib
    lib = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"
        assert s in _lib.SIGNATURES, f"{s} has no ctypes signature"
    assert set(_lib.SIGNATURES) == set(syms)
    assert lib.hn_version() >= 100


def test_conv_desc_layout_matches_header():
    """ctypes mirror of struct hn_conv_desc: field order follows the header."""
    from hn_b200 import _lib
    text = open(os.path.join(ROOT, "include", "handnet_b200.h")).read()
    body = text[text.index("typedef struct hn_conv_desc {"): text.index("} hn_conv_desc;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    names = []
    for decl in body.split("{", 1)[1].split(";"):
        decl = decl.strip()
        if not decl:
            continue
        for part in decl.split(","):
            names.append(re.findall(r"(\w+)\s*$", part.strip())[0])
    mine = [n.rstrip("_") for n, _ in _lib.ConvDesc._fields_]
    assert names == mine


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from hn_b200 import _lib, ops
    lib = _lib.load()
    with pytest.raises(RuntimeError):
        ops.a2j_aggregate(torch.zeros(1, 4, 2), torch.zeros(1, 4, 2, 2), torch.zeros(1, 4, 2), torch.zeros(4, 2))
    sm = ctypes.c_int()
    assert lib.hn_device_info(ctypes.byref(sm), None, None) != 0
    assert lib.hn_last_error()


def test_argument_validation_reports_errors():
    from hn_b200 import _lib
    lib = _lib.load()
    d = _lib.ConvDesc()
    assert lib.hn_conv2d_bf16(ctypes.byref(d), None) == -1
    assert b"null pointer" in lib.hn_last_error()
    assert lib.hn_nms_workspace_bytes(2, 17850) > 2 * 17850 * 279 * 8

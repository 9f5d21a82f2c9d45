"""Chebyshev graph convolution on the GPU kernels of hn_b200 (csrc/hn_mesh.cu).

Same call as /root/reference/pose2mesh/lib/models/backbones/cheby_graph_conv.py:5 -- graph_conv_cheby(x, cl, bn, L, Fout, K):
x [B, V, Fin] -> [B, V, Fout] = BN(Linear([T_0 x | T_1 x | ... ] interleaved per feature)), T_0 = x, T_1 = L x,
T_k = 2 L T_{k-1} - T_{k-2} (:26-31).  Eval-mode BatchNorm only (folded into the linear layer's epilogue); no ReLU here, the
caller applies it (meshnet.py:100-101) -- `relu=True` lets the caller fuse it."""
from __future__ import annotations

import torch

from hn_b200 import ops

_graph_cache = {}


def as_graph(L, device) -> "ops.CsrGraph":
    """CSR copy of a Laplacian on `device`, cached per object (the model holds its Laplacians for its lifetime)."""
    if isinstance(L, ops.CsrGraph):
        return L
    key = (id(L), str(device))
    g = _graph_cache.get(key)
    if g is None or g[0] is not L:
        g = (L, ops.CsrGraph(L, device))
        _graph_cache[key] = g
    return g[1]


_bn_cache = {}


def bn_affine(bn):
    """Eval-mode BatchNorm1d as (scale, shift), folded once per set of parameter versions."""
    if bn is None:
        return None
    if bn.training:
        raise NotImplementedError("pose2mesh on the B200 build is inference only: call .eval() (BatchNorm uses running statistics)")
    ts = (bn.weight, bn.bias, bn.running_mean, bn.running_var)
    key = tuple((t.data_ptr(), -1 if t.is_inference() else t._version) for t in ts)      # (inference tensors are immutable)
    hit = _bn_cache.get(id(bn))
    if hit is None or hit[0] != key:
        scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).detach().float().contiguous()
        shift = (bn.bias - bn.running_mean * scale).detach().float().contiguous()
        hit = (key, scale, shift)
        _bn_cache[id(bn)] = hit
    return hit[1], hit[2]


def graph_conv_cheby(x, cl, bn, L, Fout, K, relu: bool = False):
    B, V, Fin = x.shape
    assert cl.weight.shape == (Fout, Fin * K) and 1 <= K <= 3, "Chebyshev order 1..3 (meshnet.py:24,32 uses 3)"
    g = as_graph(L, x.device)
    x = x.contiguous().float()
    planes = [x]
    if K > 1:
        planes.append(ops.cheby_spmm(g, x))                                               # T1 = L x
    if K > 2:
        planes.append(ops.cheby_spmm(g, planes[1], z=x, alpha=2.0, beta=-1.0))            # T2 = 2 L T1 - T0
    y = ops.linear_f32([p.view(B * V, Fin) for p in planes], cl.weight.detach().float().contiguous(),
                       None if cl.bias is None else cl.bias.detach().float().contiguous(), out_affine=bn_affine(bn), relu=relu)
    return y.view(B, V, Fout)

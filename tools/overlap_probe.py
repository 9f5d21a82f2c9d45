"""Does the memory-bound GroupNorm kernel run UNDERNEATH a 256-wide tower convolution (blocks co-resident with the persistent
conv CTAs)?  Times 8 fused-level tower convolutions on stream A, 8 GroupNorm passes on stream B, and both together."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops

dev = "cuda"
B = 8
levels = [(100, 136), (50, 68), (25, 34)]
g = torch.Generator().manual_seed(0)
wp = ops.pack_conv_weight((torch.randn(256, 256, 3, 3, generator=g) * 0.02).to(dev))
bias = torch.zeros(256, device=dev)
gamma, beta = torch.ones(256, device=dev), torch.zeros(256, device=dev)
def mk():
    xs = [ops.Act(B, h, w, 256, 1, dev) for h, w in levels]
    for x in xs: x.interior().normal_()
    return xs
xa, oa = mk(), mk()
xb = mk()
sta = [torch.zeros(B, 32, 2, dtype=torch.int64, device=dev) for _ in levels]
stb = [torch.zeros(B, 32, 2, dtype=torch.int64, device=dev) for _ in levels]
for s_, (h, w) in zip(stb, levels): s_[..., 1] = int(h * w * 8 * (1 << 24))
sa, sb = torch.cuda.Stream(), torch.cuda.Stream()
N = 8
def conv():
    for _ in range(N): ops.conv2d_levels(xa, wp, cout=256, ksize=3, shift=bias, outs=oa, gn_stats=sta, gn_groups=32)
def gn():
    for _ in range(N): ops.groupnorm_relu_levels(xb, stb, 32, gamma, beta, 1e-5)
def timed(fa, fb):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cur = torch.cuda.current_stream()
    torch.cuda._sleep(20_000_000)
    e0.record(cur)
    sa.wait_event(e0); sb.wait_event(e0)
    if fa:
        with torch.cuda.stream(sa): fa()
    if fb:
        with torch.cuda.stream(sb): fb()
    cur.wait_stream(sa); cur.wait_stream(sb)
    e1.record(cur)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3
for _ in range(2): timed(conv, gn)
for name, fa, fb in (("conv only", conv, None), ("gn only", None, gn), ("both", conv, gn)):
    print(f"{name:10s} {min(timed(fa, fb) for _ in range(5)) / N:8.1f} us per layer", flush=True)

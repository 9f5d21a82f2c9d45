"""Latency of one small A2J-sized convolution: L2-hot (same launch replayed back to back) vs L2-cold (a 256 MiB
write between launches), both from CUDA graphs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def graph_time(fn, reps=30):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

for name, (n, h, w, cin, cout, k) in (("1x1 256->1024 @11x11", (8, 11, 11, 256, 1024, 1)), ("1x1 1024->256 @11x11", (8, 11, 11, 1024, 256, 1)),
                                      ("3x3 256->256 @11x11", (8, 11, 11, 256, 256, 3)), ("3x3 512->512 d2 @11x11", (8, 11, 11, 512, 512, 3)),
                                      ("3x3 2048->256 @11x11", (8, 11, 11, 2048, 256, 3)), ("1x1 64->256 @44x44", (8, 44, 44, 64, 256, 1))):
    g = torch.Generator().manual_seed(0)
    x = ops.Act.from_nchw(torch.randn(n, cin, h, w, generator=g).cuda(), 1)
    wt = ops.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) / 30).cuda())
    out = ops.Act(n, h, w, cout, 1, "cuda")
    sc = torch.ones(cout, device="cuda"); sh = torch.zeros(cout, device="cuda")
    conv = lambda: ops.conv2d(x, wt, cout=cout, ksize=k, scale=sc, shift=sh, relu=True, out=out)
    def hot():
        for _ in range(20): conv()
    def cold():
        for _ in range(20):
            flush.zero_(); conv()
    def only_flush():
        for _ in range(20): flush.zero_()
    t_hot = graph_time(hot) / 20
    t_cold = (graph_time(cold) - graph_time(only_flush)) / 20
    print(f"{name:26s} hot {t_hot:6.1f} us   cold {t_cold:6.1f} us   weights {wt.numel() * 2 / 1e6:5.2f} MB", flush=True)

"""Where does a narrow-N conv spend its time?  Times a layer1-shaped conv with the epilogue partly disabled."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops

def time_conv(n, h, w, cin, cout, k, debug, block_n=0, res=False, reps=20):
    g = torch.Generator().manual_seed(0)
    x = ops.Act.from_nchw(torch.randn(n, cin, h, w, generator=g).cuda(), 1)
    wt = ops.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) / 30).cuda())
    out = ops.Act(n, h, w, cout, 1, "cuda")
    r = ops.Act.from_nchw(torch.randn(n, cout, h, w, generator=g).cuda(), 1) if res else None
    sc = torch.ones(cout, device="cuda"); sh = torch.zeros(cout, device="cuda")
    def run():
        ops.conv2d(x, wt, cout=cout, ksize=k, scale=sc, shift=sh, relu=True, res=r, res_mode=1 if res else 0, out=out,
                   block_n=block_n, debug=debug)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

NOEPI = 128
for name, shp in (("layer1 64->64 3x3 @200x272", (8, 200, 272, 64, 64, 3)), ("layer2 128->128 3x3 @100x136", (8, 100, 136, 128, 128, 3)),
                  ("P3 256->256 3x3 @100x136", (8, 100, 136, 256, 256, 3))):
    for dbg, what in ((0, "full"), (1024, "full, release after lookahead"), (NOEPI | 16, "no epilogue"),
                      (NOEPI | 16 | 1024, "no epilogue, release after la"), (NOEPI | 16 | 512, "no epilogue, no TMA"),
                      (NOEPI | 16 | 512 | 1024, "no epi, no TMA, rel. after la")):
        print(f"{name:32s} {what:30s} {time_conv(*shp, debug=dbg):8.1f} us", flush=True)

"""Where does a conv launch spend its time?  Times the dominant layer shapes with parts of the kernel disabled
(debug bits of hn_conv_desc: 64 no stores, 128 no epilogue, 256 no MMA, 512 no TMA, 16 no resident weights)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops

def time_conv(n, h, w, cin, cout, k, debug, block_n=0, res=False, reps=20, cluster=0):
    g = torch.Generator().manual_seed(0)
    x = ops.Act.from_nchw(torch.randn(n, cin, h, w, generator=g).cuda(), 1)
    wt = ops.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) / 30).cuda())
    out = ops.Act(n, h, w, cout, 1, "cuda")
    r = ops.Act.from_nchw(torch.randn(n, cout, h, w, generator=g).cuda(), 1) if res else None
    sc = torch.ones(cout, device="cuda"); sh = torch.zeros(cout, device="cuda")
    def run():
        ops.conv2d(x, wt, cout=cout, ksize=k, scale=sc, shift=sh, relu=True, res=r, res_mode=1 if res else 0, out=out,
                   block_n=block_n, debug=debug, cluster=cluster)
    for _ in range(3): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3

NOEPI, NOMMA, NOTMA, NORB = 128, 256, 512, 16
shapes = (("layer1 64->64 3x3 @200x272", (8, 200, 272, 64, 64, 3)), ("layer2 128->128 3x3 @100x136", (8, 100, 136, 128, 128, 3)),
          ("P3 256->256 3x3 @100x136", (8, 100, 136, 256, 256, 3)), ("layer3 256->256 3x3 @50x68", (8, 50, 68, 256, 256, 3)),
          ("layer4 512->512 3x3 @25x34", (8, 25, 34, 512, 512, 3)))
only = sys.argv[1:] 
for name, shp in shapes:
    if only and not any(o in name for o in only):
        continue
    flops = 2.0 * shp[0] * shp[1] * shp[2] * shp[3] * shp[4] * shp[5] ** 2
    for dbg, what in ((0, "full"), (64, "no stores"), (NOEPI, "no epilogue"), (NOEPI | NOMMA, "no epilogue, no MMA"), (NOEPI | NOTMA, "no epilogue, no TMA"),
                      (NOEPI | NOMMA | NOTMA, "barriers only"), (NORB, "full, streamed weights"), (-1, "full, cluster 2"),
                      (2048 | NOEPI | NOMMA | NOTMA, "barriers only, no epi warps"), (2048 | NOEPI | NOTMA, "MMA only, no epi warps"),
                      (2048 | NOEPI, "MMA+TMA, no epi warps"), (1024, "full, 1-lane epi polling")):
        if dbg == NORB and shp[4] > 64:
            continue
        if dbg == -1 and shp[4] < 256:
            continue
        us = time_conv(*shp, debug=max(dbg, 0), cluster=2 if dbg == -1 else 0)
        print(f"{name:32s} {what:26s} {us:8.1f} us  {flops / us / 1e6:7.0f} TF/s", flush=True)

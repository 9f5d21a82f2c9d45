"""[needs a library built with the bring-up instrumentation: python -m hn_b200.build --debug (HN_CONV_DEBUG build)]
What bounds the 64->64 3x3 @200x272 x8 convolution (FCOS layer1, resident-weights path)?  Times the launch with the
kernel's timing-experiment flags (hn_conv_desc.debug bits 6..9: no stores / no epilogue / no MMA / no TMA) and prints
the L2 -> SM bytes the operand boxes need.  python tools/layer1_ablation.py [shape: layer1|layer2|p3]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops

shape = sys.argv[1] if len(sys.argv) > 1 else "layer1"
n, h, w, c, co = {"layer1": (8, 200, 272, 64, 64), "layer2": (8, 100, 136, 128, 128), "p3": (8, 100, 136, 256, 256)}[shape]
g = torch.Generator().manual_seed(0)
x = ops.Act.from_nchw(torch.randn(n, c, h, w, generator=g).cuda(), 1)
wt = ops.pack_conv_weight((torch.randn(co, c, 3, 3, generator=g) * (c * 9) ** -0.5).cuda())
scale = torch.ones(co, device="cuda"); shift = torch.zeros(co, device="cuda")
out = ops.Act(n, h, w, co, 1, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def t(debug, reps=20):
    for _ in range(3):
        ops.conv2d(x, wt, cout=co, ksize=3, scale=scale, shift=shift, relu=True, out=out, debug=debug)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d(x, wt, cout=co, ksize=3, scale=scale, shift=shift, relu=True, out=out, debug=debug)
        e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

rows = n * (h + 2) * (w + 2)
tiles = (rows + 127) // 128
print(f"{shape}: {n}x{h}x{w} {c}->{co}, {tiles} M tiles, activations {rows * c * 2 / 1e6:.1f} MB, "
      f"A boxes per launch {tiles * 3 * 136 * 128 * (c // 64) * max(1, co // 256) / 1e6:.1f} MB (3 kernel rows x 136 rows x 128 B per chunk and tile)")
base = int(os.environ.get("HN_DEBUG_BASE", "0"))     # e.g. 16384 = flattened tiles instead of patch tiles
for name, dbg in (("full", 0), ("no stores", 64), ("no epilogue", 128), ("no MMA", 256), ("no MMA, no epilogue", 256 | 128),
                  ("no TMA", 512), ("no TMA, no epilogue", 512 | 128), ("no TMA, no MMA, no epilogue (role loops only)", 512 | 256 | 128),
                  ("no TMA, no MMA", 512 | 256), ("no TMA, no MMA, no epilogue ROLE (producer <-> MMA only)", 512 | 256 | 2048),
                  ("no TMA, no epilogue ROLE (MMA issue + producer handshake)", 512 | 2048)):
    print(f"  {name:60s} {t(dbg | base):8.1f} us", flush=True)

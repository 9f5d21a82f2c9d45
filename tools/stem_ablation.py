"""[needs a library built with the bring-up instrumentation: python -m hn_b200.build --debug (HN_CONV_DEBUG build)]
What bounds the direct 7x7/2 stem (8 VGA frames: 400x544 outputs, K = 256, 64 channels)?  Same timing flags as
tools/layer1_ablation.py.  python tools/stem_ablation.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops
from hn_b200.runtime import STEM_K_RGB

B, hc, wc = 8, 800, 1088
g = torch.Generator().manual_seed(0)
frame = ops.StemFrame(B, (hc, wc), "cuda")
frame.set_canvas(torch.randn(B, hc, wc, 4, generator=g).cuda())
w = ops.pack_stem_weight((torch.randn(64, 3, 7, 7, generator=g) * 0.05).cuda(), STEM_K_RGB)
scale = torch.ones(64, device="cuda"); shift = torch.zeros(64, device="cuda")
out = ops.Act(B, hc // 2, wc // 2, 64, 0, "cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def run(debug):
    ops.conv2d(frame, w, cout=64, ksize=1, scale=scale, shift=shift, relu=True, out=out, algo_k=147, debug=debug)

def t(debug, reps=15):
    for _ in range(3): run(debug)
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(debug); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]

print(f"stem: {B} x {hc // 2}x{wc // 2} outputs, {B * (hc // 2) * ((wc // 2 + 127) // 128)} tiles of 128 columns, output {B * hc * wc // 4 * 128 / 1e6:.0f} MB")
for name, dbg in (("full", 0), ("alternate-tile epilogue off", 32768), ("no stores", 64), ("no epilogue", 128), ("no MMA", 256),
                  ("no MMA, no epilogue (TMA only)", 256 | 128), ("no TMA", 512), ("no TMA, no epilogue (MMA only)", 512 | 128),
                  ("no TMA, no MMA (epilogue only)", 512 | 256), ("role loops only", 512 | 256 | 128),
                  ("producer <-> MMA loops only (no epilogue role)", 512 | 256 | 2048), ("MMA + producer, no TMA, no epilogue role", 512 | 2048),
                  ("TMA + MMA, no epilogue role", 2048)):
    print(f"  {name:50s} {t(dbg):8.1f} us", flush=True)

import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch, torch.nn.functional as F
from hn_b200 import ops
torch.manual_seed(0)
n, h, w = 1, 64, 96
mode = sys.argv[1] if len(sys.argv) > 1 else "rand"
canvas = torch.zeros(n, h, w, 4)
if mode == "ones":
    canvas[..., :3] = 1.0
else:
    canvas[..., :3] = torch.randn(n, h, w, 3)
wt = torch.zeros(64, 3, 7, 7)
if mode == "ones":
    wt[:, 0, 3, 3] = 1.0          # centre tap only: output == input at (2oy, 2ox)
else:
    wt = torch.randn(64, 3, 7, 7) * 0.05
frame = ops.StemFrame(n, (h, w), "cuda")
frame.set_canvas(canvas.to(torch.bfloat16))
wp = ops.pack_stem_weight(wt.cuda(), 256)
stem = ops.Act(n, h // 2, w // 2, 64, 0, "cuda")
ops.conv2d(frame, wp, cout=64, ksize=1, out=stem)
torch.cuda.synchronize()
got = stem.to_nchw().cpu()
ref = F.conv2d(canvas[..., :3].permute(0, 3, 1, 2).to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), stride=2, padding=3)
print("finite frac", torch.isfinite(got).float().mean().item(), "max abs err", (got - ref).abs()[torch.isfinite(got)].max().item() if torch.isfinite(got).any() else None)
print("got[0,0,:4,:8]\n", got[0, 0, :4, :8]); print("ref[0,0,:4,:8]\n", ref[0, 0, :4, :8])
if mode == "time":
    n, h, w = 8, 800, 1088
    frame = ops.StemFrame(n, (h, w), "cuda")
    frame.set_canvas(torch.randn(frame.n, frame.hc, frame.wc, 4, device="cuda"))
    stem = ops.Act(n, h // 2, w // 2, 64, 0, "cuda")
    sc = torch.ones(64, device="cuda"); sh = torch.zeros(64, device="cuda")
    trace = torch.zeros(3 * 2048 * 2, dtype=torch.int64, device="cuda")
    for dbg, what in ((0, "full"), (64, "no stores"), (128, "no epilogue"), (128 | 256, "no epilogue, no MMA"), (128 | 512, "no epilogue, no TMA")):
        for _ in range(3): ops.conv2d(frame, wp, cout=64, ksize=1, scale=sc, shift=sh, relu=True, out=stem, debug=dbg)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): ops.conv2d(frame, wp, cout=64, ksize=1, scale=sc, shift=sh, relu=True, out=stem, debug=dbg)
        e1.record(); torch.cuda.synchronize()
        print(f"stem 8x800x1088 {what:24s} {e0.elapsed_time(e1) / 10 * 1e3:8.1f} us")
    ops.conv2d(frame, wp, cout=64, ksize=1, scale=sc, shift=sh, relu=True, out=stem, trace=trace)
    torch.cuda.synchronize()
    tr = trace.cpu().view(3, 2048, 2)
    for r, name in enumerate(("producer", "mma", "epilogue")):
        ev = [int(a) for a, b in tr[r].tolist() if a > 0]
        d = [b - a for a, b in zip(ev[:-1], ev[1:])]
        mid = d[len(d) // 4: 3 * len(d) // 4]
        print(name, len(ev), "events; mean gap mid", sum(mid) / max(1, len(mid)), "first gaps", d[:12])

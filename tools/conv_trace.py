"""[needs a library built with the bring-up instrumentation: python -m hn_b200.build --debug (HN_CONV_DEBUG build)]
Timeline of CTA 0 of one conv launch (hn_conv_desc.trace): per role, the clock64 deltas between events.
    python tools/conv_trace.py layer1|layer2|P3|layer3 [debug_flags]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "handnet-pipeline_b200"))
import torch
from hn_b200 import ops

SHAPES = {"layer1": (8, 200, 272, 64, 64, 3), "layer2": (8, 100, 136, 128, 128, 3), "P3": (8, 100, 136, 256, 256, 3),
          "layer3": (8, 50, 68, 256, 256, 3), "a2j3x3": (8, 11, 11, 256, 256, 3), "a2j1x1": (8, 11, 11, 1024, 256, 1),
          "a2j1x1b": (8, 11, 11, 256, 1024, 1)}
name = sys.argv[1] if len(sys.argv) > 1 else "layer1"
debug = int(sys.argv[2]) if len(sys.argv) > 2 else 0
n, h, w, cin, cout, k = SHAPES[name]
g = torch.Generator().manual_seed(0)
x = ops.Act.from_nchw(torch.randn(n, cin, h, w, generator=g).cuda(), 1)
wt = ops.pack_conv_weight((torch.randn(cout, cin, k, k, generator=g) / 30).cuda())
out = ops.Act(n, h, w, cout, 1, "cuda")
sc = torch.ones(cout, device="cuda"); sh = torch.zeros(cout, device="cuda")
trace = torch.zeros(3 * 2048 * 2, dtype=torch.int64, device="cuda")
for i in range(3):
    ops.conv2d(x, wt, cout=cout, ksize=k, scale=sc, shift=sh, relu=True, out=out, debug=debug, trace=trace if i == 2 else None,
               block_n=int(os.environ.get("BN", "0")))
torch.cuda.synchronize()
tr = trace.cpu().view(3, 2048, 2)
t0 = min(int(tr[r, 0, 0]) for r in range(3) if tr[r, 0, 0] > 0)
ROLE = ("producer", "mma", "epilogue")
TAGS = ({1: "a_empty ok", 2: "b_empty ok"}, {1: "a_full ok", 2: "b_full ok", 3: "commits issued", 4: "tmem_empty ok", 5: "before a_full wait", 6: "mmas issued"},
        {1: "tile start", 2: "tmem_full ok", 3: "tile done", 4: "tmem ld done", 5: "stored", 6: "math done", 7: "packed"})
limit = int(os.environ.get("TRACE_N", "60"))
for r in range(3):
    ev = [(int(a), int(b)) for a, b in tr[r].tolist() if a > 0]
    print(f"== {ROLE[r]}: {len(ev)} events, span {ev[-1][0] - ev[0][0] if ev else 0} cycles")
    prev = None
    for i, (t, tag) in enumerate(ev[:limit]):
        print(f"  {t - t0:8d}  (+{0 if prev is None else t - prev:5d})  {TAGS[r].get(tag, tag)}")
        prev = t
    # steady-state statistics from the middle of the trace
    if len(ev) > 40:
        mid = ev[len(ev) // 4: 3 * len(ev) // 4]
        span = mid[-1][0] - mid[0][0]
        by = {}
        for (ta, _), (tb, tag) in zip(mid[:-1], mid[1:]):
            by.setdefault(tag, []).append(tb - ta)
        print("  steady state: " + ", ".join(f"{TAGS[r].get(k, k)}: n={len(v)} mean +{sum(v) / len(v):.0f}" for k, v in sorted(by.items()))
              + f"; {span / max(1, len(mid) - 1):.0f} cycles/event")

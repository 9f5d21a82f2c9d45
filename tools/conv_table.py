"""Aggregate a per-launch conv table (HN_CONV_TABLE=... python bench.py) by layer shape.
    python tools/conv_table.py gpurun_out/conv_table.json [peak_tflops]"""
import json, sys
from collections import OrderedDict
rows = json.load(open(sys.argv[1]))
peak = float(sys.argv[2]) if len(sys.argv) > 2 else 1356.9
agg = OrderedDict()
for r in rows:
    key = (r["h"], r["w"], r["cin"], r["cout"], r["k"], r["stride"], r.get("dil", 1), bool(r.get("gn")), bool(r.get("f32")), r.get("levels", 1))
    a = agg.setdefault(key, [0, 0.0, 0.0])
    a[0] += 1; a[1] += r["ms"]; a[2] += r["gflop"]
tot = sum(a[1] for a in agg.values())
print("| map (HxW) | Cin->Cout | k/s/d | levels | epilogue | launches | us total | TF/s | % of peak | share |")
print("|---|---|---|---|---|---|---|---|---|---|")
for key, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    h, w, cin, cout, k, s, d, gn, ok, lv = key
    tf = a[2] / a[1]
    epi = "GN stats" if gn else ("fp32 rows" if ok else "")
    print(f"| {h}x{w} | {cin}->{cout} | {k}/{s}/{d} | {lv} | {epi} | {a[0]} | {a[1]*1e3:.1f} | {tf:.0f} | {100*tf/peak:.0f}% | {100*a[1]/tot:.1f}% |")
gf = sum(a[2] for a in agg.values())
print(f"| **all** | | | | | {len(rows)} | {tot*1e3:.1f} | {gf/tot:.0f} | {100*gf/tot/peak:.0f}% | 100% |")

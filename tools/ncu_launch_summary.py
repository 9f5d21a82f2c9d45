"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel:
python tools/ncu_launch_summary.py gpurun_out/launches.csv "<header line>" > profiles/rNN_launches_summary.txt"""
import collections
import csv
import sys

tot = collections.OrderedDict()
for r in csv.reader(open(sys.argv[1])):
    if len(r) < 10 or not r[0].isdigit() or r[-3] != "gpu__time_duration.sum":
        continue
    name = r[6] if "Kernel Name" not in r else None
    k = tot.setdefault(name[:70], [0, 0.0])
    k[0] += 1
    k[1] += float(r[-1].replace(",", "")) / 1e3
for line in sys.argv[2:]:
    print("# " + line)
print("# per-launch times are cold-cache and serialised: compare shares.\n")
total = sum(v[1] for v in tot.values())
print(f"{'kernel':70s} {'launches':>8} {'total us':>10} {'share':>7}")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:70s} {n:8d} {us:10.1f} {100 * us / total:6.1f}%")
print(f"{'total':70s} {sum(v[0] for v in tot.values()):8d} {total:10.1f}")

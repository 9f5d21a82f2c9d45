"""Summarise the ncu metric CSV of tools/membound_roofline.py (see the command in the header it writes):
python tools/ncu_membound_summary.py gpurun_out/membound_ncu.csv > profiles/rNN_membound_ncu.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for x in rows:
    k = x[4].split("(")[0].replace("<unnamed>::", "").replace("void ", "")
    agg.setdefault((int(x[0]), k), {})[x[-3]] = float(x[-1].replace(",", ""))
print("# ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,"
      "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed --clock-control none")
print("#   -k 'regex:preprocess|maxpool|groupnorm|select_|nms_|a2j_|gather_kernel'   HN_MEMBOUND_REPS=1 python tools/membound_roofline.py")
print("# One B200, the large-batch shapes of tools/membound_roofline.py (64 frames / 256 frames / 512 crops / 8 x ~10k candidates);")
print("# a warm-up and a measured launch of each kernel, both listed.  DRAM bytes = what the launch really moved (cold L2);")
print("# GB/s = DRAM bytes / duration; 'ncu dram %' = gpu__dram_throughput against ncu's own peak (~8 TB/s).")
print()
print(f"{'id':>3} {'kernel':28s} {'us':>9} {'DRAM rd MB':>11} {'DRAM wr MB':>11} {'GB/s':>7} {'ncu dram %':>10}")
for (i, k), d in agg.items():
    t = d["gpu__time_duration.sum"] / 1e3
    rd, wr = d["dram__bytes_read.sum"] / 1e6, d["dram__bytes_write.sum"] / 1e6
    print(f"{i:>3} {k:28s} {t:9.1f} {rd:11.1f} {wr:11.1f} {(rd + wr) / t * 1e3:7.0f} "
          f"{d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']:10.1f}")

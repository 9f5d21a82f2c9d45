"""Key metrics of one or more `ncu --set full` captures as text: python tools/ncu_full_summary.py "<title>" file.ncu-rep [...]
(reads the report with `ncu -i ... --page raw --csv`; one column per captured launch)."""
import csv, io, subprocess, sys
KEYS = ["gpu__time_duration.sum", "gpc__cycles_elapsed.avg.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "lts__t_sector_hit_rate.pct", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active"]
STALL = "smsp__average_warps_issue_stalled_"
for title, path in zip(sys.argv[1::2], sys.argv[2::2]):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"## {title}\n# ({path}; one column per captured launch: {', '.join(r[hdr.index('Kernel Name')][:60] for r in data[:1])})\n")
    keys = KEYS + sorted(h for h in hdr if h.startswith(STALL) and h.endswith("_per_issue_active.ratio")
                         and any(float(r[hdr.index(h)].replace(',', '') or 0) >= 0.05 for r in data))
    for k in keys:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:92s} {units[i]:16s} " + "  ".join(r[i] for r in data))
    print()

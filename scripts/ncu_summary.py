#!/usr/bin/env python
"""Turn an .ncu-rep (ncu --set full) into the small per-kernel CSV kept under profiles/.
  python scripts/ncu_summary.py gpurun_out/prof_x.ncu-rep profiles/r01_x_ncu_full.csv"""
import csv
import subprocess
import sys

COLS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__cycles_elapsed.avg.per_second",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"]


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = [hdr.index(c) for c in COLS if c in hdr]
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([hdr[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in data:
            w.writerow([r[i] for i in idx])
    for r in data:
        d = dict(zip(hdr, r))
        gb = float(d["dram__bytes_read.sum"]) + float(d["dram__bytes_write.sum"])
        u = units[hdr.index("dram__bytes_read.sum")]
        ms = float(d["gpu__time_duration.sum"])
        tu = units[hdr.index("gpu__time_duration.sum")]
        print(d["Kernel Name"][:70], "| %.3f %s | dram %.3f %s | dram %% %s | regs %s" % (
            ms, tu, gb, u, d.get("dram__throughput.avg.pct_of_peak_sustained_elapsed", d.get("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")),
            d["launch__registers_per_thread"]))


if __name__ == "__main__":
    main()

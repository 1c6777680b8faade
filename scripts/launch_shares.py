#!/usr/bin/env python
"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --clock-control none --csv).
  python scripts/launch_shares.py gpurun_out/x_launches.csv "header line" > profiles/x_launch_shares.txt"""
import csv, sys, collections


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 8]
    hdr = rows[0]
    iname, imet, ival, iunit = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in rows[1:]:
        if r[imet] != "gpu__time_duration.sum":
            continue
        v = float(r[ival].replace(",", ""))
        v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[iunit], 1e-6)
        name = r[iname].split("(")[0][:64]
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    if len(sys.argv) > 2:
        print("# " + sys.argv[2])
    for name, v in tot.most_common():
        print(f"{name:64s} {cnt[name]:5d} launches {v:10.3f} ms {100 * v / s:5.1f}%")
    print(f"{'total':64s} {sum(cnt.values()):5d} launches {s:10.3f} ms")


if __name__ == "__main__":
    main()

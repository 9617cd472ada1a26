#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list as the markdown table kept under profiles/:
kernel, launches, total / average duration, share of GPU kernel time (runs here, no GPU needed).

    python scripts/ncu_launch_list.py gpurun_out/r2ncu/launches_r2.csv > profiles/launches_r2.md"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1], errors="replace")))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r and "Metric Value" in r)
    names = rows[hdr]
    k, v, u, m = names.index("Kernel Name"), names.index("Metric Value"), names.index("Metric Unit"), names.index("Metric Name")
    agg = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) != len(names) or r[m] != "gpu__time_duration.sum":
            continue
        t = float(r[v].replace(",", ""))
        t_us = t / 1e3 if r[u].startswith("ns") else t * 1e3 if r[u].startswith("ms") else t
        name = r[k].split("(")[0].split("::")[-1]
        a = agg.setdefault(name, [0, 0.0, []])
        a[0] += 1; a[1] += t_us; a[2].append(t_us)
    total = sum(a[1] for a in agg.values())
    print("| kernel | launches | total us | avg us | median us | share of GPU kernel time |")
    print("|---|---:|---:|---:|---:|---:|")
    for name, (n, t, ts) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        ts.sort()
        print(f"| `{name}` | {n} | {t:.1f} | {t / n:.1f} | {ts[len(ts) // 2]:.1f} | {t / total:.3f} |")


if __name__ == "__main__":
    main()

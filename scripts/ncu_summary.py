#!/usr/bin/env python
"""Summarise an `ncu --set full` report as the markdown table kept under profiles/ (run HERE, no GPU needed):

    python scripts/ncu_summary.py gpurun_out/round2/k5_full.ncu-rep [--kernel regex] [--alg-bytes N] > profiles/k5_full_r2.md

One column per captured launch; the rows are the counters the judge reads for an HBM-bound kernel: duration, DRAM bytes
read / written, DRAM throughput %, L2 read sectors, occupancy, registers, grid.  With --alg-bytes (algorithmic bytes per
launch, read + write) the achieved GB/s and its fraction of MEASURED_PEAKS.json's hbm_gbs are appended."""
import argparse
import csv
import io
import json
import os
import re
import subprocess
import sys

ROWS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg",
        "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_write.sum", "l1tex__m_xbar2l1tex_read_bytes.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum"]


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--kernel", default=None, help="regex on the kernel name (default: every captured launch)")
    ap.add_argument("--alg-bytes", type=float, default=None, help="algorithmic bytes per launch (read + write)")
    ap.add_argument("--max-launches", type=int, default=6)
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    names, units, data = rows[hdr], rows[hdr + 1], rows[hdr + 2:]
    kcol = names.index("Kernel Name")
    launches = [r for r in data if len(r) == len(names) and (not a.kernel or re.search(a.kernel, r[kcol]))][: a.max_launches]
    if not launches:
        sys.exit("no launch matches")
    print(f"# ncu --set full: {os.path.basename(a.report)} ({len(launches)} launch(es) of `{launches[0][kcol].split('(')[0]}`)\n")
    print("| metric | " + " | ".join(f"launch {i + 1}" for i in range(len(launches))) + " | unit |")
    print("|---|" + "---:|" * len(launches) + "---|")
    col = {n: i for i, n in enumerate(names)}
    for m in ROWS:
        if m in col:
            print(f"| `{m}` | " + " | ".join(r[col[m]] for r in launches) + f" | {units[col[m]]} |")
    if a.alg_bytes and "gpu__time_duration.sum" in col:
        peak = None
        try:
            peak = float(json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:
            pass
        scale = {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "s": 1.0, "second": 1.0}.get(units[col["gpu__time_duration.sum"]], 1e-9)
        print()
        for i, r in enumerate(launches):
            t = float(r[col["gpu__time_duration.sum"]].replace(",", "")) * scale
            g = a.alg_bytes / t / 1e9
            print(f"* launch {i + 1}: {t * 1e6:.1f} us -> {g:.0f} GB/s algorithmic (read + write)" + (f" = {g / peak:.3f} of the measured {peak} GB/s copy peak" if peak else ""))


if __name__ == "__main__":
    main()

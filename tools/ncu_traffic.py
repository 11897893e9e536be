#!/usr/bin/env python
"""Write profiles/ncu_traffic.json from an `ncu --set full` report of `python bench.py` (same build, same command):

    python tools/ncu_traffic.py gpurun_out/<tag>.ncu-rep [--tag <tag>]

For every kernel in the report: dram__bytes_read.sum + dram__bytes_write.sum and gpu__time_duration.sum, averaged over
its captured launches, keyed by the short function name bench.py uses (`k_row_product_tma<4096>` ...).  bench.py reads
the file for `roofline.traffic`; it never hard-codes a number.  Also prints a text summary (copy it under profiles/)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__registers_per_thread"]
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6,
        "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3, "second": 1e6}


def short_name(full):
    """`void gfb::k_row_product_tma<4096>(...)` -> `k_row_product_tma<4096>`; bool template arguments as bench.py prints them"""
    m = re.search(r"(k_[a-z0-9_]+)(<[^>]*>)?", full)
    if not m:
        return full
    targs = m.group(2) or ""
    targs = targs.replace(" ", "")
    return m.group(1) + targs


def main():
    rep = sys.argv[1]
    tag = sys.argv[sys.argv.index("--tag") + 1] if "--tag" in sys.argv else os.path.basename(rep).split(".")[0]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(METRICS)],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, units, body = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(head)}
    acc = {}
    for r in body:
        name = short_name(r[col["Kernel Name"]])
        a = acc.setdefault(name, {"n": 0, "full_name": r[col["Kernel Name"]], "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]})
        a["n"] += 1
        for m in METRICS:
            if m not in col or r[col[m]] == "":
                continue
            v = float(r[col[m]].replace(",", "")) * UNIT.get(units[col[m]], 1.0)
            a[m] = a.get(m, 0.0) + v
    res = {}
    lines = ["ncu --set full, report %s (per launch, averaged over the captured launches)" % os.path.basename(rep)]
    for name, a in acc.items():
        n = a["n"]
        rd, wr, us = a.get(METRICS[0], 0) / n, a.get(METRICS[1], 0) / n, a.get(METRICS[2], 0) / n
        res[name] = {"dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr, "ncu_duration_us": us, "launches_captured": n,
                     "fp64_pipe_pct": a.get(METRICS[5], 0) / n, "warps_active_pct": a.get(METRICS[6], 0) / n, "report": tag}
        lines.append("%-34s grid %-14s block %-12s n=%d  %8.1f us  DRAM read %7.1f MB  write %7.1f MB  (%.0f GB/s)  regs %3.0f  "
                     "SM%% %4.1f  DRAM%% %4.1f  fp64 pipe%% %4.1f  warps active%% %4.1f  smem wavefronts %.1f M"
                     % (name, a["grid"], a["block"], n, us, rd / 1e6, wr / 1e6, (rd + wr) / us / 1e3 if us else 0,
                        a.get(METRICS[8], 0) / n, a.get(METRICS[3], 0) / n, a.get(METRICS[4], 0) / n, a.get(METRICS[5], 0) / n,
                        a.get(METRICS[6], 0) / n, a.get(METRICS[7], 0) / n / 1e6))
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    old = {}
    if os.path.exists(path) and "--merge" in sys.argv:
        old = json.load(open(path))
    old.update(res)
    json.dump(old, open(path, "w"), indent=1, sort_keys=True)
    print("\n".join(lines))


if __name__ == "__main__":
    main()

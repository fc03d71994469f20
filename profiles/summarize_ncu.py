#!/usr/bin/env python
"""Turns an .ncu-rep (ncu --set full) into the short per-kernel summary kept under profiles/.

    python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep > profiles/rNN_<what>.txt
    python profiles/summarize_ncu.py --launches gpurun_out/launches.csv > profiles/rNN_launches.txt
"""
import csv
import io
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block", "launch__waves_per_multiprocessor", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_shared_loads",
    "sass__inst_executed_shared_stores", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warp_latency_issue_stalled_barrier.pct", "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], stdout=subprocess.PIPE,
                         stderr=subprocess.DEVNULL, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    return rows[0], rows[1], rows[2:]


def summarize(rep):
    hdr, units, data = raw_rows(rep)
    col = {h: i for i, h in enumerate(hdr)}
    for r in data:
        print(f"kernel: {r[col['Kernel Name']]}  (launch id {r[col['ID']]})")
        for k in KEYS:
            if k in col:
                print(f"  {k:86s} {r[col[k]]:>16s} {units[col[k]]}")
        rd, wr = col.get("dram__bytes_read.sum"), col.get("dram__bytes_write.sum")
        print()


def launches(path):
    rows = [r for r in csv.reader(open(path)) if r and not r[0].startswith("==")]
    hdr = rows[0]
    ci = {h: i for i, h in enumerate(hdr)}
    agg = OrderedDict()
    total = 0.0
    for r in rows[1:]:
        if len(r) != len(hdr) or r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]].split("(")[0]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        unit = r[ci["Metric Unit"]]
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1.0)
        n, t = agg.get(name, (0, 0.0))
        agg[name] = (n + 1, t + v)
        total += v
    print(f"{'kernel':90s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:90]:90s} {n:8d} {t:12.2f} {t / n:10.2f} {100 * t / total:6.1f}%")
    print(f"{'TOTAL':90s} {sum(n for n, _ in agg.values()):8d} {total:12.2f}")


if __name__ == "__main__":
    if sys.argv[1] == "--launches":
        launches(sys.argv[2])
    else:
        summarize(sys.argv[1])

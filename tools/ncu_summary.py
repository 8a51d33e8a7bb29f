#!/usr/bin/env python
"""Markdown summary of one kernel in an `ncu --set full` report (read with `ncu -i ... --page raw
--csv`) plus the per-source-line stall hot spots (tools/ncu_hotspots.py).

    python tools/ncu_summary.py REPORT.ncu-rep KERNEL_MANGLED_SUBSTR "title / command" > profiles/xxx.md
    python tools/ncu_summary.py ... --traffic-json profiles/das_kernel_traffic.json --workload c2
"""
import argparse
import csv
import io
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

KEYS = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__inst_executed.sum",
]

UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("title")
    ap.add_argument("--top", type=int, default=14)
    ap.add_argument("--outer", default=None, help="file whose lines the hot spots are attributed to")
    ap.add_argument("--traffic-json", default=None)
    ap.add_argument("--workload", default=None)
    args = ap.parse_args()
    out = subprocess.run(["ncu", "-i", args.report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    print(f"# ncu --set full summary: {args.title}\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in col:
            print(f"| {k} | {vals[col[k]]} | {units[col[k]]} |")
    print("\nWarp stall reasons (warps per issue-active cycle):\n\n| reason | value |\n|---|---|")
    st = []
    for h in hdr:
        if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio"):
            st.append((float(vals[col[h]] or 0), h[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]))
    for v, n in sorted(st, reverse=True):
        if v >= 0.01:
            print(f"| {n} | {v:.3f} |")
    print("\nStall samples per source line (tools/ncu_hotspots.py"
          + (f", attributed to the calling line in {args.outer}" if args.outer else "") + "):\n")
    sys.stdout.flush()
    cmd = [sys.executable, os.path.join(HERE, "ncu_hotspots.py"), args.report, args.kernel, "--top", str(args.top)]
    if args.outer:
        cmd += ["--outer", args.outer]
    print(subprocess.run(cmd, capture_output=True, text=True).stdout)
    if args.traffic_json and args.workload:
        def nbytes(k):
            return float(vals[col[k]]) * UNIT_SCALE[units[col[k]]]
        t = nbytes("dram__bytes_read.sum") + nbytes("dram__bytes_write.sum")
        d = {}
        if os.path.exists(args.traffic_json):
            d = json.load(open(args.traffic_json))
        d[args.workload] = t
        d.setdefault("_source", {})[args.workload] = os.path.basename(args.report) + ": " + args.title
        json.dump(d, open(args.traffic_json, "w"), indent=1)


if __name__ == "__main__":
    main()

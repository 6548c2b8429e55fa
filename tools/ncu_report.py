#!/usr/bin/env python
"""Summarise an .ncu-rep: key raw metrics + samples / executed instructions per source line.
usage: python tools/ncu_report.py <file.ncu-rep> [topN]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__registers_per_thread",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_fp64.sum",
        "sm__cycles_elapsed.avg", "sm__cycles_active.avg"]
print("| metric | unit | value |\n|---|---|---|")
for h, u, v in zip(hdr, units, vals):
    if h in keep or (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio")):
        print(f"| {h} | {u} | {v} |")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Line No"][0]
h = rows[hi]
iS, iI, iN, iA = h.index("# Samples"), h.index("Instructions Executed"), h.index("stall_no_inst"), 2
agg = []
for r in rows[hi + 1:]:
    if len(r) > 10 and r[iA] == "-":
        try:
            agg.append((int(r[iS] or 0), int(r[iI] or 0), int(r[iN] or 0), r[0], r[1][:100]))
        except ValueError:
            pass
ts, ti = sum(a[0] for a in agg) or 1, sum(a[1] for a in agg) or 1
print(f"\ntotal samples {ts}, total warp instructions {ti}\n")
for a in sorted(agg, reverse=True)[:top]:
    print(f"{100*a[0]/ts:5.1f}% samp {100*a[1]/ti:5.1f}% inst  noinst {a[2]:6d}  L{a[3]:>4s} {a[4]}")

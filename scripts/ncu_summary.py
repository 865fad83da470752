#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): key raw metrics + per-phase stall attribution.
usage: python scripts/ncu_summary.py gpurun_out/X.ncu-rep [out.txt]"""
import csv, io, subprocess, sys

rep = sys.argv[1]
out = open(sys.argv[2], "w") if len(sys.argv) > 2 else sys.stdout
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__grid_size", "launch__block_size",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warp_latency_per_inst_issued.ratio", "sass__inst_executed_local_stores", "sass__inst_executed_local_loads"]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("kernel:", d.get("Kernel Name"), file=out)
    for k in KEYS:
        if k in d:
            print("  %-90s %s %s" % (k, d[k], units[hdr.index(k)]), file=out)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h = rows[1]
data = [r for r in rows[2:] if len(r) == len(h) and r != h]
iS, iI, iSrc = h.index("# Samples"), h.index("Instructions Executed"), h.index("Source")
cols = {n: i for i, n in enumerate(h)}
bars = [k for k, r in enumerate(data) if "BAR.SYNC" in r[iSrc]]
cut = bars[-1] if bars else len(data)
def summ(lo, hi, name):
    s = sum(int(r[iS]) for r in data[lo:hi]); ins = sum(int(r[iI]) for r in data[lo:hi])
    st = {}
    for key in h:
        if key.startswith("stall_") and "Not Issued" not in key:
            v = sum(int(r[cols[key]]) for r in data[lo:hi])
            if v * 50 > max(s, 1):
                st[key] = v
    f64 = sum(int(r[iI]) for r in data[lo:hi] if any(op in r[iSrc] for op in ("DFMA", "DMUL", "DADD", "MUFU")))
    print("%s: static SASS %d, samples %d, warp-instructions %d, fp64 warp-instructions %d, stalls %s" % (name, hi - lo, s, ins, f64, st), file=out)
print("phase split at the last BAR.SYNC (phase 1 = per-element math, phase 2 = tile reduction):", file=out)
summ(0, cut, "phase1"); summ(cut, len(data), "phase2")
top = sorted(range(len(data)), key=lambda k: -int(data[k][iS]))[:12]
print("top sampled SASS:", file=out)
for k in top:
    print("  %5d %-60s samples %s executed %s" % (k, data[k][iSrc].strip()[:60], data[k][iS], data[k][iI]), file=out)

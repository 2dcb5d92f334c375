#!/usr/bin/env python
"""Key numbers of one ncu capture: ncu_summary.py X.ncu-rep  (time, instructions, pipes, stalls, DRAM traffic)."""
import csv, subprocess, sys
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(out.splitlines()))
h = r[0]; v = r[2] if len(r) > 2 else r[1]
d = dict(zip(h, v))
def g(k):
    try: return float(d[k].replace(",", ""))
    except Exception: return float("nan")
print("kernel", d.get("Kernel Name", "?")[:80])
print("time_us", g("gpu__time_duration.sum"), "grid", d.get("launch__grid_size"), "regs", d.get("launch__registers_per_thread"))
print("warp_inst", g("smsp__inst_executed.sum"), "lanes/inst", g("smsp__thread_inst_executed_per_inst_executed.ratio"))
print("issue_active%", g("smsp__issue_active.avg.pct_of_peak_sustained_active"), "alu%", g("sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
      "fma%", g("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active"), "warps_active/SM", g("sm__warps_active.avg.per_cycle_active"))
print("dram_read", d.get("dram__bytes_read.sum"), d.get("dram__bytes_read.sum.per_second"), "dram_write", d.get("dram__bytes_write.sum"))
st = []
for k in h:
    if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio") and "not_issued" not in k:
        x = g(k)
        if x > 0.1: st.append((x, k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")))
print("stalls", ", ".join(f"{n} {x:.2f}" for x, n in sorted(st, reverse=True)))

#!/usr/bin/env python3
"""Markdown table of the kernels in an ncu report (one row per kernel launch):
   python tools/ncu_summary.py report.ncu-rep [first_id last_id] > profiles/xxx.md"""
import csv, subprocess, sys
rep = sys.argv[1]
METRICS = [("time_ms", "gpu__time_duration.sum"), ("dram_read_GB", "dram__bytes_read.sum"), ("dram_write_GB", "dram__bytes_write.sum"),
           ("dram_pct_peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"), ("l2_hit_pct", "lts__t_sector_hit_rate.pct"),
           ("l1_hit_pct", "l1tex__t_sector_hit_rate.pct"), ("issue_active_pct", "sm__issue_active.avg.pct_of_peak_sustained_elapsed"),
           ("warps_active_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
           ("threads_per_inst", "smsp__thread_inst_executed_per_inst_executed.ratio"), ("regs", "launch__registers_per_thread"),
           ("warp_inst", "smsp__inst_executed.sum"),
           ("stall_long_sb", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio"),
           ("stall_short_sb", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"),
           ("stall_wait", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"),
           ("stall_math", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"),
           ("stall_lg", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio")]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(hdr)}
print("| kernel | " + " | ".join(n for n, _ in METRICS) + " |")
print("|---|" + "---|" * len(METRICS))
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
    vals = []
    for n, m in METRICS:
        v, u = r[ix[m]], units[ix[m]]
        try:
            f = float(v)
            if u == "us": f /= 1e3
            if u == "s": f *= 1e3
            if u == "Mbyte": f /= 1e3
            if u == "Kbyte": f /= 1e6
            if u == "byte": f /= 1e9
            vals.append("%.3g" % f if n != "warp_inst" else "%.3g" % f)
        except ValueError:
            vals.append(v)
    print("| `%s` | " % name + " | ".join(vals) + " |")

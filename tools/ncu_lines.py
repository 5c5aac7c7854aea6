#!/usr/bin/env python3
"""Hot source lines of one kernel from an ncu report captured with --import-source on:
   python tools/ncu_lines.py report.ncu-rep <kernel-id (1-based launch index)> [top]
Aggregates the `--page source --print-source cuda,sass` view per CUDA source line: stall samples, warp instructions
executed and average active threads."""
import csv, subprocess, sys
rep, kid = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-id", ":::" + kid],
                     capture_output=True, text=True).stdout.splitlines()
rows = list(csv.reader(out))
hdr = None
lines = []
for r in rows:
    if len(r) > 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0].isdigit():
        d = dict(zip(hdr, r))
        def num(k):
            try: return float(d[k])
            except Exception: return 0.0
        lines.append((int(r[0]), r[1].strip(), num("# Samples"), num("Instructions Executed"), num("Thread Instructions Executed"),
                      num("stall_long_sb"), num("stall_short_sb"), num("stall_mio"), num("stall_lg"), num("stall_wait"), num("stall_math"),
                      num("L1 Wavefronts Shared Excessive"), num("L2 Theoretical Sectors Local")))
name = [r[1] for r in rows if r and r[0] == "Function Name"]
tot_s = sum(l[2] for l in lines) or 1
tot_i = sum(l[3] for l in lines) or 1
print(name[0] if name else "", "samples", int(tot_s), "warp-inst %.3g" % tot_i)
print("%5s %6s %6s %5s | %6s %6s %6s %6s %6s %6s | %s" % ("line", "smp%", "inst%", "thr", "longsb", "shrtsb", "mio", "lg", "wait", "math", "source"))
for l in sorted(lines, key=lambda x: -x[2])[:top]:
    thr = l[4] / l[3] if l[3] else 0
    print("%5d %6.2f %6.2f %5.1f | %6d %6d %6d %6d %6d %6d | %s" % (l[0], 100 * l[2] / tot_s, 100 * l[3] / tot_i, thr, l[5], l[6], l[7], l[8], l[9], l[10], l[1][:110]))

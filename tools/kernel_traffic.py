#!/usr/bin/env python3
"""DRAM traffic and issue-slot use per kernel from one `ncu --set full` capture of tools/prof_run.py (one search):
   python tools/kernel_traffic.py report.ncu-rep <reads of the run> > profiles/r02_kernel_traffic.json
bench.py scales dram_bytes_per_launch to its own reads per GPU for `roofline.traffic`.  Kernels launched more than
once in the capture (chunks) are summed."""
import csv, json, subprocess, sys
rep, reads = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
ix = {k: i for i, k in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
dram, issue, ms = {}, {}, {}
for r in rows[2:]:
    name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "").split("<")[0]
    b = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        b += float(r[ix[m]]) * scale[units[ix[m]]]
    dram[name] = dram.get(name, 0.0) + b
    t = float(r[ix["gpu__time_duration.sum"]]) * {"us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6}[units[ix["gpu__time_duration.sum"]]]
    ms[name] = ms.get(name, 0.0) + t
    issue[name] = float(r[ix["sm__issue_active.avg.pct_of_peak_sustained_elapsed"]])
json.dump({"source": rep.split("/")[-1], "reads": reads, "dram_bytes_per_launch": dram, "issue_active_pct": issue, "ms_under_ncu": ms}, sys.stdout, indent=1)
print()

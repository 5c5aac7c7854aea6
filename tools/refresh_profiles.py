#!/usr/bin/env python3
"""Rebuild the tracked profile artefacts from one pair of captures brought back in gpurun_out/:
   python tools/refresh_profiles.py <round tag, e.g. r01> <launches.csv> <full.ncu-rep>
   launches.csv : ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ... python bench.py --steps 2 --warmup 3 --no-cpu-baseline
   full.ncu-rep : ncu --set full --clock-control none --import-source on -k regex:"k_qc|k_frames|k_seg|k_probe|k_seed|k_walk|k_gap|k_cls" -s 13 -c 13 ... python tools/prof_run.py 2000000 100 2
Writes profiles/<tag>_launches.csv, <tag>_launch_summary.csv, <tag>_k_probe_traffic.json and replaces the table of <tag>_ncu_summary.md."""
import collections, csv, json, os, re, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag, launches, rep = sys.argv[1], sys.argv[2], sys.argv[3]
P = os.path.join(ROOT, "profiles")
rows = [r for r in csv.reader(open(launches)) if len(r) > 14 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    name = re.sub(r"<.*", "", r[4].split("(")[0].replace("void ", ""))
    if "cub::" in name or "Device" in name:
        name = "cub scan / merge sort / radix sort"
    ns = float(r[14]) * (1e3 if r[13] == "us" else 1.0)
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ns / 1e6
tot = sum(a[1] for a in agg.values())
with open(os.path.join(P, tag + "_launch_summary.csv"), "w") as f:
    f.write("# ncu launch list of `bench.py --steps 2 --warmup 3 --no-cpu-baseline` (2M x 100 bp, 1 B200; 8 searches: 3 warm-up + 2 timed device-resident, 1 + 2 end to end;\n")
    f.write("# k_dpx_bench is the DPX microbenchmark that runs after the timed steps); cold-cache serialised times: compare SHARES with the live stage times, not absolutes\n")
    f.write("kernel,launches,total_ms,share\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("%s,%d,%.3f,%.1f%%\n" % (k, a[0], a[1], 100 * a[1] / tot))
shutil.copy(launches, os.path.join(P, tag + "_launches.csv"))
tab = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
md = os.path.join(P, tag + "_ncu_summary.md")
s = open(md).read()
a, b = s.index("| kernel | time_ms"), s.index("## Reading")
pre = s[a:b]
note = "\n".join(l for l in pre.splitlines() if l.startswith("(Captured"))      # drop stale capture notes
open(md, "w").write(s[:a] + tab + "\n" + s[b:])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", "dram__bytes_read.sum,dram__bytes_write.sum"], capture_output=True, text=True).stdout
rr = list(csv.reader(raw.splitlines()))
ix = {k: i for i, k in enumerate(rr[0])}
for r in rr[2:]:
    if "k_probe" in r[ix["Kernel Name"]]:
        def gb(col):
            v, u = float(r[ix[col]]), rr[1][ix[col]]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        rd, wr = gb("dram__bytes_read.sum"), gb("dram__bytes_write.sum")
        json.dump({"kernel": "k_probe", "dram_bytes_per_launch": rd + wr, "reads_per_launch": 2000000,
                   "source": "profiles/%s_ncu_summary.md (ncu --set full on tools/prof_run.py 2000000 100 2: dram__bytes_read.sum %.2f GB + dram__bytes_write.sum %.2f GB)" % (tag, rd / 1e9, wr / 1e9)},
                  open(os.path.join(P, tag + "_k_probe_traffic.json"), "w"), indent=1)
        break
print(open(os.path.join(P, tag + "_launch_summary.csv")).read())
print(tab)

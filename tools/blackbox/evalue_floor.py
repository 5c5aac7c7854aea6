#!/usr/bin/env python3
"""Measure, per read length, the lowest raw score RAPsearch2 reports for a single-HSP (query, subject) pair
with `-e 1` (the m8 floor used by microbecensus_b200.markers.report_floor).  Needs the reference tree and
baseline/_ref; reads come from the reference's own example.fa.gz (500 bp) and test metagenome."""
import collections, math, os, subprocess, sys, tempfile, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from microbe_census import microbe_census as mc
RAP = "/root/reference/microbe_census/bin/rapsearch_Linux_2.15"
DB = os.path.join(ROOT, "baseline", "_ref", "microbe_census", "data", "rapdb_2.15")
recs = [r.seq for r in mc.parse_seqs(mc.open_file("/root/reference/microbe_census/example/example.fa.gz"))]
meta = [r.seq for r in mc.parse_seqs(mc.open_file("/root/reference/tests/data/metagenome.fa.gz"))][:30000]
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for L in [50, 60, 70, 80, 90, 100, 110, 120, 130, 140, 150, 175, 200, 225, 250, 300, 350, 400, 450, 500]:
        pool = [s[k:k + L] for s in recs for k in range(0, len(s) - L + 1, L)] if L > 100 else [s[:L] for s in meta]
        pool = pool[:30000]
        fa = os.path.join(tmp, "r.fa")
        with open(fa, "w") as fh:
            for i, s in enumerate(pool):
                fh.write(">%d\n%s\n" % (i, s))
        subprocess.check_call("%s -q %s -d %s -o %s -z 8 -e 1 -t n -p f -b 0" % (RAP, fa, DB, os.path.join(tmp, "o")), shell=True,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        pairs = collections.defaultdict(list)
        for line in open(os.path.join(tmp, "o.m8")):
            if line[0] == "#": continue
            f = line.split("\t"); pairs[(f[0], f[1])].append(float(f[11]))
        single = [v[0] for v in pairs.values() if len(v) == 1]
        raw = lambda b: round((b * math.log(2) - math.log(1 / 0.041)) / 0.267)
        lo = min(single)
        out[L] = raw(lo)
        print(L, "reads", len(pool), "lines", sum(len(v) for v in pairs.values()), "min single-HSP bits", lo, "raw", raw(lo), flush=True)
print(out)

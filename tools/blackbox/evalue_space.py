#!/usr/bin/env python3
"""Fit, per read length, the search space N(L) RAPsearch2 v2.15 uses for its E-values:
log10 E = log10(K N) - lambda S log10(e)  (K = 0.041, lambda = 0.267, S = raw score).  Every single-HSP m8 line gives
an interval for log10(K N): its log(e-value) is printed with two decimals, rounded AWAY from zero (with round-to-
nearest the intervals of one run do not intersect; with this rule they do, for every length); the intersection over
a few thousand lines pins log10(K N) to ~1e-4.  Needs the reference tree and baseline/_ref.  Output: the table pasted into
microbecensus_b200/markers.py (LOG10_KN)."""
import collections, math, os, subprocess, sys, tempfile, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from microbe_census import microbe_census as mc
RAP = "/root/reference/microbe_census/bin/rapsearch_Linux_2.15"
DB = os.path.join(ROOT, "baseline", "_ref", "microbe_census", "data", "rapdb_2.15")
LAM, K = 0.267, 0.041
recs = [r.seq for r in mc.parse_seqs(mc.open_file("/root/reference/microbe_census/example/example.fa.gz"))]
meta = [r.seq for r in mc.parse_seqs(mc.open_file("/root/reference/tests/data/metagenome.fa.gz"))][:12000]
out = {}
with tempfile.TemporaryDirectory() as tmp:
    for L in [50, 60, 70, 80, 90, 100, 110, 120, 130, 140, 150, 175, 200, 225, 250, 300, 350, 400, 450, 500]:
        pool = [s[k:k + L] for s in recs for k in range(0, len(s) - L + 1, L)] if L > 100 else [s[:L] for s in meta]
        pool = pool[:12000]
        fa = os.path.join(tmp, "r.fa")
        with open(fa, "w") as fh:
            for i, s in enumerate(pool):
                fh.write(">%d\n%s\n" % (i, s))
        subprocess.check_call("%s -q %s -d %s -o %s -z 8 -e 1 -t n -p f -b 0" % (RAP, fa, DB, os.path.join(tmp, "o")), shell=True,
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        pairs = collections.defaultdict(list)
        for line in open(os.path.join(tmp, "o.m8")):
            if line[0] == "#":
                continue
            f = line.rstrip("\n").split("\t")
            pairs[(f[0], f[1])].append((f[10], float(f[11])))
        lo, hi, n = -1e9, 1e9, 0
        for v in pairs.values():
            if len(v) != 1 or "." not in v[0][0] or len(v[0][0].split(".")[1]) != 2:
                continue                               # sum-statistics lines print six digits: different formula
            loge, bits = float(v[0][0]), v[0][1]
            S = round((bits * math.log(2) - math.log(1 / K)) / LAM)
            c = LAM * S * math.log10(math.e)
            x0, x1 = (loge - 0.01, loge) if loge > 0 else ((loge, loge + 0.01) if loge < 0 else (-0.01, 0.01))
            lo, hi, n = max(lo, x0 + c), min(hi, x1 + c), n + 1
        out[L] = (lo + hi) / 2
        print(L, "lines", n, "log10(K N) in [%.6f, %.6f]" % (lo, hi), "N = %.4g" % (10 ** ((lo + hi) / 2) / K), flush=True)
print("LOG10_KN = {" + ", ".join("%d: %.5f" % kv for kv in sorted(out.items())) + "}")

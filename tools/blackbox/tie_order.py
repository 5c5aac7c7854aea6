#!/usr/bin/env python3
"""Which of several EQUAL-SCORE lines of different subjects does RAPsearch2 v2.15 print first for a read (the line
classify_reads keeps, mc.py:450-453)?  Runs the binary on a FASTA / FASTQ input and scores simple rules on the reads
whose best bit score is shared by several subjects (single-HSP lines only).
usage: tie_order.py <fasta/fastq(.gz)> <read length> [threads].  Result on tests/data/metagenome.fa.gz at 100 bp: 643
such reads of 3,273 with hits; lowest subject index first in 62 % (reads with <= 16 lines) / 37 % (more lines); highest
identity, leftmost query start, shortest subject: 52-61 %; the order does not depend on -z.  DESIGN.md section 9 has the
mechanism read off the binary (multimap by subject index, then an unstable std::sort over printed and unprinted hits)."""
import collections, os, subprocess, sys, warnings
warnings.filterwarnings("ignore")
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/baseline/_ref")
from microbe_census import microbe_census as mc
from microbecensus_b200.markers import Markers
src, L = sys.argv[1], int(sys.argv[2])
z = int(sys.argv[3]) if len(sys.argv) > 3 else 8
RAP = "/root/reference/microbe_census/bin/rapsearch_Linux_2.15"
DB = "/root/repo/baseline/_ref/microbe_census/data/rapdb_2.15"
tmp = "/tmp/tie_order"; os.makedirs(tmp, exist_ok=True)
seqs = [r.seq[:L] for r in mc.parse_seqs(mc.open_file(src)) if len(r.seq) >= L]
with open(os.path.join(tmp, "r.fa"), "w") as fh:
    for i, s in enumerate(seqs):
        fh.write(">%d\n%s\n" % (i, s))
subprocess.check_call("%s -q %s/r.fa -d %s -o %s/o -z %d -e 1 -t n -p f -b 0" % (RAP, tmp, DB, tmp, z), shell=True,
                      stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
m = Markers(); idx = {n: i for i, n in enumerate(m.names)}
d = collections.OrderedDict()
for l in open(os.path.join(tmp, "o.m8")):
    if l[0] != "#":
        f = l.rstrip("\n").split("\t"); d.setdefault(f[0], []).append(f)
single = lambda f: len(f[10].split(".")[-1]) <= 2          # two-decimal E-values: single-HSP lines
rules = {"lowest subject index": lambda g: min(g, key=lambda f: idx[f[1]]),
         "highest subject index": lambda g: max(g, key=lambda f: idx[f[1]]),
         "highest identity": lambda g: max(g, key=lambda f: (float(f[2]), -idx[f[1]])),
         "leftmost query start": lambda g: min(g, key=lambda f: (min(int(f[6]), int(f[7])), idx[f[1]])),
         "shortest subject": lambda g: min(g, key=lambda f: (m.subj_len[idx[f[1]]], idx[f[1]]))}
for small in (True, False):
    tot, ok = 0, collections.Counter()
    for q, lines in d.items():
        best = max(float(f[11]) for f in lines)
        grp = [f for f in lines if float(f[11]) == best]
        if len(set(f[1] for f in grp)) < 2 or not all(single(f) for f in grp) or (len(lines) <= 16) != small:
            continue
        tot += 1
        for name, rule in rules.items():
            ok[name] += rule(grp)[1] == grp[0][1]
    print("reads with <= 16 lines" if small else "reads with > 16 lines", tot, {k: round(v / max(tot, 1), 3) for k, v in ok.items()})

#!/usr/bin/env python3
"""Small search for compute-sanitizer: golden reads at 100 / 150 / 500 bp, with and without -d / qualities.
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import golden_io
from microbecensus_b200.engine import MarkerSearch, ReadBatch
eng = MarkerSearch()
for name, L in (("meta.fa.gz", 100), ("long.fa.gz", 150), ("long.fa.gz", 500)):
    seqs = golden_io.read_fasta(name)[:200]
    eng.set_params(L, filter_dups=(L == 150))
    eng.push(ReadBatch.from_strings(seqs))
    res = eng.search(-1)
    print(name, L, res.sampled_reads, res.reads_classified, len(eng.hits()))
recs = golden_io.read_fastq("short.fq.gz")[:300]
eng.set_params(100, quality_offset=33, min_quality=5, mean_quality=20, max_unknown=5)
eng.push(ReadBatch.from_strings([r[1] for r in recs], [r[2] for r in recs]))
res = eng.search(100)
print("fastq", res.sampled_reads, res.low_qual, res.reads_classified)

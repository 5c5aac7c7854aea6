#!/usr/bin/env python3
"""One search of N synthetic reads (for ncu): python tools/prof_run.py [n_reads] [L] [repeats]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microbecensus_b200 import synth, _lib
if os.environ.get("MCX_LIB"):                 # A/B runs against another build of the library
    _lib.LIB_PATH = os.environ["MCX_LIB"]
from microbecensus_b200.engine import MarkerSearch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 2
eng = MarkerSearch()
eng.set_params(L)
batch = synth.reads(2, 0, n, L)
for _ in range(rep):
    eng.push(batch)
    res = eng.search(-1)
print(res.sampled_reads, res.reads_classified, eng.timings())

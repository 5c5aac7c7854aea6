#!/usr/bin/env python3
"""Latency of small batches (what a streamed run with tiny files sees): python tools/small_batch.py [n_reads] [L] [repeats]
prints the wall time and the stage timers of the LAST of `repeats` push + search rounds (buffers allocated, tables warm)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microbecensus_b200 import synth
from microbecensus_b200.engine import MarkerSearch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 300
L = int(sys.argv[2]) if len(sys.argv) > 2 else 100
rep = int(sys.argv[3]) if len(sys.argv) > 3 else 5
eng = MarkerSearch()
eng.set_params(L)
batch = synth.reads(2, 0, n, L)
for _ in range(rep):
    t0 = time.perf_counter()
    eng.push(batch)
    res = eng.search(-1)
    wall = time.perf_counter() - t0
tm, launches = eng.timings()
print("%d reads x %d bp: %.3f ms wall per push + search, %d launches, stages %s" % (n, L, wall * 1e3, launches, {k: round(v, 3) for k, v in tm.items()}))

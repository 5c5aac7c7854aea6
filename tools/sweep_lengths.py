#!/usr/bin/env python3
"""BASELINE config 5 on one GPU: reads/s of the whole search at every supported read length.
python tools/sweep_lengths.py [reads_per_length] > gpurun_out/length_sweep.jsonl"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from microbecensus_b200 import synth, microbe_census as mcb
from microbecensus_b200.engine import MarkerSearch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
eng = MarkerSearch()
for L in mcb.VALID_LENGTHS:
    batch = synth.reads(5, 0, n, L)
    eng.set_params(L)
    for _ in range(2):
        eng.push(batch); eng.search(-1)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        eng.push(batch); res = eng.search(-1)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    tm, _ = eng.timings()
    ags = mcb.estimate_average_genome_size({"read_length": L, "sampled_reads": res.sampled_reads, "verbose": False}, None, res.agg_hits())
    print(json.dumps({"read_length": L, "reads": n, "reads_per_s_e2e": n / dt, "ms": dt * 1e3, "stages_ms": tm, "reads_with_hits": res.reads_with_hits,
                      "reads_classified": res.reads_classified, "n_hsp": res.n_hsp, "gapped_cells": res.gapped_cells, "ags": ags}), flush=True)

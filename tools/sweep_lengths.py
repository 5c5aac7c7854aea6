#!/usr/bin/env python3
"""BASELINE config 5: reads/s of the whole search at every supported read length (-l 50 .. 500), sharded over the GPUs of
the box.   python tools/sweep_lengths.py [reads_per_length_total]        (1 GPU)
           torchrun --nproc-per-node 8 tools/sweep_lengths.py 20000000   (8 GPUs, 2.5M reads per GPU and length)
Every rank draws its share of 500 bp reads once and keeps them on its GPU; the reads of length L are their first L bases
(a read of L bp starting at the same genome position), so nothing but the search runs per length.  Device-resident
timing (CUDA events, max over ranks), ASCII reads packed on the device by k_pack_ascii inside the timed region."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from microbecensus_b200 import synth, microbe_census as mcb
from microbecensus_b200.engine import MarkerSearch

total = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n = total // world
eng = MarkerSearch(device=local)
eng.set_stream(torch.cuda.current_stream().cuda_stream)
t0 = time.time()
pool = synth.reads(5, rank * n, (rank + 1) * n, 500)
d_pool = torch.from_numpy(pool.bases.reshape(n, 500)).to(dev)
gen_s = time.time() - t0
del pool
LENGTHS = [int(x) for x in os.environ['SWEEP_LENGTHS'].split(',')] if os.environ.get('SWEEP_LENGTHS') else mcb.VALID_LENGTHS
for L in LENGTHS:
    d_b = d_pool[:, :L].contiguous()
    d_o = torch.arange(n + 1, dtype=torch.int64, device=dev) * L
    eng.set_params(L)
    def step():
        eng.push_device(d_b.data_ptr(), 0, d_o.data_ptr(), n, n * L)
        return eng.search(-1)
    step()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 2
    e0.record(); w0 = time.perf_counter()
    for _ in range(reps):
        res = step()
    e1.record(); torch.cuda.synchronize()
    ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3) / reps
    tm, _ = eng.timings()
    v = torch.from_numpy(res.counts_vector()).to(dev)
    t = torch.tensor([ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(v); dist.all_reduce(t, op=dist.ReduceOp.MAX)
    res.load_counts_vector(v.cpu().numpy())
    ms = float(t[0])
    if rank == 0:
        ags = mcb.estimate_average_genome_size({"read_length": L, "sampled_reads": res.sampled_reads, "verbose": False}, None, res.agg_hits())
        print(json.dumps({"read_length": L, "n_gpus": world, "reads": res.sampled_reads, "reads_per_s": res.sampled_reads / (ms * 1e-3), "ms": ms,
                          "stages_ms_rank0": tm, "reads_with_hits": res.reads_with_hits, "reads_classified": res.reads_classified, "n_hsp": res.n_hsp,
                          "gapped_cells": res.gapped_cells, "gapped_gcups_per_gpu": res.gapped_cells / world / (tm["gapped"] * 1e-3) / 1e9 if tm["gapped"] else None,
                          "ags": ags, "genome_pack": os.path.basename(synth.genome_pack_path()), "host_generation_s": gen_s}), flush=True)
if world > 1:
    dist.destroy_process_group()

#!/usr/bin/env python3
"""torchrun --nproc-per-node N tools/check_sharded.py : the sharded search (-n, -d across ranks, all-reduce) must give
exactly the numbers of one GPU searching everything."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from microbecensus_b200 import synth
from microbecensus_b200.engine import MarkerSearch, ReadBatch
from microbecensus_b200.distributed import sharded_search

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
n, L = 400000, 100
base = synth.reads(9, 0, n, L, with_quals=True)
# inject duplicates: 5 % exact copies and 1 % reverse complements of earlier reads
rng = np.random.default_rng(3)
b = base.bases.reshape(n, L).copy(); q = base.quals.reshape(n, L).copy()
comp = np.zeros(256, np.uint8); comp[[65, 67, 71, 84, 78]] = [84, 71, 67, 65, 78]
for i in np.flatnonzero(rng.random(n) < 0.06):
    if i < 10: continue
    j = int(rng.integers(0, i))
    b[i] = b[j] if rng.random() < 0.83 else comp[b[j][::-1]]
whole = ReadBatch(b.reshape(-1), base.offsets, q.reshape(-1))
eng = MarkerSearch(device=local)
ok = True
for nreads, dups in ((None, False), (250000, False), (None, True), (300000, True)):
    eng.set_params(L, quality_offset=33, min_quality=5, mean_quality=20, max_unknown=5, filter_dups=False)
    lo, hi = rank * n // world, (rank + 1) * n // world
    res = sharded_search(eng, whole.slice(lo, hi), lo, nreads=nreads, filter_dups=dups)
    if rank == 0:
        eng.set_params(L, quality_offset=33, min_quality=5, mean_quality=20, max_unknown=5, filter_dups=dups)
        eng.push(whole)
        ref = eng.search(-1 if nreads is None else nreads)
        same = np.array_equal(ref.counts_vector(), res.counts_vector())
        print("nreads", nreads, "dups", dups, "sampled", res.sampled_reads, "dups", res.dups, "low_qual", res.low_qual,
              "classified", res.reads_classified, "EQUAL" if same else "DIFFERENT", flush=True)
        ok &= same
    dist.barrier()
# ---- run_pipeline under torchrun: the ranks walk the files together in batches, every rank keeps its own; -n, -d (with the
# owners remembering the reads kept in earlier rounds) and the sums must give what one GPU reading everything gives
import tempfile
from microbecensus_b200 import microbe_census as mcb
tmp = os.environ.get("TMPDIR", "/tmp")
half = n // 2
paths = [os.path.join(tmp, "shard_check_%d.fq" % k) for k in (1, 2)]
if rank == 0:
    synth.write_fastq(ReadBatch(b[:half].reshape(-1), base.offsets[:half + 1], q[:half].reshape(-1)), paths[0])
    synth.write_fastq(ReadBatch(b[half:].reshape(-1), base.offsets[:n - half + 1], q[half:].reshape(-1)), paths[1])
dist.barrier()
os.environ["MCX_BATCH_READS"] = "30000"
for nreads, dups in ((None, False), (300000, True), (None, True)):
    args = {"seqfiles": list(paths), "verbose": False, "nreads": nreads, "read_length": L, "threads": 4, "filter_dups": dups,
            "min_quality": 5, "mean_quality": 20, "max_unknown": 5}
    est, out = mcb.run_pipeline(args)
    if rank == 0:
        eng.set_params(L, quality_offset=out["quality_offset"], min_quality=5, mean_quality=20, max_unknown=5, filter_dups=dups)
        eng.push(whole)
        ref = eng.search(-1 if nreads is None else nreads)
        want = mcb.estimate_average_genome_size({"read_length": L, "sampled_reads": ref.sampled_reads, "verbose": False}, None, ref.agg_hits())
        same = est == want and out["sampled_reads"] == ref.sampled_reads
        print("run_pipeline nreads", nreads, "dups", dups, "sampled", out["sampled_reads"], "AGS", est, "EQUAL" if same else "DIFFERENT (want %s, %s)" % (ref.sampled_reads, want), flush=True)
        ok &= same
    dist.barrier()
if rank == 0:
    print("SHARDED OK" if ok else "SHARDED MISMATCH")
dist.destroy_process_group()

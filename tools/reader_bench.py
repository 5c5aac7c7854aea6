#!/usr/bin/env python3
"""Throughput of the file reader (libmcxio) and of run_pipeline from a file: python tools/reader_bench.py [reads] [L]
Writes a synthetic FASTQ (plain and .gz) to $TMPDIR, parses it with 1 .. nproc threads, then runs the whole drop-in on it."""
import json, os, sys, time, gzip, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from microbecensus_b200 import synth, seqio, microbe_census as mcb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000000
L = int(sys.argv[2]) if len(sys.argv) > 2 else 150
tmp = os.environ.get("TMPDIR", "/tmp")
path = os.path.join(tmp, "reader_bench.fq")
b = synth.reads(3, 0, n, L, with_quals=True)
synth.write_fastq(b, path)
out = {"reads": n, "read_length": L, "file_MB": os.path.getsize(path) / 1e6, "cores": os.cpu_count(), "plain": {}, "gz": {}}
for threads in sorted({1, 2, 4, 8, 16, os.cpu_count() or 1}):
    if threads > (os.cpu_count() or 1):
        continue
    best = 0
    for rep in range(2):
        t0 = time.perf_counter()
        with seqio.SeqFile(path) as rd:
            k = 0
            while not rd.eof:
                p = rd.next_packed(4000000, threads); k += p.n; del p
        best = max(best, k / (time.perf_counter() - t0))
    out["plain"][threads] = best
if len(sys.argv) <= 3:
    sub = os.path.join(tmp, "reader_bench_small.fq.gz")
    with open(path, "rb") as fi, gzip.open(sub, "wb", compresslevel=4) as fo:
        fo.write(fi.read(200 << 20))
    t0 = time.perf_counter()
    with seqio.SeqFile(sub) as rd:
        k = 0
        while not rd.eof:
            p = rd.next_packed(4000000, 8); k += p.n; del p
    out["gz"] = {"reads_per_s": k / (time.perf_counter() - t0), "reads": k}
try:
    import torch
    if torch.cuda.is_available():
        for threads in (1, os.cpu_count() or 1):
            args = {"seqfiles": [path], "verbose": False, "nreads": None, "read_length": L, "threads": threads, "min_quality": 5, "mean_quality": 20, "max_unknown": 5}
            mcb.run_pipeline(dict(args))                      # warm-up: index upload, buffers
            t0 = time.perf_counter()
            est, o = mcb.run_pipeline(dict(args))
            dt = time.perf_counter() - t0
            out["run_pipeline_t%d" % threads] = {"seconds": dt, "reads_per_s": n / dt, "sampled": o["sampled_reads"], "ags": est}
except ImportError:
    pass
print(json.dumps(out))

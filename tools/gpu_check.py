#!/usr/bin/env python3
"""Quick on-GPU parity run: CUDA path vs the CPU oracle on the golden reads (bit-exact expected)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from microbecensus_b200.markers import Markers, report_floor
from microbecensus_b200.engine import MarkerSearch, ReadBatch
from oracle_lib import Oracle, MCX_ORDER
import golden_io

m = Markers()
t0 = time.time(); ms = MarkerSearch(m); print("create %.2fs" % (time.time() - t0))
orc = Oracle(m)
ok = True
for name, L in (("meta.fa.gz", 100), ("meta50.fa.gz", 50), ("long.fa.gz", 500), ("long.fa.gz", 250), ("long.fa.gz", 150)):
    seqs = golden_io.read_fasta(name)
    batch = ReadBatch.from_strings(seqs)
    ms.set_params(L)
    qc = ms.push(batch)
    res = ms.search(-1)
    hits = ms.hits()
    oh, nseeds = orc.search(batch, L, ms.min_report_raw)
    oh = oh[:, MCX_ORDER]
    same = hits.shape == oh.shape and np.array_equal(hits, oh)
    oc = orc.classify(orc.search(batch, L, ms.min_report_raw)[0], L, m, batch.n)
    cls_same = (np.array_equal(oc["fam_hits"], res.fam_hits) and np.array_equal(oc["fam_aln"], res.fam_aln)
                and np.array_equal(oc["aln_by_len"], res.aln_by_len) and oc["classified"] == res.reads_classified
                and np.array_equal(oc["best_subject"], ms.classified(batch.n)))
    print(name, L, "qc", qc, "hsp gpu", len(hits), "oracle", len(oh), "hits_equal", same, "classify_equal", cls_same,
          "classified", res.reads_classified, "with_hits", res.reads_with_hits, "seed_hits", res.n_seed_hits,
          "gapped", res.n_gapped, "cells", res.gapped_cells, ms.timings())
    if not same:
        ok = False
        a = set(map(tuple, hits.tolist())); b = set(map(tuple, oh.tolist()))
        print("  gpu-only", sorted(a - b)[:5]); print("  oracle-only", sorted(b - a)[:5])
    ok &= cls_same
print("ALL OK" if ok else "MISMATCH")
sys.exit(0 if ok else 1)

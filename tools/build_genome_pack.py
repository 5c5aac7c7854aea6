#!/usr/bin/env python3
"""Pack a subset of the reference's training genomes (training/input/genomes/*.fna.gz) into 2 bits per
base for the synthetic read generator (microbecensus_b200/synth.py).  Ten single-contig, pure-ACGT
genomes (~17.5 Mbp) are enough to give genome-like seed statistics and ~1 % marker-gene reads; shipping
all 30 (85 Mbp) would add 21 MB to the repository for no change in what the benchmark exercises.

layout: char[8] "MCXGEN01"; int32 n; int64 length[n]; uint8 packed[ceil(sum/4)]  (A0 C1 G2 T3, base i in
bits 2*(i%4) of byte i/4)
"""
import gzip, os, struct, sys
import numpy as np

GENOMES = ["2504756006", "2513237181", "638154521", "639633019", "641522611", "642555132", "649633067",
           "2511231212", "640753014", "641228511"]


def main(ref="/root/reference"):
    out = os.path.join(os.path.dirname(__file__), "..", "microbecensus_b200", "data", "genomes.pack")
    code = np.full(256, 255, np.uint8)
    for i, c in enumerate("ACGT"):
        code[ord(c)] = i
    seqs = []
    for g in GENOMES:
        s = "".join(l.strip() for l in gzip.open(os.path.join(ref, "training", "input", "genomes", g + ".fna.gz"), "rt") if l[0] != ">")
        a = code[np.frombuffer(s.encode(), np.uint8)]
        assert a.max() < 4, g
        seqs.append(a)
    allb = np.concatenate(seqs)
    pad = (-len(allb)) % 4
    allb = np.concatenate([allb, np.zeros(pad, np.uint8)]).reshape(-1, 4)
    packed = (allb[:, 0] | (allb[:, 1] << 2) | (allb[:, 2] << 4) | (allb[:, 3] << 6)).astype(np.uint8)
    with open(out, "wb") as fh:
        fh.write(b"MCXGEN01" + struct.pack("<i", len(seqs)) + np.array([len(s) for s in seqs], np.int64).tobytes() + packed.tobytes())
    print("wrote", out, sum(len(s) for s in seqs), "bp", os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main(*sys.argv[1:])

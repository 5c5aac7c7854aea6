#!/usr/bin/env python3
"""Pack the reference's training genomes (training/input/genomes/*.fna.gz) into 2 bits per base for the
synthetic read generator (microbecensus_b200/synth.py).

  python tools/build_genome_pack.py          ten single-contig, pure-ACGT genomes (17.5 Mbp) -> data/genomes.pack
                                             (4.4 MB, committed: what a checkout without the reference tree has)
  python tools/build_genome_pack.py --all    all 30 genomes (84.8 Mbp; every contig, cut at letters other than ACGT,
                                             pieces under 1,000 bp dropped) -> data/genomes30.pack (21 MB, built by
                                             __graft_entry__.build() when /root/reference is mounted, kept out of git;
                                             synth.py prefers it).  100 M x 150 bp reads drawn from 17.5 Mbp repeat
                                             themselves (70 % chance duplicates under -d); from 84.8 Mbp they do not.

layout: char[8] "MCXGEN01"; int32 n; int64 length[n]; uint8 packed[ceil(sum/4)]  (A0 C1 G2 T3, base i in
bits 2*(i%4) of byte i/4)
"""
import gzip, os, struct, sys
import numpy as np

GENOMES = ["2504756006", "2513237181", "638154521", "639633019", "641522611", "642555132", "649633067",
           "2511231212", "640753014", "641228511"]


def main(ref="/root/reference", all_genomes=False):
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "microbecensus_b200", "data", "genomes30.pack" if all_genomes else "genomes.pack")
    code = np.full(256, 255, np.uint8)
    for i, c in enumerate("ACGT"):
        code[ord(c)] = i
    seqs = []
    gdir = os.path.join(ref, "training", "input", "genomes")
    if all_genomes:
        for fn in sorted(os.listdir(gdir)):
            if not fn.endswith(".fna.gz"):
                continue
            contigs, cur = [], []
            for l in gzip.open(os.path.join(gdir, fn), "rt"):
                if l[0] == ">":
                    if cur:
                        contigs.append("".join(cur))
                    cur = []
                else:
                    cur.append(l.strip())
            if cur:
                contigs.append("".join(cur))
            for c in contigs:
                a = code[np.frombuffer(c.encode(), np.uint8)]
                bad = np.flatnonzero(a > 3)
                cuts = np.concatenate([[-1], bad, [len(a)]])
                for lo, hi in zip(cuts[:-1] + 1, cuts[1:]):
                    if hi - lo >= 1000:
                        seqs.append(a[lo:hi])
    for g in ([] if all_genomes else GENOMES):
        s = "".join(l.strip() for l in gzip.open(os.path.join(gdir, g + ".fna.gz"), "rt") if l[0] != ">")
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
    args = [a for a in sys.argv[1:] if a != "--all"]
    main(*args, all_genomes="--all" in sys.argv[1:])

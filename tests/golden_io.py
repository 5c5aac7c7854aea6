"""Readers for tests/golden fixtures (plain Python; the product's reader is tested separately)."""
import gzip
import json
import os

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def read_fasta(name):
    seqs, cur = [], None
    for line in gzip.open(os.path.join(GOLD, name), "rt"):
        line = line.rstrip("\n")
        if line.startswith(">"):
            if cur is not None:
                seqs.append("".join(cur))
            cur = []
        else:
            cur.append(line)
    if cur is not None:
        seqs.append("".join(cur))
    return seqs


def read_fastq(name):
    lines = gzip.open(os.path.join(GOLD, name), "rt").read().split("\n")
    recs = []
    for i in range(0, len(lines) - 3, 4):
        recs.append((lines[i][1:], lines[i + 1], lines[i + 3]))
    return recs


def read_m8(name):
    rows = []
    for line in gzip.open(os.path.join(GOLD, name), "rt"):
        if line[0] == "#":
            continue
        f = line.rstrip("\n").split("\t")
        rows.append((int(f[0]), f[1], float(f[2]), int(f[3]), int(f[4]), int(f[5]), int(f[6]), int(f[7]),
                     int(f[8]), int(f[9]), float(f[10]), float(f[11])))
    return rows


def read_json(name):
    return json.load(open(os.path.join(GOLD, name)))

#!/usr/bin/env python3
"""False-positive rate of libmcx's blocked presence filter (mcx.cu: filt_hash_a / filt_hash_b / filt_bits) on the marker
set, in numpy, with the hashes and bit layout of the kernel: python tools/filter_sim.py [log2 blocks = 21] [reads = 3000]
Queries = every 10-letter murphy10 window of the six SEG-masked frames of synthetic 150 bp reads (tests/oracle_lib for the
frames).  Also checks that the filter has no false negatives.  Results (3,000 reads, 744,000 windows): 2^21 blocks 1.2 %
false positives (pass rate 4.2 % with 3.0 % true word hits), 2^20 blocks 2.1 %, 2^19 blocks 3.8 %."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from microbecensus_b200 import markers, synth
import oracle_lib

NB = int(sys.argv[1]) if len(sys.argv) > 1 else 21
NREADS = int(sys.argv[2]) if len(sys.argv) > 2 else 3000
U = np.uint64
M = markers.Markers()
MUR = np.array([0, 1, 2, 2, 3, 2, 2, 4, 5, 6, 6, 1, 6, 7, 8, 9, 9, 7, 7, 6, 10], np.uint8)


def windows(red):                         # letter k of the window at bits 4k
    n = len(red)
    pad = np.concatenate([red.astype(U), np.full(10, 15, U)])
    w = np.zeros(n, U)
    for k in range(10):
        w |= pad[k:k + n] << U(4 * k)
    return w


def bad(w, nibs):
    b = np.zeros(len(w), bool)
    for k in nibs:
        b |= ((w >> U(4 * k)) & U(15)) >= 10
    return b


def mask(nibs):
    return U(sum(15 << (4 * k) for k in nibs))


PAT = [list(range(9))] + [[k for k in range(10) if k != w] for w in (3, 4, 5, 6)]
m32 = lambda x: x & U(0xffffffff)


def hash_a(w):
    lo, hi = m32(w), w >> U(32)
    h = m32((lo & U(0xFFF00FFF)) * U(0x9E3779B1)); h ^= h >> U(15)
    h = m32((h + (hi & U(0xF)) * U(0x85EBCA77)) * U(0xC2B2AE3D)); h ^= h >> U(13)
    return h


def hash_b(w):
    lo, hi = m32(w), w >> U(32)
    h = m32((lo & U(0xF00FFFFF)) * U(0x9E3779B1)); h ^= h >> U(15)
    h = m32((h + (hi & U(0xFF)) * U(0x85EBCA77) + U(0x68E31DA4)) * U(0xC2B2AE3D)); h ^= h >> U(13)
    return h


nib = lambda w, k: (w >> U(4 * k)) & U(15)


def bits(w, p, h):
    free = [nib(w, 3) | (nib(w, 4) << U(4)), nib(w, 4) | (nib(w, 9) << U(4)), nib(w, 3) | (nib(w, 9) << U(4)), nib(w, 6), nib(w, 5)][p]
    x = m32(((free | U((p + 1) << 8)) ^ m32(h << U(11))) * U(0x2C1B3C6D))
    return (x >> U(27)).astype(np.int64), ((x >> U(22)) & U(31)).astype(np.int64)


def locate(w, p):
    h = hash_a(w) if p <= 2 else hash_b(w)
    b1, b2 = bits(w, p, h)
    return (h >> U(32 - NB)).astype(np.int64), b1, b2


# the words of the database
res = M.res.copy(); res[res > 20] = 20
red = MUR[res]
subj = np.repeat(np.arange(M.n_subj), np.diff(M.off))
remain = M.off[subj + 1] - np.arange(len(red))
dbw = windows(red)
keys = []
for p, nibs in enumerate(PAT):
    ok = (remain >= (9 if p == 0 else 10)) & ~bad(dbw, nibs)
    keys.append(np.unique(dbw[ok] & mask(nibs)))
A = np.zeros((1 << NB, 4, 32), bool); B = np.zeros((1 << NB, 2, 32), bool)
for p in range(5):
    blk, b1, b2 = locate(keys[p], p)
    if p == 0:
        A[blk, 0, b1] = True; A[blk, 3, b2] = True
    elif p <= 2:
        A[blk, p, b1] = True; A[blk, p, b2] = True
    else:
        B[blk, p - 3, b1] = True; B[blk, p - 3, b2] = True
# the query windows
orc = oracle_lib.Oracle(M)
L = 150
bases = synth.reads(3, 0, NREADS, L).bases.reshape(-1, L)
qs = []
for i in range(bases.shape[0]):
    s = bases[i].tobytes().decode()
    for f in range(6):
        aa = np.frombuffer(bytes(orc.frame(s, L, f, True)), np.uint8)
        w = windows(MUR[np.minimum(aa, 20)])
        qs.append(w[:max(len(aa) - 8, 0)])
q = np.concatenate(qs)
tot = passed = true = 0
for p in range(5):
    valid = ~bad(q, PAT[p])
    truth = np.isin(q & mask(PAT[p]), keys[p]) & valid
    blk, b1, b2 = locate(q, p)
    ok = (A[blk, 0, b1] & A[blk, 3, b2]) if p == 0 else (A[blk, p, b1] & A[blk, p, b2]) if p <= 2 else (B[blk, p - 3, b1] & B[blk, p - 3, b2])
    ok &= valid
    assert not (truth & ~ok).any(), "false negative"
    tot += valid.sum(); passed += ok.sum(); true += truth.sum()
    print("pattern %d: %d distinct words, false positives %.4f" % (p, len(keys[p]), (ok & ~truth).sum() / valid.sum()))
print("2^%d blocks (%d MB), %d windows: pass %.4f, true %.4f, false positives %.4f" % (NB, (24 << NB) >> 20, len(q), passed / tot, true / tot, (passed - true) / tot))

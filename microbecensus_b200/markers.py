"""Marker database blob (``data/markers.mcxdb.gz``, written by tools/build_marker_db.py).

One file replaces the reference's data/rapdb_2.15 (RAPsearch2 database), gene_fam.map, gene_len.map,
pars.map, coefficients.map, weights.map and read_len.map (loaders: find_opt_pars / read_dic,
microbe_census.py:61-88).
"""
import gzip
import math
import os
import struct

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PATH = os.path.join(_HERE, "data", "markers.mcxdb.gz")
STAT_NAMES = ("hits", "cov", "aln")

# Raw-score floor of the HSPs RAPsearch2 reports with `-e 1` (log10 E <= 1) for single-HSP hits, by read
# length: measured from its output at each of the 20 supported lengths (tools/blackbox/evalue_floor.py;
# E = K * space * exp(-lambda * S) with lambda 0.267, K 0.041 and RAPsearch2's length-adjusted search space).
# The floor only decides which HSPs count as "reported"; every family cutoff of pars.map lies at or above it.
_REPORT_FLOOR = {50: 47, 60: 48, 400: 50, 450: 51, 500: 52}

# log10(K * search space) of RAPsearch2's E-values by read length: log10 E = LOG10_KN[L] - 0.267 S log10(e).
# Fitted from the binary's own output (tools/blackbox/evalue_space.py): every m8 line bounds the constant to an
# interval 0.01 wide, a few thousand lines pin it to ~1e-4.  The search space is not the textbook (m - l)(n - N l);
# it is flat (~1.01e8) from 80 to 350 bp and the same for all six frames of a read.  The floors above follow from it:
# smallest S with log10 E <= 1.
LOG10_KN = {50: 6.42177, 60: 6.51877, 70: 6.57940, 80: 6.61364, 90: 6.60555, 100: 6.61851, 110: 6.61234, 120: 6.62247,
            130: 6.61617, 140: 6.60965, 150: 6.61891, 175: 6.61844, 200: 6.61687, 225: 6.61174, 250: 6.60781,
            300: 6.60900, 350: 6.62083, 400: 6.75603, 450: 6.85623, 500: 6.93852}


def log10_evalue(raw, read_length):
    """log10 of the E-value RAPsearch2 assigns to a single HSP of raw score `raw` at this read length."""
    return LOG10_KN[int(read_length)] - 0.267 * raw * math.log10(math.e)


def log10_evalue_printed(raw, read_length):
    """... as its m8 shows it: two decimals, rounded away from zero, except that values between -0.01 and 0 print as 0
    (observed on ~200,000 lines at all 20 lengths: positive values are rounded up, negative ones down, and the three
    (length, score) pairs with -0.01 < log10 E < 0 -- S = 57 at 90, 250 and 300 bp -- print "0")."""
    h = log10_evalue(raw, read_length) * 100.0
    if h > 0:
        return math.ceil(h - 1e-9) / 100.0
    t = math.trunc(h + 1e-9)
    return (t - 1) / 100.0 if t < 0 else 0.0


def bits_printed(raw):
    """Bit score as RAPsearch2 prints it: (0.267 S + ln(1/0.041)) / ln 2 with two decimals."""
    return float("%.2f" % ((0.267 * raw + math.log(1.0 / 0.041)) / math.log(2.0)))


def min_raw_for_bits(cutoff):
    """Smallest raw score whose printed bit score is >= cutoff (mc.py:425 compares the printed text)."""
    s = 1
    while bits_printed(s) < cutoff:
        s += 1
    return s


def report_floor(read_length):
    return _REPORT_FLOOR.get(int(read_length), 49)


class Markers:
    def __init__(self, path=None):
        path = path or DEFAULT_PATH
        opener = gzip.open if path.endswith(".gz") else open
        with opener(path, "rb") as fh:
            blob = fh.read()
        if blob[:8] != b"MCXDB001":
            raise ValueError("not a marker blob: %s" % path)
        n_subj, n_res, n_fam, n_len, names_bytes = struct.unpack_from("<5i", blob, 8)
        p = 8 + 32
        self.off = np.frombuffer(blob, np.int32, n_subj + 1, p).copy(); p += 4 * (n_subj + 1)
        self.fam = np.frombuffer(blob, np.uint8, n_subj, p).copy(); p += (n_subj + 3) & ~3
        self.res = np.frombuffer(blob, np.uint8, n_res, p).copy(); p += (n_res + 3) & ~3
        self.read_lengths = [int(x) for x in np.frombuffer(blob, np.int32, n_len, p)]; p += 4 * n_len
        self.fam_names = [blob[p + 8 * i:p + 8 * i + 8].rstrip(b"\0").decode() for i in range(n_fam)]; p += 8 * n_fam
        rec = np.dtype([("min_cov", "<f8"), ("max_aaid", "<f8"), ("min_score", "<f8"), ("stat", "<i4"), ("pad", "<i4")])
        self.pars = np.frombuffer(blob, rec, n_len * n_fam, p).reshape(n_len, n_fam).copy(); p += 32 * n_len * n_fam
        self.coeff = np.frombuffer(blob, np.float64, n_len * n_fam, p).reshape(n_len, n_fam).copy(); p += 8 * n_len * n_fam
        self.weight = np.frombuffer(blob, np.float64, n_len * n_fam, p).reshape(n_len, n_fam).copy(); p += 8 * n_len * n_fam
        self.names = blob[p:p + names_bytes].decode().split("\n")
        self.n_subj, self.n_res, self.n_fam = n_subj, n_res, n_fam
        self.subj_len = np.diff(self.off)

    def length_index(self, read_length):
        try:
            return self.read_lengths.index(int(read_length))
        except ValueError:
            raise ValueError("read length %s is not one of %s" % (read_length, self.read_lengths))

    def cutoffs(self, read_length):
        """Rows of pars.map at this read length, family order = self.fam_names."""
        return self.pars[self.length_index(read_length)]

    def raw_cutoffs(self, read_length):
        return [min_raw_for_bits(float(r["min_score"])) for r in self.cutoffs(read_length)]

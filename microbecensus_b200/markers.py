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

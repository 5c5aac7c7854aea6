"""ctypes access to the CPU oracle (oracle/libmcoracle.so).  Test infrastructure only."""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB = os.path.join(ORACLE_DIR, "libmcoracle.so")


class OcDb(C.Structure):
    _fields_ = [("n_subj", C.c_int32), ("off", C.c_void_p), ("res", C.c_void_p), ("fam", C.c_void_p)]


class OcCutoff(C.Structure):
    _fields_ = [("min_cov", C.c_double), ("max_aaid", C.c_double), ("min_score", C.c_double),
                ("stat", C.c_int32), ("pad", C.c_int32)]


OC_HIT_FIELDS = ("read", "subject", "frame", "diag", "score", "aln", "ident", "mism", "gapo", "q0", "q1", "t0", "t1")
# column order of libmcx's mcx_hit
MCX_ORDER = [OC_HIT_FIELDS.index(k) for k in ("read", "subject", "frame", "score", "aln", "ident", "mism", "gapo", "q0", "q1", "t0", "t1")]


def build():
    if not os.path.isfile(LIB) or os.path.getmtime(LIB) < os.path.getmtime(os.path.join(ORACLE_DIR, "mc_oracle.c")):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "libmcoracle.so"], stdout=subprocess.DEVNULL)


class Oracle:
    def __init__(self, markers):
        build()
        self.lib = C.CDLL(LIB)
        self.m = markers
        self.db = OcDb(markers.n_subj, markers.off.ctypes.data, markers.res.ctypes.data, markers.fam.ctypes.data)
        L = self.lib
        L.oc_index_build.restype = C.c_void_p
        L.oc_index_build.argtypes = [C.POINTER(OcDb)]
        L.oc_index_free.argtypes = [C.c_void_p]
        L.oc_search_batch.restype = C.c_int64
        L.oc_search_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_int64, C.c_void_p]
        L.oc_classify.restype = C.c_int64
        L.oc_classify.argtypes = [C.POINTER(OcDb), C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_int64]
        L.oc_process_reads_d.restype = C.c_int64
        L.oc_process_reads_d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int,
                                         C.c_int, C.c_int, C.c_int64, C.c_void_p, C.c_void_p]
        L.oc_bits.restype = C.c_double
        L.oc_bits.argtypes = [C.c_int]
        L.oc_min_raw_for_bits.restype = C.c_int
        L.oc_min_raw_for_bits.argtypes = [C.c_double]
        L.oc_index_counts.restype = C.POINTER(C.c_int32)
        L.oc_index_counts.argtypes = [C.c_void_p]
        L.oc_frame.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.oc_alignment_coverage.restype = C.c_double
        L.oc_alignment_coverage.argtypes = [C.c_double] * 7
        L.oc_fingerprint.argtypes = [C.c_char_p, C.c_int, C.c_void_p]
        self.ix = L.oc_index_build(C.byref(self.db))

    def close(self):
        if self.ix:
            self.lib.oc_index_free(self.ix)
            self.ix = None

    def search(self, batch, L, min_raw, use_seg=True, cap=4_000_000):
        """-> int32 array (n_hits, 13) in OC_HIT_FIELDS order, grouped by read, (subject, score desc) within."""
        out = np.zeros((cap, 13), np.int32)
        nseeds = C.c_int64(0)
        n = self.lib.oc_search_batch(self.ix, batch.bases.ctypes.data, batch.offsets.ctypes.data, batch.n, int(L),
                                     1 if use_seg else 0, int(min_raw), out.ctypes.data, cap, C.byref(nseeds))
        assert n < cap, "oracle hit buffer too small"
        return out[:n].copy(), nseeds.value

    def classify(self, hits13, L, markers, n_reads):
        cut = (OcCutoff * 30)()
        rows = markers.cutoffs(L)
        for f in range(30):
            cut[f].min_cov, cut[f].max_aaid, cut[f].min_score = float(rows[f]["min_cov"]), float(rows[f]["max_aaid"]), float(rows[f]["min_score"])
            cut[f].stat = int(rows[f]["stat"])
        fam_hits = np.zeros(30, np.int64); fam_aln = np.zeros(30, np.int64); abl = np.zeros(30 * 1280, np.int64)
        best = np.full(max(n_reads, 1), -1, np.int32)
        hits13 = np.ascontiguousarray(hits13, np.int32)
        nc = self.lib.oc_classify(C.byref(self.db), hits13.ctypes.data, len(hits13), int(L), C.byref(cut),
                                  fam_hits.ctypes.data, fam_aln.ctypes.data, abl.ctypes.data, best.ctypes.data, n_reads)
        return {"classified": int(nc), "fam_hits": fam_hits, "fam_aln": fam_aln, "aln_by_len": abl.reshape(30, 1280),
                "best_subject": best[:n_reads]}

    def process_reads(self, batch, L, quality_offset, min_quality, mean_quality, max_unknown, nreads, filter_dups=False):
        code = np.zeros(max(batch.n, 1), np.uint8)
        counters = np.zeros(3, np.int64)
        q = batch.quals.ctypes.data if batch.quals is not None else None
        sampled = self.lib.oc_process_reads_d(batch.bases.ctypes.data, q, batch.offsets.ctypes.data, batch.n, int(L),
                                              int(quality_offset or 0), int(min_quality), int(mean_quality), int(max_unknown),
                                              1 if filter_dups else 0, -1 if nreads is None else int(nreads),
                                              code.ctypes.data, counters.ctypes.data)
        return int(sampled), code[:batch.n], {"too_short": int(counters[0]), "low_qual": int(counters[1]), "dups": int(counters[2])}

    def frame(self, seq, L, frame, use_seg=True):
        buf = (C.c_uint8 * 200)()
        m = self.lib.oc_frame(seq.encode(), int(L), int(frame), 1 if use_seg else 0, buf)
        return bytes(buf[:m])

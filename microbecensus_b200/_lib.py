"""ctypes binding of libmcx.so (include/mcx.h).

The shared library is built in-tree by ``microbecensus_b200/csrc/Makefile`` (``__graft_entry__.build()``)
and must be present: there is no Python or CPU fallback for the search.
"""
import ctypes as C
import os

N_FAM = 30
LEN_BINS = 1280

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libmcx.so")


class McxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libmcx error %d: %s" % (code, msg))
        self.code = code
        self.msg = msg


class Db(C.Structure):
    _fields_ = [("n_subj", C.c_int32), ("off", C.c_void_p), ("res", C.c_void_p), ("fam", C.c_void_p)]


class Cutoff(C.Structure):
    _fields_ = [("min_cov", C.c_double), ("max_aaid", C.c_double), ("min_raw", C.c_int32), ("stat", C.c_int32)]


class Params(C.Structure):
    _fields_ = [("read_length", C.c_int32), ("has_quality", C.c_int32), ("quality_offset", C.c_int32),
                ("min_quality", C.c_int32), ("mean_quality", C.c_int32), ("max_unknown", C.c_int32),
                ("filter_dups", C.c_int32), ("min_report_raw", C.c_int32), ("cut", Cutoff * N_FAM)]


class Qc(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("kept", C.c_int64), ("too_short", C.c_int64),
                ("low_qual", C.c_int64), ("dups", C.c_int64)]


class Result(C.Structure):
    _fields_ = [("sampled_reads", C.c_int64), ("too_short", C.c_int64), ("low_qual", C.c_int64),
                ("dups", C.c_int64), ("reads_with_hits", C.c_int64), ("reads_classified", C.c_int64),
                ("n_hsp", C.c_int64), ("n_seed_hits", C.c_int64), ("n_gapped", C.c_int64),
                ("gapped_cells", C.c_int64), ("n_capped_reads", C.c_int64), ("fam_hits", C.c_int64 * N_FAM), ("fam_aln", C.c_int64 * N_FAM),
                ("aln_by_len", C.c_int64 * (N_FAM * LEN_BINS))]


class Hit(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("read", "subject", "frame", "score", "aln", "ident", "mism", "gapo",
                                         "q0", "q1", "t0", "t1")]


EXPORTS = ("mcx_create", "mcx_destroy", "mcx_set_params", "mcx_set_stream", "mcx_push_reads", "mcx_push_reads_dev",
           "mcx_push_reads_packed", "mcx_push_reads_packed_dev", "mcx_host_alloc", "mcx_host_free",
           "mcx_qc_counts", "mcx_qc_export", "mcx_qc_import", "mcx_qc_device", "mcx_qc_refresh", "mcx_dedup_reset", "mcx_dedup_begin", "mcx_dedup_owner", "mcx_dedup_finish", "mcx_search", "mcx_result_get", "mcx_get_hits", "mcx_get_classified",
           "mcx_timings", "mcx_timings_detail", "mcx_search_counters", "mcx_dpx_peak", "mcx_l2_peak", "mcx_last_error", "mcx_version")

_lib = None


def load():
    """Load libmcx.so (once).  Raises OSError when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise OSError("libmcx.so not found at %s: build it with `make -C microbecensus_b200/csrc` "
                      "(__graft_entry__.build()); the search has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    lib.mcx_create.argtypes = [C.POINTER(vp), C.POINTER(Db), C.c_int]
    lib.mcx_destroy.argtypes = [vp]
    lib.mcx_destroy.restype = None
    lib.mcx_set_params.argtypes = [vp, C.POINTER(Params)]
    lib.mcx_set_stream.argtypes = [vp, vp]
    lib.mcx_push_reads.argtypes = [vp, vp, vp, vp, i64]
    lib.mcx_push_reads_dev.argtypes = [vp, vp, vp, vp, i64, i64]
    lib.mcx_push_reads_packed.argtypes = [vp, vp, i64, vp, vp, i64, i64]
    lib.mcx_push_reads_packed_dev.argtypes = [vp, vp, i64, vp, vp, i64, i64]
    lib.mcx_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.mcx_host_free.argtypes = [vp]
    lib.mcx_host_free.restype = None
    lib.mcx_qc_counts.argtypes = [vp, C.POINTER(Qc)]
    lib.mcx_qc_export.argtypes = [vp, vp, vp]
    lib.mcx_qc_import.argtypes = [vp, vp]
    lib.mcx_qc_device.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
    lib.mcx_qc_refresh.argtypes = [vp]
    lib.mcx_dedup_reset.argtypes = [vp]
    lib.mcx_dedup_begin.argtypes = [vp, C.c_int, i64, C.POINTER(vp), C.POINTER(i64)]
    lib.mcx_dedup_owner.argtypes = [vp, vp, i64, C.POINTER(vp)]
    lib.mcx_dedup_finish.argtypes = [vp, vp]
    lib.mcx_search.argtypes = [vp, i64]
    lib.mcx_result_get.argtypes = [vp, C.POINTER(Result)]
    lib.mcx_get_hits.argtypes = [vp, vp, i64, C.POINTER(i64)]
    lib.mcx_get_classified.argtypes = [vp, vp, i64]
    lib.mcx_timings.argtypes = [vp, C.POINTER(C.c_float * 12), C.POINTER(i64)]
    lib.mcx_timings_detail.argtypes = [vp, C.POINTER(C.c_float * 4)]
    lib.mcx_search_counters.argtypes = [vp, C.POINTER(C.c_int64 * 8)]
    lib.mcx_dpx_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.mcx_l2_peak.argtypes = [vp, C.POINTER(C.c_double)]
    lib.mcx_last_error.argtypes = [vp]
    lib.mcx_last_error.restype = C.c_char_p
    lib.mcx_version.restype = C.c_char_p
    for name in EXPORTS:
        if name not in ("mcx_destroy", "mcx_last_error", "mcx_version", "mcx_host_free"):
            getattr(lib, name).restype = C.c_int
    _lib = lib
    return lib


def check(lib, ctx, rc):
    if rc != 0:
        msg = lib.mcx_last_error(ctx)
        raise McxError(rc, msg.decode() if msg else "")

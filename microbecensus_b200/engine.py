"""Host-side driver of libmcx: one MarkerSearch = one GPU context with the marker index resident.

Mirrors the seam of the reference pipeline (microbe_census.py:611-620): reads in, sampled_reads +
agg_hits out.  All computation happens in the CUDA library; this module only moves arrays.
"""
import ctypes as C

import numpy as np

from . import _lib
from .markers import Markers, min_raw_for_bits, report_floor, STAT_NAMES


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


class ReadBatch:
    """Reads as the library wants them: concatenated ASCII bases (+ qualities) and n+1 offsets."""

    def __init__(self, bases, offsets, quals=None):
        self.bases = np.ascontiguousarray(bases, dtype=np.uint8)
        self.offsets = np.ascontiguousarray(offsets, dtype=np.int64)
        self.quals = None if quals is None else np.ascontiguousarray(quals, dtype=np.uint8)
        self.n = len(self.offsets) - 1

    @classmethod
    def from_strings(cls, seqs, quals=None):
        offs = np.zeros(len(seqs) + 1, np.int64)
        if len(seqs):
            offs[1:] = np.cumsum([len(s) for s in seqs])
        b = np.frombuffer("".join(seqs).encode("ascii", "replace"), np.uint8) if len(seqs) else np.zeros(0, np.uint8)
        q = None
        if quals is not None:
            q = np.frombuffer("".join(quals).encode("ascii", "replace"), np.uint8) if len(quals) else np.zeros(0, np.uint8)
        return cls(b, offs, q)

    def slice(self, lo, hi):
        o = self.offsets[lo:hi + 1]
        b0, b1 = int(o[0]), int(o[-1])
        return ReadBatch(self.bases[b0:b1], o - b0, None if self.quals is None else self.quals[b0:b1])

    @property
    def nbytes(self):
        return self.bases.nbytes + self.offsets.nbytes + (0 if self.quals is None else self.quals.nbytes)


class SearchResult:
    def __init__(self, raw, markers, read_length):
        self.sampled_reads = int(raw.sampled_reads)
        self.too_short = int(raw.too_short)
        self.low_qual = int(raw.low_qual)
        self.dups = int(raw.dups)
        self.reads_with_hits = int(raw.reads_with_hits)
        self.reads_classified = int(raw.reads_classified)
        self.n_hsp = int(raw.n_hsp)
        self.n_seed_hits = int(raw.n_seed_hits)
        self.n_gapped = int(raw.n_gapped)
        self.gapped_cells = int(raw.gapped_cells)
        self.fam_hits = np.array(raw.fam_hits, np.int64)
        self.fam_aln = np.array(raw.fam_aln, np.int64)
        self.aln_by_len = np.array(raw.aln_by_len, np.int64).reshape(_lib.N_FAM, _lib.LEN_BINS)
        self._markers = markers
        self.read_length = read_length

    def counts_vector(self):
        """Everything additive, as one int64 vector (what a multi-GPU run all-reduces)."""
        head = np.array([self.sampled_reads, self.too_short, self.low_qual, self.dups, self.reads_with_hits,
                         self.reads_classified, self.n_hsp, self.n_seed_hits, self.n_gapped, self.gapped_cells], np.int64)
        return np.concatenate([head, self.fam_hits, self.fam_aln, self.aln_by_len.ravel()])

    def load_counts_vector(self, v):
        v = np.asarray(v, np.int64)
        (self.sampled_reads, self.too_short, self.low_qual, self.dups, self.reads_with_hits, self.reads_classified,
         self.n_hsp, self.n_seed_hits, self.n_gapped, self.gapped_cells) = (int(x) for x in v[:10])
        nf = _lib.N_FAM
        self.fam_hits = v[10:10 + nf].copy()
        self.fam_aln = v[10 + nf:10 + 2 * nf].copy()
        self.aln_by_len = v[10 + 2 * nf:].reshape(nf, _lib.LEN_BINS).copy()

    def agg_hits(self):
        """{family id: weighted count} as aggregate_hits returns it (microbe_census.py:462-472).

        `hits` families count reads, `aln` families sum alignment lengths, `cov` families sum
        aln/target_len, formed here as sum over subject lengths of (integer sum of aln)/length in
        ascending length order -- independent of read order and of the number of GPUs."""
        cut = self._markers.cutoffs(self.read_length)
        out = {}
        for f, name in enumerate(self._markers.fam_names):
            if self.fam_hits[f] == 0:
                continue
            stat = int(cut[f]["stat"])
            if stat == 0:
                out[name] = float(self.fam_hits[f])
            elif stat == 2:
                out[name] = float(self.fam_aln[f])
            else:
                total = 0.0
                row = self.aln_by_len[f]
                for ln in np.nonzero(row)[0]:
                    total += float(row[ln]) / float(ln)
                out[name] = total
        return out


class MarkerSearch:
    """GPU context: marker residues + seed index on the device, ready to search batches of reads."""

    def __init__(self, markers=None, device=0):
        self.markers = markers if isinstance(markers, Markers) else Markers(markers)
        self.lib = _lib.load()
        m = self.markers
        self._db = _lib.Db(m.n_subj, m.off.ctypes.data, m.res.ctypes.data, m.fam.ctypes.data)
        self.ctx = C.c_void_p(0)
        rc = self.lib.mcx_create(C.byref(self.ctx), C.byref(self._db), int(device))
        if rc != 0:
            _lib.check(self.lib, None, rc)
        self.read_length = None
        self._batch = None

    def close(self):
        if self.ctx:
            self.lib.mcx_destroy(self.ctx)
            self.ctx = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        _lib.check(self.lib, self.ctx, rc)

    def set_params(self, read_length, quality_offset=None, min_quality=-5, mean_quality=-5, max_unknown=100,
                   filter_dups=False, min_report_raw=None):
        m = self.markers
        cut = m.cutoffs(read_length)
        raws = m.raw_cutoffs(read_length)
        p = _lib.Params()
        p.read_length = int(read_length)
        p.has_quality = 0 if quality_offset is None else 1
        p.quality_offset = int(quality_offset or 0)
        p.min_quality, p.mean_quality, p.max_unknown = int(min_quality), int(mean_quality), int(max_unknown)
        p.filter_dups = 1 if filter_dups else 0
        floor = report_floor(read_length) if min_report_raw is None else int(min_report_raw)
        p.min_report_raw = min(floor, min(raws))
        for f in range(_lib.N_FAM):
            p.cut[f].min_cov = float(cut[f]["min_cov"])
            p.cut[f].max_aaid = float(cut[f]["max_aaid"])
            p.cut[f].min_raw = int(raws[f])
            p.cut[f].stat = int(cut[f]["stat"])
        self._ck(self.lib.mcx_set_params(self.ctx, C.byref(p)))
        self.read_length = int(read_length)
        self.min_report_raw = int(p.min_report_raw)

    def set_stream(self, cuda_stream):
        """Run on the caller's stream (int handle of a cudaStream_t)."""
        self._ck(self.lib.mcx_set_stream(self.ctx, C.c_void_p(int(cuda_stream))))

    def push(self, batch):
        """Host -> device copy of the reads + QC kernel.  Returns the QC counters over all pushed reads."""
        self._batch = batch  # keep the arrays alive
        self._ck(self.lib.mcx_push_reads(self.ctx, _ptr(batch.bases), _ptr(batch.quals), _ptr(batch.offsets), batch.n))
        return self.qc()

    def push_device(self, d_bases_ptr, d_quals_ptr, d_offsets_ptr, n, total_bytes):
        self._ck(self.lib.mcx_push_reads_dev(self.ctx, C.c_void_p(d_bases_ptr), C.c_void_p(d_quals_ptr or 0),
                                             C.c_void_p(d_offsets_ptr), int(n), int(total_bytes)))
        return self.qc()

    def qc(self):
        q = _lib.Qc()
        self._ck(self.lib.mcx_qc_counts(self.ctx, C.byref(q)))
        return {"n_reads": q.n_reads, "kept": q.kept, "too_short": q.too_short, "low_qual": q.low_qual, "dups": q.dups}

    def qc_export(self, with_fingerprints=False):
        """(codes uint8[n], fingerprints uint64[n, 2] or None) of the pushed reads"""
        n = self.qc()["n_reads"]
        code = np.zeros(max(n, 1), np.uint8)
        fp = np.zeros((max(n, 1), 2), np.uint64) if with_fingerprints else None
        self._ck(self.lib.mcx_qc_export(self.ctx, _ptr(code), _ptr(fp)))
        return code[:n], (None if fp is None else fp[:n])

    def qc_device(self, with_fingerprints=True):
        """(pointer to the n verdict bytes, pointer to the n 24-byte fingerprint records or 0, n) in device memory"""
        dc, df, n = C.c_void_p(0), C.c_void_p(0), C.c_int64(0)
        self._ck(self.lib.mcx_qc_device(self.ctx, C.byref(dc), C.byref(df) if with_fingerprints else None, C.byref(n)))
        return dc.value or 0, df.value or 0, n.value

    def qc_refresh(self):
        """rebuild the kept list after the verdicts were rewritten in device memory"""
        self._ck(self.lib.mcx_qc_refresh(self.ctx))
        return self.qc()

    def qc_import(self, code):
        code = np.ascontiguousarray(code, np.uint8)
        self._ck(self.lib.mcx_qc_import(self.ctx, _ptr(code)))
        return self.qc()

    def search(self, quota=-1):
        self._ck(self.lib.mcx_search(self.ctx, -1 if quota is None else int(quota)))
        raw = _lib.Result()
        self._ck(self.lib.mcx_result_get(self.ctx, C.byref(raw)))
        return SearchResult(raw, self.markers, self.read_length)

    def hits(self):
        n = C.c_int64(0)
        self._ck(self.lib.mcx_get_hits(self.ctx, C.c_void_p(0), 0, C.byref(n)))
        arr = np.zeros((max(n.value, 1), 12), np.int32)
        self._ck(self.lib.mcx_get_hits(self.ctx, _ptr(arr), n.value, C.byref(n)))
        return arr[:n.value]

    def classified(self, n):
        out = np.full(int(n), -1, np.int32)
        self._ck(self.lib.mcx_get_classified(self.ctx, _ptr(out), int(n)))
        return out

    def dpx_peak(self):
        """1e9 DPX thread-instructions per second this GPU issues (viaddmax_s32 / vimax3_s32_relu microbenchmark)."""
        v = C.c_double(0.0)
        self._ck(self.lib.mcx_dpx_peak(self.ctx, C.byref(v)))
        return v.value

    def timings(self):
        ms = (C.c_float * 10)()
        launches = C.c_int64(0)
        self._ck(self.lib.mcx_timings(self.ctx, C.byref(ms), C.byref(launches)))
        names = ("h2d", "qc", "probe", "gapped", "sort", "classify", "d2h", "extend", "frames", "seg")
        return {k: float(ms[i]) for i, k in enumerate(names)}, int(launches.value)


HIT_FIELDS = ("read", "subject", "frame", "score", "aln", "ident", "mism", "gapo", "q0", "q1", "t0", "t1")


def dna_coords(L, frame, q0, q1):
    """aa range (0-based inclusive) on a frame -> 1-based q.start/q.end on the read, as RAPsearch2 prints them."""
    a0, a1 = q0 + 1, q1 + 1
    if frame < 3:
        return 3 * (a0 - 1) + frame + 1, 3 * a1 + frame
    o = frame - 3
    return L - o - 3 * (a0 - 1), L - o - 3 * a1 + 1


def format_m8(hits, markers, read_length, read_names=None):
    """m8 lines in RAPsearch2's layout (the file search_seqs leaves for parse_rapsearch, mc.py:391-398): -b 0
    subject coordinates, identity with six significant digits, log10 E and bit score with two decimals."""
    from .markers import bits_printed, log10_evalue_printed
    lines = []
    for h in hits:
        d = dict(zip(HIT_FIELDS, (int(x) for x in h)))
        qs, qe = dna_coords(read_length, d["frame"], d["q0"], d["q1"])
        name = str(d["read"]) if read_names is None else read_names[d["read"]]
        lines.append("%s\t%s\t%g\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%s\t%s" % (
            name, markers.names[d["subject"]], 100.0 * d["ident"] / d["aln"], d["aln"], d["mism"], d["gapo"],
            qs, qe, d["t0"], d["t1"], "%g" % log10_evalue_printed(d["score"], read_length), "%g" % bits_printed(d["score"])))
    return lines

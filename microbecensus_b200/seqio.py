"""ctypes binding of libmcxio.so (include/mcxio.h): the streaming FASTA/FASTQ reader that stands in for
open_file() + parse_seqs() (mc.py:47-59, 294-325) and folds count_bases() (mc.py:573-584) into the same pass."""
import bz2
import ctypes as C
import os

import numpy as np

from .engine import ReadBatch, PackedBatch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmcxio.so")
EXPORTS = ("mcxio_open", "mcxio_open_mem", "mcxio_next_batch", "mcxio_next_packed", "mcxio_skip_packed", "mcxio_state", "mcxio_free_packed", "mcxio_set_allocator",
           "mcxio_skip_rest", "mcxio_close", "mcxio_last_error")


class Batch(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("quals", C.c_void_p), ("offsets", C.c_void_p), ("n", C.c_int64),
                ("records_total", C.c_int64), ("bases_total", C.c_int64), ("eof", C.c_int32)]


class Packed(C.Structure):
    _fields_ = [("packed", C.c_void_p), ("lengths", C.c_void_p), ("quals", C.c_void_p), ("n", C.c_int64), ("n_words", C.c_int64),
                ("n_bases", C.c_int64), ("records_total", C.c_int64), ("bases_total", C.c_int64), ("eof", C.c_int32),
                ("last_without_quality", C.c_int32), ("reparsed", C.c_int64)]


class _PackedOwner:
    """keeps the reader's output buffers alive for the numpy views and releases them through the reader's allocator"""
    def __init__(self, lib, rec):
        self.lib, self.rec = lib, rec

    def __del__(self):
        try:
            if self.rec is not None:
                self.lib.mcxio_free_packed(C.byref(self.rec))
                self.rec = None
        except Exception:
            pass


def use_pinned_buffers(libmcx):
    """Packed batches from now on come in page-locked memory from libmcx (mcx_host_alloc): the push to the GPU is then an
    asynchronous DMA that overlaps the search (called by MarkerSearch once a GPU context exists)."""
    lib = load()
    lib.mcxio_set_allocator(C.cast(libmcx.mcx_host_alloc, C.c_void_p), C.cast(libmcx.mcx_host_free, C.c_void_p))


class SeqIOError(IOError):
    pass


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise OSError("libmcxio.so not found at %s: build it with `make -C microbecensus_b200/csrc`" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i64 = C.c_void_p, C.c_int64
    lib.mcxio_open.argtypes = [C.POINTER(vp), C.c_char_p]
    lib.mcxio_open_mem.argtypes = [C.POINTER(vp), vp, i64]
    lib.mcxio_next_batch.argtypes = [vp, i64, C.POINTER(Batch)]
    lib.mcxio_next_packed.argtypes = [vp, i64, C.c_int, C.POINTER(Packed)]
    lib.mcxio_skip_packed.argtypes = [vp, i64, C.c_int, C.POINTER(i64)]
    lib.mcxio_state.argtypes = [vp, C.POINTER(i64), C.POINTER(i64), C.POINTER(C.c_int32)]
    lib.mcxio_free_packed.argtypes = [C.POINTER(Packed)]
    lib.mcxio_free_packed.restype = None
    lib.mcxio_set_allocator.argtypes = [vp, vp]
    lib.mcxio_skip_rest.argtypes = [vp, C.POINTER(i64), C.POINTER(i64)]
    lib.mcxio_close.argtypes = [vp]
    lib.mcxio_close.restype = None
    lib.mcxio_last_error.argtypes = [vp]
    lib.mcxio_last_error.restype = C.c_char_p
    _lib = lib
    return lib


def _view(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=n)


class SeqFile:
    """One FASTA/FASTQ file (plain, .gz or .bz2) read record by record in the reference's readfq semantics."""

    def __init__(self, path):
        self._lib = load()
        self._h = C.c_void_p()
        self._keep = None
        with open(path, "rb") as fh:
            magic = fh.read(3)
        if magic == b"BZh":                      # open_file(): bz2 by extension; here by magic, inflated on the host
            with bz2.open(path, "rb") as fh:
                data = fh.read()
            self._keep = np.frombuffer(data, np.uint8)
            rc = self._lib.mcxio_open_mem(C.byref(self._h), self._keep.ctypes.data, len(data))
        else:
            rc = self._lib.mcxio_open(C.byref(self._h), os.fsencode(path))
        if rc != 0:
            raise SeqIOError(self._lib.mcxio_last_error(None).decode())
        self.records_total = 0
        self.bases_total = 0
        self.eof = False

    @classmethod
    def from_bytes(cls, data):
        self = cls.__new__(cls)
        self._lib = load()
        self._h = C.c_void_p()
        self._keep = np.frombuffer(bytes(data), np.uint8)
        rc = self._lib.mcxio_open_mem(C.byref(self._h), self._keep.ctypes.data if len(self._keep) else None, len(self._keep))
        if rc != 0:
            raise SeqIOError(self._lib.mcxio_last_error(None).decode())
        self.records_total = 0
        self.bases_total = 0
        self.eof = False
        return self

    def next_batch(self, max_records=None, copy=True):
        """Up to max_records further records (None: the rest of the file) as a ReadBatch.  copy=False returns views
        of the reader's own buffers, valid until the next call on this reader (the streaming path pushes each batch
        to the GPU before asking for the next one)."""
        b = Batch()
        rc = self._lib.mcxio_next_batch(self._h, -1 if max_records is None else int(max_records), C.byref(b))
        if rc != 0:
            raise SeqIOError(self._lib.mcxio_last_error(self._h).decode())
        offs = _view(b.offsets, b.n + 1, np.int64)
        total = int(offs[-1]) if b.n else 0
        bases = _view(b.bases, total, np.uint8)
        quals = _view(b.quals, total, np.uint8) if b.quals else None
        if copy:
            offs, bases, quals = offs.copy(), bases.copy(), None if quals is None else quals.copy()
        self.records_total, self.bases_total, self.eof = b.records_total, b.bases_total, bool(b.eof)
        return ReadBatch(bases, offs, quals)

    def next_packed(self, target_records=None, threads=1):
        """About target_records further records (None: the rest of the file) as a PackedBatch -- the layout of the reads
        in HBM -- parsed and packed by `threads` threads (include/mcxio.h).  The batch owns its buffers."""
        rec = Packed()
        rc = self._lib.mcxio_next_packed(self._h, -1 if target_records is None else int(target_records), int(threads), C.byref(rec))
        if rc != 0:
            raise SeqIOError(self._lib.mcxio_last_error(self._h).decode())
        owner = _PackedOwner(self._lib, rec)
        pb = PackedBatch.__new__(PackedBatch)
        pb.packed = _view(rec.packed, rec.n_words, np.uint32)
        pb.lengths = _view(rec.lengths, rec.n, np.uint32)
        pb.quals = _view(rec.quals, rec.n_bases, np.uint8) if rec.quals else None
        pb.n, pb.n_bases = int(rec.n), int(rec.n_bases)
        pb._owners = [owner]
        self.records_total, self.bases_total, self.eof = rec.records_total, rec.bases_total, bool(rec.eof)
        self.last_without_quality = bool(rec.last_without_quality)
        self.reparsed = int(rec.reparsed)
        return pb

    def skip_packed(self, target_records=None, threads=1):
        """Advance by the records the same next_packed call would return, without storing them; returns how many."""
        n = C.c_int64(0)
        rc = self._lib.mcxio_skip_packed(self._h, -1 if target_records is None else int(target_records), int(threads), C.byref(n))
        if rc != 0:
            raise SeqIOError(self._lib.mcxio_last_error(self._h).decode())
        rt, bt, eof = C.c_int64(0), C.c_int64(0), C.c_int32(0)
        self._lib.mcxio_state(self._h, C.byref(rt), C.byref(bt), C.byref(eof))
        self.records_total, self.bases_total, self.eof = rt.value, bt.value, bool(eof.value)
        return int(n.value)

    def skip_rest(self):
        """Parse to the end of the file without storing records; returns (records, bases) of the whole file."""
        n, t = C.c_int64(), C.c_int64()
        rc = self._lib.mcxio_skip_rest(self._h, C.byref(n), C.byref(t))
        if rc != 0:
            raise SeqIOError(self._lib.mcxio_last_error(self._h).decode())
        self.records_total, self.bases_total, self.eof = n.value, t.value, True
        return n.value, t.value

    def close(self):
        if self._h:
            self._lib.mcxio_close(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

"""ctypes binding of libmcxio.so (include/mcxio.h): the streaming FASTA/FASTQ reader that stands in for
open_file() + parse_seqs() (mc.py:47-59, 294-325) and folds count_bases() (mc.py:573-584) into the same pass."""
import bz2
import ctypes as C
import os

import numpy as np

from .engine import ReadBatch

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libmcxio.so")
EXPORTS = ("mcxio_open", "mcxio_open_mem", "mcxio_next_batch", "mcxio_skip_rest", "mcxio_close", "mcxio_last_error")


class Batch(C.Structure):
    _fields_ = [("bases", C.c_void_p), ("quals", C.c_void_p), ("offsets", C.c_void_p), ("n", C.c_int64),
                ("records_total", C.c_int64), ("bases_total", C.c_int64), ("eof", C.c_int32)]


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

"""Drop-in for ``microbe_census.microbe_census`` (MicrobeCensus 1.1.0) with the read sampling, marker
search and classification stages running on a B200 through libmcx.

Public surface kept from the reference (file:line in /root/reference/microbe_census/microbe_census.py):
``run_pipeline(args) -> (est_ags, args)`` (:586), ``count_bases(args)`` (:573), ``report_results`` (:514),
``estimate_average_genome_size`` (:474), ``auto_detect_file_type`` / ``_quality_offset`` / ``_read_length``
(:166, :174, :147), ``read_list`` (:90), ``open_file`` (:47), ``parse_seqs`` (:294), the ``args`` keys the
pipeline fills in, the verbose messages and the ``sys.exit`` texts.  What changed: the four calls at
:611-:620 (process_seqfile, search_seqs -> RAPsearch2 child process, classify_reads, aggregate_hits) are one
GPU search; ``-t`` is accepted and ignored by the search; ``-r`` (alternative rapsearch binary) has nothing
to point at and is ignored.  ``-d`` compares 128-bit strand-canonical fingerprints of the untrimmed reads
instead of whole strings (the reference's KeyError on non-ACGTN characters under ``-d`` is not reproduced).
"""
import bz2
import gzip
import io
import os
import platform
import sys

import numpy as np
from numpy import median

from .engine import MarkerSearch, ReadBatch
from .seqio import SeqFile
from .markers import Markers

__version__ = "1.1.0"

VALID_LENGTHS = [50, 60, 70, 80, 90, 100, 110, 120, 130, 140, 150, 175, 200, 225, 250, 300, 350, 400, 450, 500]

_markers = None
_engines = {}


def get_markers():
    global _markers
    if _markers is None:
        _markers = Markers()
    return _markers


def default_device():
    """The GPU this process searches on: $MCX_DEVICE if set; under torchrun (one process per GPU) the rank's own GPU --
    torch's current device once a process group is up, else $LOCAL_RANK; otherwise GPU 0."""
    if os.environ.get("MCX_DEVICE"):
        return int(os.environ["MCX_DEVICE"])
    if "torch" in sys.modules:
        try:
            import torch
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl":
                return int(torch.cuda.current_device()) if os.environ.get("LOCAL_RANK") is None else int(os.environ["LOCAL_RANK"])
        except Exception:
            pass
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("LOCAL_RANK") is not None:
        return int(os.environ["LOCAL_RANK"])
    return 0


def get_engine(device=None):
    """One GPU context per device, created on first use (marker index upload ~ seconds) and reused."""
    device = default_device() if device is None else int(device)
    if device not in _engines:
        _engines[device] = MarkerSearch(get_markers(), device)
    return _engines[device]


# ------------------------------------------------------------------------------------------------ utilities
def mad(x, const=1.48):
    """Median absolute deviation, scaled (mc.py:43-45)."""
    m = median(x)
    return const * median([abs(i - m) for i in x])


def open_file(inpath):
    """Text handle on a plain / .gz / .bz2 file (mc.py:47-59)."""
    ext = inpath.split(".")[-1]
    if ext == "gz":
        return io.TextIOWrapper(gzip.open(inpath))
    if ext == "bz2":
        return io.TextIOWrapper(bz2.BZ2File(inpath))
    return open(inpath)


def read_list(file, header, dtype):
    """One value per line (mc.py:90-99)."""
    conv = float if dtype == "float" else int if dtype == "int" else (lambda v: v)
    with open_file(file) as f_in:
        if header is True:
            next(f_in)
        return [conv(line.rstrip()) for line in f_in]


def check_os():
    if platform.system() not in ["Linux", "Darwin"]:
        sys.exit("Operating system '%s' not supported" % platform.system())


class Sequence:
    def __init__(self, name, seq, quality=None):
        self.id = name
        self.seq = seq
        self.quality = quality

    def phred(self, offset):
        return [ord(c) - offset for c in self.quality]


def parse_seqs(fp):
    """FASTA/FASTQ records with readfq semantics (mc.py:294-325): a record starts at a line beginning with
    '>' or '@'; the name ends at the first space; sequence lines run until a line beginning with '@', '+'
    or '>'; after a '+' line quality lines are consumed until they are at least as long as the sequence;
    a FASTQ record cut short by EOF is yielded without qualities."""
    pending = None            # readfq's `last`: tested for truth, so an empty stripped line counts as "none"
    lines = iter(fp)
    while True:
        if not pending:
            for line in lines:
                if line[0] in ">@":
                    pending = line[:-1]
                    break
            if not pending:
                return
        name = pending[1:].partition(" ")[0]
        pending = None
        chunks = []
        for line in lines:
            if line[0] in "@+>":
                pending = line[:-1]
                break
            chunks.append(line[:-1])
        seq = "".join(chunks)
        if not pending or pending[0] != "+":
            yield Sequence(name, seq)
            if not pending:
                return
            continue
        got, qchunks, complete = 0, [], False
        for line in lines:
            qchunks.append(line[:-1])
            got += len(line) - 1
            if got >= len(seq):
                complete = True
                break
        pending = None
        if complete:
            yield Sequence(name, seq, "".join(qchunks))
        else:
            yield Sequence(name, seq)
            return


def auto_detect_file_type(seqfile):
    with open_file(seqfile) as f_in:
        for line in f_in:
            if line[0] == ">":
                return "fasta"
            if line[0] == "@":
                return "fastq"
            sys.exit("Filetype [fasta, fastq] of %s could not be recognized" % seqfile)


def auto_detect_quality_offset(seqfile):
    """32 if a character only Phred+33 files use shows up first, 64 for the Phred+64-only range (mc.py:174-187)."""
    low = set("""!"#$%&'()*+,-./0123456789""")
    high = set("""KLMNOPQRSTUVWXYZ[\\]^_`abcdefghijklmnopqrstuvwxyz{|}~""")
    with open_file(seqfile) as f_in:
        for rec in parse_seqs(f_in):
            for ch in rec.quality:
                if ch in low:
                    return 32
                if ch in high:
                    return 64
    return 32


def auto_detect_read_length(seqfile, file_type):
    """Largest supported length not above the median of the first 10,000 reads (mc.py:147-164)."""
    lengths = []
    try:
        with open_file(seqfile) as f_in:
            for index, rec in enumerate(parse_seqs(f_in)):
                if index == 10000:
                    break
                lengths.append(len(rec.seq))
    except Exception:
        sys.exit("Could not detect read length of: %s\nThis may be due to an invalid format\nTry specifying it with -l" % seqfile)
    med = int(median(lengths))
    if med < VALID_LENGTHS[0]:
        sys.exit("Median read length is %s. Cannot compute AGS using reads shorter than 50 bp." % med)
    for index, length in enumerate(VALID_LENGTHS):
        if length > med:
            return VALID_LENGTHS[index - 1]
    return VALID_LENGTHS[-1]


def impute_missing_args(args):
    for key, value in (("verbose", False), ("outfile", None), ("nreads", 1000000), ("threads", 1),
                       ("filter_dups", False), ("keep_tmp", False), ("mean_quality", -5), ("min_quality", -5),
                       ("max_unknown", 100)):
        if key not in args:
            args[key] = value
    args["file_type"] = auto_detect_file_type(args["seqfiles"][0])
    if args["file_type"] == "fastq":
        args["quality_offset"] = auto_detect_quality_offset(args["seqfiles"][0])
    if "read_length" not in args or args["read_length"] is None:
        args["read_length"] = auto_detect_read_length(args["seqfiles"][0], args["file_type"])


def check_input(args):
    for seqfile in args["seqfiles"]:
        if not os.path.isfile(seqfile):
            sys.exit("Input file %s not found" % seqfile)


def check_arguments(args):
    if args["file_type"] == "fasta" and any([args["min_quality"] > -5, args["mean_quality"] > -5]):
        sys.exit("Quality filtering options are only available for FASTQ files")
    if args["threads"] < 1:
        sys.exit("Invalid number of threads: %s\nMust be a positive integer." % args["threads"])
    if args["nreads"] is not None and args["nreads"] < 1:
        sys.exit("Invalid number of reads: %s\nMust be a positive integer." % args["nreads"])


def print_copyright():
    print("\nMicrobeCensus - estimation of average genome size from shotgun sequence data")
    print("version %s; github.com/snayfach/MicrobeCensus" % __version__)
    print("Copyright (C) 2013-2015 Stephen Nayfach")
    print("Freely distributed under the GNU General Public License (GPLv3)\n")


def print_parameters(args):
    fq = args["file_type"] == "fastq"
    print("=============Parameters==============")
    print("Input metagenome: %s" % args["seqfiles"])
    print("Output file: %s" % args["outfile"])
    print("Reads trimmed to: %s bp" % args["read_length"])
    print("Maximum reads sampled: %s" % args["nreads"])
    print("Threads to use for db search: %s" % args["threads"])
    print("Minimum base-level quality score: %s" % (args["min_quality"] if fq else "NA"))
    print("Minimum read-level quality score: %s" % (args["mean_quality"] if fq else "NA"))
    print("Maximum percent unknown bases/read: %s" % args["max_unknown"])
    print("Filter duplicate reads: %s" % args["filter_dups"])
    print("Keep temporary files: %s\n" % args["keep_tmp"])


# ------------------------------------------------------------------------------------------------ read loading
# Files are read by libmcxio (microbecensus_b200/csrc/mcxio.cpp, include/mcxio.h): the readfq state machine of
# parse_seqs() in C++, inflating / reading ahead in a producer thread and returning packed batches.  parse_seqs()
# above stays as the mirror of the reference's generator (auto-detection uses it; tests check the reader against it).
_base_counts = {}          # (path, size, mtime) -> total bases of the file, filled when a file was read to its end


def _file_key(path):
    st = os.stat(path)
    return (os.path.abspath(path), st.st_size, st.st_mtime_ns)


def load_reads(seqfile, file_type=None, max_records=None):
    """All (or the first max_records) records of one file as a ReadBatch."""
    with SeqFile(seqfile) as rd:
        batch = rd.next_batch(max_records)
        if rd.eof:
            _base_counts[_file_key(seqfile)] = rd.bases_total
    return batch


def concat_batches(batches):
    if len(batches) == 1:
        return batches[0]
    bases = np.concatenate([b.bases for b in batches])
    has_q = all(b.quals is not None for b in batches)
    quals = np.concatenate([b.quals for b in batches]) if has_q else None
    offs = [np.zeros(1, np.int64)]
    base = 0
    for b in batches:
        offs.append(b.offsets[1:] + base)
        base += int(b.offsets[-1])
    return ReadBatch(bases, np.concatenate(offs), quals)


def _dump_m8(eng, fh, read_length, first_id):
    """Append the reported HSPs of the last search to `fh` in RAPsearch2's m8 layout (the <tmp>.m8 of mc.py:375 that
    parse_rapsearch reads): queries are named by their running index among the sampled reads, as process_seqfile
    names them (mc.py:352), lines grouped by query in input order, best score first."""
    from .engine import format_m8
    hits = eng.hits()
    if len(hits) == 0:
        return
    codes, _ = eng.qc_export(False)
    rank = np.cumsum(codes == 0) - 1 + first_id
    order = np.lexsort((hits[:, 1], -hits[:, 3], hits[:, 0]))
    names = {int(r): str(int(rank[r])) for r in np.unique(hits[:, 0])}
    for line in format_m8(hits[order], eng.markers, read_length, names):
        fh.write(line + "\n")


# ------------------------------------------------------------------------------------------------ the GPU seam
class _Prefetch:
    """Batches of one input file, parsed and packed ahead of the search by a thread of their own (libmcxio does the work
    with the GIL released, on `threads` threads: the -t option).  Paired files get one of these each, so the second file is
    inflating / being parsed while the first is searched.  Under a sharded run every rank walks the whole file but only
    keeps every world-th batch (the others are skipped without being stored)."""

    def __init__(self, path, per_batch, threads, first_batch_no, world, rank, want_total):
        import queue
        import threading
        self.path, self.q = path, queue.Queue(maxsize=2)
        self.halt = threading.Event()
        self.error = None
        self.n_batches = None                        # batches this file was cut into, known once the thread is done
        self.bases_total = None
        self.records_seen = 0                        # records in the batches handed out (or skipped) so far
        self._args = (per_batch, threads, first_batch_no, world, rank, want_total)
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()

    def _run(self):
        per_batch, threads, batch_no, world, rank, want_total = self._args
        try:
            with SeqFile(self.path) as rd:
                first_record = 0
                while not rd.eof and not self.halt.is_set():
                    if batch_no % world == rank:
                        pb = rd.next_packed(per_batch, threads)
                        n = pb.n
                        item = (batch_no, first_record, pb, getattr(rd, "last_without_quality", False))
                    else:
                        n = rd.skip_packed(per_batch, threads)
                        item = (batch_no, first_record, None, False)
                    if n == 0 and rd.eof:
                        break
                    self.q.put(item)
                    first_record += n
                    self.records_seen = first_record
                    batch_no += 1
                if not rd.eof and want_total:
                    rd.skip_rest()
                if rd.eof:
                    self.bases_total = rd.bases_total
                self.n_batches = batch_no
        except BaseException as exc:              # handed to the consumer
            self.error = exc
        self.q.put(None)

    def __iter__(self):
        while True:
            item = self.q.get()
            if item is None:
                if self.error is not None:
                    raise self.error
                return
            yield item

    def stop(self):
        self.halt.set()
        while self.thread.is_alive():           # let a producer blocked on the full queue see the flag
            try:
                self.q.get(timeout=0.05)
            except Exception:
                pass
        self.thread.join()


# ------------------------------------------------------------------------------------------------ the GPU seam
def sample_and_search(args, engine=None):
    """process_seqfile + search_seqs + classify_reads + aggregate_hits (mc.py:611-620) on the GPU.

    Files are taken in order (mc.py:337: paired files are processed one after the other) and streamed: libmcxio parses
    and packs batches of records on args['threads'] threads into page-locked buffers, each batch is pushed (an
    asynchronous copy) and searched while the next one is being parsed; reading stops with the batch in which the
    `-n`-th read was kept (mc.py:356).  The device applies the filter chain too-short -> duplicate -> low-quality per read
    (the duplicate filter remembers the reads kept by earlier batches) and the `-n` cut as "first nreads kept reads";
    the counters are those of the reference loop up to the read that filled the quota.  Under torchrun (one process per
    GPU) the ranks walk the files together, rank r keeps every world-th batch, and -n / -d / the sums are made global by
    microbecensus_b200.distributed -- no rank ever holds more than its batches."""
    eng = engine or get_engine()
    if args["verbose"]:
        print("====Estimating Average Genome Size====")
        print("Sampling & trimming reads...")
    L = args["read_length"]
    fastq = args["file_type"] == "fastq"
    nreads = args["nreads"]
    threads = max(1, int(args.get("threads") or 1))
    world, rank = 1, 0
    if "torch" in sys.modules or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # (a plain single-GPU run never imports torch: the import alone costs more than searching a million reads)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(), dist.get_rank()
        except ImportError:
            pass
    dups = bool(args.get("filter_dups"))
    # in a sharded run the duplicates are settled between the ranks (the engine's own filter only sees its batches)
    eng.set_params(L, quality_offset=args.get("quality_offset") if fastq else None,
                   min_quality=args["min_quality"], mean_quality=args["mean_quality"],
                   max_unknown=args["max_unknown"], filter_dups=dups and world == 1)

    want_total = args.get("no_equivs") is False      # the CLI will ask count_bases() next: finish the files in this pass
    # optional m8-compatible dump of the reported HSPs (args["m8_out"] or $MCX_M8_OUT; single-GPU runs)
    m8_path = args.get("m8_out") or os.environ.get("MCX_M8_OUT")
    m8 = open(m8_path, "w") if m8_path and world == 1 else None
    if m8:
        m8.write("# Fields: Query\tSubject\tidentity\taln-len\tmismatch\tgap-openings\tq.start\tq.end\ts.start\ts.end\tlog(e-value)\tbit-score\n")
    per_batch = int(os.environ.get("MCX_BATCH_READS", "4000000"))
    empty = ReadBatch(np.zeros(0, np.uint8), np.zeros(1, np.int64), None if not fastq else np.zeros(0, np.uint8))

    def checked(pb, cut_short):
        if fastq and pb.n and pb.quals is None:
            raise ValueError("FASTQ input without qualities")
        if fastq and cut_short and pb.n and int(pb.lengths[-1]) >= L:
            # the file ends inside a FASTQ record: the reference yields it without qualities (mc.py:323) and fails in
            # quality_filter (mc.py:272, rec.phred() of None) -- the same error text, caught by run_pipeline the same way
            raise TypeError("'NoneType' object is not iterable")
        if not fastq and pb.quals is not None:
            pb.quals = None
        return pb

    res, remaining = None, nreads
    records_before = 0            # records of the files already finished (global read index = this + index in the file)
    batch_no = 0
    if world > 1:
        from .distributed import sharded_round, allreduce_result
    readers = [_Prefetch(path, per_batch, threads, 0, world, rank, want_total) for path in args["seqfiles"][:1]]
    try:
        for fi, path in enumerate(args["seqfiles"]):
            if remaining is not None and remaining <= 0:
                break
            rd = readers[fi]
            if fi + 1 < len(args["seqfiles"]) and len(readers) == fi + 1:
                # the next file starts being read now (its batch numbers continue where this file will end only matter for
                # the round-robin of a sharded run, which therefore reads the files one after the other)
                if world == 1:
                    readers.append(_Prefetch(args["seqfiles"][fi + 1], per_batch, threads, 0, 1, 0, want_total))
            pending = []                      # sharded: the batches of the current round
            for bno, first_record, pb, cut_short in rd:
                if world == 1:
                    eng.push(checked(pb, cut_short))
                    part = eng.search(-1 if remaining is None else remaining)
                    if m8:
                        _dump_m8(eng, m8, L, 0 if res is None else res.sampled_reads)
                    if remaining is not None:
                        remaining -= part.sampled_reads
                    res = part if res is None else _add_results(res, part)
                    if remaining is not None and remaining <= 0:
                        break
                else:
                    pending.append((bno, first_record, pb, cut_short))
                    if len(pending) == world:
                        res, remaining = _sharded_round(eng, pending, rank, records_before, remaining, dups, res, checked, empty, sharded_round)
                        pending = []
                        if remaining is not None and remaining <= 0:
                            break
            if world > 1 and pending and not (remaining is not None and remaining <= 0):
                res, remaining = _sharded_round(eng, pending, rank, records_before, remaining, dups, res, checked, empty, sharded_round)
            if remaining is not None and remaining <= 0:
                rd.stop()
            else:
                rd.thread.join()
            if world > 1 and fi + 1 < len(args["seqfiles"]) and not (remaining is not None and remaining <= 0):
                readers.append(_Prefetch(args["seqfiles"][fi + 1], per_batch, threads, 0, world, rank, want_total))
            if rd.bases_total is not None:
                _base_counts[_file_key(path)] = rd.bases_total
            records_before += rd.records_seen
    finally:
        for r in readers:
            if r.thread.is_alive():
                r.stop()
    if res is None:
        eng.push(empty)
        res = eng.search(-1)
    if world > 1:
        res = allreduce_result(res)
    if m8:
        m8.close()
    if res.sampled_reads == 0:
        sys.exit("\nError! No reads remaining after filtering!")
    args["sampled_reads"] = res.sampled_reads
    if args["verbose"]:
        print("\t%s reads shorter than %s bp and skipped" % (res.too_short, L))
        print("\t%s low quality reads found and skipped" % res.low_qual)
        print("\t%s duplicate reads found and skipped" % res.dups)
        print("\t%s reads sampled from seqfile" % res.sampled_reads)
        print("Searching reads against marker proteins...")
        print("\t%s reads hit marker proteins" % res.reads_with_hits)
        print("Filtering hits...")
    if res.reads_classified == 0:
        sys.exit("\nError: No hits to marker proteins - cannot estimate genome size! Rerun program with more reads.")
    if args["verbose"]:
        print("\t%s reads assigned to a marker protein" % res.reads_classified)
    return res.agg_hits(), res


def _add_results(res, part):
    res.load_counts_vector(res.counts_vector() + part.counts_vector())
    return res


def _sharded_round(eng, pending, rank, records_before, remaining, dups, res, checked, empty, sharded_round):
    """one round of a sharded streamed run: `pending` = the (up to world) consecutive batches of the round, of which this
    rank owns at most one (the one with a payload)"""
    mine = [p for p in pending if p[2] is not None]
    if mine:
        bno, first_record, pb, cut_short = mine[0]
        batch, first_index = checked(pb, cut_short), records_before + first_record
    else:
        batch, first_index = empty, records_before
    part, sampled = sharded_round(eng, batch, first_index, remaining, filter_dups=dups)
    res = part if res is None else _add_results(res, part)
    if remaining is not None:
        remaining -= sampled
    return res, remaining


def estimate_average_genome_size(args, paths, agg_hits):
    """Per-family AGS from the trained proportionality constants, MAD outlier cut, trained weighted mean
    (mc.py:474-512).  `paths` is unused (tables come from the marker blob) and kept for signature parity."""
    if args["verbose"]:
        print("Computing average genome size...")
    m = get_markers()
    li = m.length_index(args["read_length"])
    fam_index = {name: i for i, name in enumerate(m.fam_names)}
    estimates = {}
    for fam_id, hits in agg_hits.items():
        rate = hits / (args["sampled_reads"] * args["read_length"])
        if rate == 0:
            continue
        estimates[fam_id] = m.coeff[li, fam_index[fam_id]] / rate
    values = list(estimates.values())
    mad_estimate = mad(values)
    median_estimate = median(values)
    est_ags, sum_weights = 0, 0
    for fam_id, estimate in estimates.items():
        if abs(estimate - median_estimate) >= mad_estimate:
            continue
        weight = m.weight[li, fam_index[fam_id]]
        est_ags += estimate * weight
        sum_weights += weight
    est_ags = est_ags / sum_weights
    if args["verbose"]:
        print("\t%s bp" % str(round(est_ags, 2)))
    return est_ags


def report_results(args, est_ags, count_bases):
    """Tab-delimited report (mc.py:514-529); layout frozen."""
    with open(args["outfile"], "w") as out:
        out.write("Parameters\n")
        out.write("%s:\t%s\n" % ("metagenome", ",".join(args["seqfiles"])))
        for key, arg in (("reads_sampled", "sampled_reads"), ("trimmed_length", "read_length"),
                         ("min_quality", "min_quality"), ("mean_quality", "mean_quality"),
                         ("filter_dups", "filter_dups"), ("max_unknown", "max_unknown")):
            out.write("%s:\t%s\n" % (key, args[arg]))
        out.write("\nResults\n")
        out.write("%s:\t%s\n" % ("average_genome_size", est_ags))
        if count_bases:
            out.write("%s:\t%s\n" % ("total_bases", count_bases))
            out.write("%s:\t%s\n" % ("genome_equivalents", count_bases / est_ags))


def count_bases(args):
    """Total bp over every record of every input file (mc.py:573-584)."""
    if args["verbose"]:
        print("Computing number of genome equivalents...")
    total = 0
    for path in args["seqfiles"]:
        key = _file_key(path)
        if key not in _base_counts:                  # not read to its end by the sampling pass: count it now
            with SeqFile(path) as rd:
                _base_counts[key] = rd.skip_rest()[1]
        total += _base_counts[key]
    return total


def run_pipeline(args):
    if "verbose" in args and args["verbose"]:
        print_copyright()
    check_os()
    try:
        check_input(args)
        impute_missing_args(args)
        check_arguments(args)
        if args["verbose"]:
            print_parameters(args)
        agg_hits, _ = sample_and_search(args)
        est_ags = estimate_average_genome_size(args, None, agg_hits)
        return est_ags, args
    except Exception as error:
        print(error)

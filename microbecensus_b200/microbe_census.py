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
def sample_and_search(args, engine=None):
    """process_seqfile + search_seqs + classify_reads + aggregate_hits (mc.py:611-620) on the GPU.

    Files are taken in order and concatenated (mc.py:337: paired files are processed one after the other);
    the device applies the filter chain too-short -> low-quality per read and the `-n` cut as "first nreads
    kept reads"; counters are those of the reference loop up to the read that filled the quota."""
    eng = engine or get_engine()
    if args["verbose"]:
        print("====Estimating Average Genome Size====")
        print("Sampling & trimming reads...")
    L = args["read_length"]
    fastq = args["file_type"] == "fastq"
    eng.set_params(L, quality_offset=args.get("quality_offset") if fastq else None,
                   min_quality=args["min_quality"], mean_quality=args["mean_quality"],
                   max_unknown=args["max_unknown"], filter_dups=bool(args.get("filter_dups")))
    nreads = args["nreads"]
    # under torchrun every rank parses the input and searches its contiguous block of the read stream; -n, -d and
    # the sums are made global by microbecensus_b200.distributed (one all-reduce + two small all-gathers)
    world, rank = 1, 0
    if "torch" in sys.modules or int(os.environ.get("WORLD_SIZE", "1")) > 1:
        # (a plain single-GPU run never imports torch: the import alone costs more than searching a million reads)
        try:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                world, rank = dist.get_world_size(), dist.get_rank()
        except ImportError:
            pass

    def checked(batch):
        if fastq and batch.quals is None and batch.n:
            raise ValueError("FASTQ input without qualities")
        return batch if fastq else ReadBatch(batch.bases, batch.offsets, None)

    want_total = args.get("no_equivs") is False      # the CLI will ask count_bases() next: finish the files in this pass
    # optional m8-compatible dump of the reported HSPs (args["m8_out"] or $MCX_M8_OUT; single-GPU runs)
    m8_path = args.get("m8_out") or os.environ.get("MCX_M8_OUT")
    m8 = open(m8_path, "w") if m8_path and world == 1 else None
    if m8:
        m8.write("# Fields: Query\tSubject\tidentity\taln-len\tmismatch\tgap-openings\tq.start\tq.end\ts.start\ts.end\tlog(e-value)\tbit-score\n")
    if world > 1 or args.get("filter_dups"):
        # -d and the sharded run need the whole read stream at once (duplicates are decided over all reads)
        batch = checked(concat_batches([load_reads(f) for f in args["seqfiles"]]))
        if world > 1:
            from .distributed import sharded_search
            eng.set_params(L, quality_offset=args.get("quality_offset") if fastq else None, min_quality=args["min_quality"],
                           mean_quality=args["mean_quality"], max_unknown=args["max_unknown"], filter_dups=False)
            lo, hi = rank * batch.n // world, (rank + 1) * batch.n // world
            res = sharded_search(eng, batch.slice(lo, hi), lo, nreads=nreads, filter_dups=bool(args.get("filter_dups")))
        else:
            eng.push(batch)
            res = eng.search(-1 if nreads is None else nreads)
            if m8:
                _dump_m8(eng, m8, L, 0)
    else:
        # stream: batches of records go to the GPU as they are parsed; reading stops with the read that fills -n
        # (mc.py:356) and the additive results of the batches are summed
        per_batch = int(os.environ.get("MCX_BATCH_READS", "8000000"))
        res, remaining = None, nreads
        for path in args["seqfiles"]:
            if remaining is not None and remaining <= 0:
                break
            with SeqFile(path) as rd:
                while not rd.eof and (remaining is None or remaining > 0):
                    batch = checked(rd.next_batch(per_batch, copy=False))
                    if batch.n == 0:
                        break
                    eng.push(batch)
                    part = eng.search(-1 if remaining is None else remaining)
                    if m8:
                        _dump_m8(eng, m8, L, 0 if res is None else res.sampled_reads)
                    if remaining is not None:
                        remaining -= part.sampled_reads
                    if res is None:
                        res = part
                    else:
                        res.load_counts_vector(res.counts_vector() + part.counts_vector())
                if want_total and not rd.eof:
                    rd.skip_rest()
                if rd.eof:
                    _base_counts[_file_key(path)] = rd.bases_total
        if res is None:
            eng.push(ReadBatch(np.zeros(0, np.uint8), np.zeros(1, np.int64), None if not fastq else np.zeros(0, np.uint8)))
            res = eng.search(-1)
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

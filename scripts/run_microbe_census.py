#!/usr/bin/env python3
"""Command line of MicrobeCensus (same options and output file as the reference's
scripts/run_microbe_census.py:9-67) on top of the GPU search."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from microbecensus_b200 import microbe_census  # noqa: E402

LENGTHS = microbe_census.VALID_LENGTHS


def parse_args(argv=None):
    p = argparse.ArgumentParser(description="Estimate average genome size from metagenomic data.")
    p.add_argument("-v", dest="verbose", action="store_true", default=False, help="print program's progress to stdout (default = False)")
    p.add_argument("-V", "--version", action="version", version="MicrobeCensus (version %s)" % microbe_census.__version__)
    p.add_argument("-r", dest="rapsearch", default=None, help="accepted for compatibility; the GPU search does not use RAPsearch2")
    p.add_argument("seqfiles", metavar="SEQFILES", type=str, help="input metagenome(s), comma separated for paired files; FASTQ/FASTA, optionally .gz/.bz2")
    p.add_argument("outfile", metavar="OUTFILE", type=str, help="path to output file containing results")
    g = p.add_argument_group("Pipeline throughput (optional)")
    g.add_argument("-n", dest="nreads", type=int, default=2000000, help="number of reads to sample (default = 2000000)")
    g.add_argument("-t", dest="threads", type=int, default=1, help="accepted for compatibility (host-side parsing only)")
    g.add_argument("-e", dest="no_equivs", action="store_true", default=False, help="do not estimate the number of genome equivalents")
    q = p.add_argument_group("Quality control (optional)")
    q.add_argument("-l", dest="read_length", type=int, choices=LENGTHS, help="all reads trimmed to this length; shorter reads discarded (default = median read length)")
    q.add_argument("-q", dest="min_quality", type=int, default=-5, help="minimum base-level PHRED quality score (default = -5; no filtering)")
    q.add_argument("-m", dest="mean_quality", type=int, default=-5, help="minimum read-level PHRED quality score (default = -5; no filtering)")
    q.add_argument("-d", dest="filter_dups", action="store_true", default=False, help="filter duplicate reads (default = False)")
    q.add_argument("-u", dest="max_unknown", type=int, default=100, help="max percent of unknown bases per read (default = 100 percent; no filtering)")
    return vars(p.parse_args(argv))


if __name__ == "__main__":
    args = parse_args()
    args["seqfiles"] = args["seqfiles"].split(",")
    est_ags, args = microbe_census.run_pipeline(args)
    count_bases = None if args["no_equivs"] else microbe_census.count_bases(args)
    microbe_census.report_results(args, est_ags, count_bases)

#!/usr/bin/env python3
"""Canonical text of the marker tables -- cut-offs, coefficients, weights per (read length, family); family and length
per gene -- from (a) the reference's own map files, read the way microbe_census.py reads them (read_dic :74-88,
find_opt_pars :61-72), or (b) the packed blob the library loads.  tests/test_host.py::test_marker_blob_equals_
reference_maps compares the two line by line where the reference tree is mounted and by digest elsewhere:
   python tools/marker_digest.py /root/reference > tests/golden/marker_maps.sha256"""
import hashlib, os, sys


def lines_from_reference(ref, genes=None):
    """genes: keep only these gene records (the blob holds one record per distinct sequence, as the shipped RAPsearch2
    database does; the maps list all 19,951 genes)"""
    data = os.path.join(ref, "microbe_census", "data")
    out = []
    with open(os.path.join(data, "pars.map")) as fh:
        next(fh)
        for line in fh:
            fid, rl, cov, aaid, score, stat = line.rstrip("\n").split("\t")
            out.append("pars\t%d\t%s\t%r\t%r\t%r\t%s" % (int(rl), fid, float(cov), float(aaid), float(score), stat))
    for name in ("coefficients", "weights"):
        for line in open(os.path.join(data, name + ".map")):
            k, v = line.split()
            out.append("%s\t%s\t%r" % (name, k, float(v)))
    fam = dict(l.split() for l in open(os.path.join(data, "gene_fam.map")))        # read_dic: a repeated key keeps its last value
    length = dict(l.split() for l in open(os.path.join(data, "gene_len.map")))
    for k, v in length.items():
        if genes is None or k in genes:
            out.append("gene\t%s\t%s\t%d" % (k, fam[k], int(float(v))))
    return sorted(out)


def lines_from_blob(markers):
    out = []
    stat = {0: "hits", 1: "cov", 2: "aln"}
    for li, rl in enumerate(markers.read_lengths):
        for fi, f in enumerate(markers.fam_names):
            r = markers.pars[li, fi]
            out.append("pars\t%d\t%s\t%r\t%r\t%r\t%s" % (rl, f, float(r["min_cov"]), float(r["max_aaid"]), float(r["min_score"]), stat[int(r["stat"])]))
            out.append("coefficients\t%d_%s\t%r" % (rl, f, float(markers.coeff[li, fi])))
            out.append("weights\t%d_%s\t%r" % (rl, f, float(markers.weight[li, fi])))
    for k, name in enumerate(markers.names):
        out.append("gene\t%s\t%s\t%d" % (name, markers.fam_names[int(markers.fam[k])], int(markers.subj_len[k])))
    return sorted(out)


def digest(lines):
    return hashlib.sha256("\n".join(lines).encode()).hexdigest()


if __name__ == "__main__":
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from microbecensus_b200.markers import Markers
    ls = lines_from_reference(sys.argv[1] if len(sys.argv) > 1 else "/root/reference", set(Markers().names))
    print(digest(ls), len(ls))

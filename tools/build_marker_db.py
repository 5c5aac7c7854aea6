#!/usr/bin/env python3
"""Pack the MicrobeCensus marker data into one binary blob (``markers.mcxdb``).

Inputs (all from a MicrobeCensus checkout, default /root/reference):
  training/input/gene_fams/*.faa.gz      marker protein families (the only copy of the sequences;
                                         the shipped RAPsearch2 database is built from them, see
                                         training/search_reads.py:37-54)
  microbe_census/data/gene_fam.map       gene -> family        (read_dic, microbe_census.py:74-88)
  microbe_census/data/gene_len.map       gene -> length
  microbe_census/data/pars.map           cutoffs per (family, read length)  (find_opt_pars, :61-72)
  microbe_census/data/coefficients.map, weights.map, read_len.map

Sequences are de-duplicated by residue string (first record wins, files in sorted order), which
reproduces the shipped rapdb_2.15.info byte for byte.  Output layout (little endian):

  char[8] "MCXDB001"; int32 n_subj, n_res, n_fam, n_len, names_bytes, reserved[3]
  int32 off[n_subj+1]; uint8 fam[n_subj] (+pad4); uint8 res[n_res] (+pad4)
  int32 read_len[n_len]; char fam_name[n_fam][8]
  {f64 min_cov, f64 max_aaid, f64 min_score, i32 stat, i32 pad}[n_len][n_fam]     stat: 0 hits 1 cov 2 aln
  f64 coeff[n_len][n_fam]; f64 weight[n_len][n_fam]; char names[names_bytes] ('\n' separated)
"""
import glob, gzip, os, struct, sys
import numpy as np

AA = "ARNDCQEGHILKMFPSTWYV"

def main(ref="/root/reference", out=None):
    out = out or os.path.join(os.path.dirname(__file__), "..", "microbecensus_b200", "data", "markers.mcxdb.gz")
    data = os.path.join(ref, "microbe_census", "data")
    gene2fam = {}
    for line in open(os.path.join(data, "gene_fam.map")):
        k, v = line.split(); gene2fam[k] = v            # last entry wins, as read_dic does
    gene2len = {}
    for line in open(os.path.join(data, "gene_len.map")):
        k, v = line.split(); gene2len[k] = float(v)
    read_lens = [int(x) for x in open(os.path.join(data, "read_len.map")).read().split()]
    names, seqs, seen = [], [], set()
    for f in sorted(glob.glob(os.path.join(ref, "training", "input", "gene_fams", "*.faa.gz"))):
        hdr, buf = None, []
        def flush():
            if hdr is None: return
            s = "".join(buf)
            if s in seen: return
            seen.add(s); names.append(hdr[1:].split()[0]); seqs.append(s)
        for line in gzip.open(f, "rt"):
            line = line.rstrip("\n")
            if line.startswith(">"):
                flush(); hdr, buf = line, []
            else:
                buf.append(line)
        flush()
    fams = sorted(set(gene2fam.values()))
    assert len(fams) == 30
    fam_idx = {f: i for i, f in enumerate(fams)}
    off = np.zeros(len(seqs) + 1, np.int32)
    off[1:] = np.cumsum([len(s) for s in seqs])
    code = np.full(256, 20, np.uint8)
    for i, c in enumerate(AA): code[ord(c)] = i
    res = code[np.frombuffer("".join(seqs).encode(), np.uint8)]
    fam = np.array([fam_idx[gene2fam[n]] for n in names], np.uint8)
    for n, s in zip(names, seqs):
        assert gene2len[n] == len(s), n
    pars = {}
    with open(os.path.join(data, "pars.map")) as fh:
        next(fh)
        for line in fh:
            fid, rl, cov, aaid, sc, stat = line.split()
            pars[(int(rl), fid)] = (float(cov), float(aaid), float(sc), {"hits": 0, "cov": 1, "aln": 2}[stat])
    def table(path):
        d = {}
        for line in open(path):
            k, v = line.split(); d[k] = float(v)
        return np.array([[d["%d_%s" % (rl, f)] for f in fams] for rl in read_lens], np.float64)
    coeff = table(os.path.join(data, "coefficients.map"))
    weight = table(os.path.join(data, "weights.map"))
    names_blob = "\n".join(names).encode()
    pad4 = lambda b: b + b"\0" * (-len(b) % 4)
    blob = b"MCXDB001" + struct.pack("<8i", len(seqs), len(res), len(fams), len(read_lens), len(names_blob), 0, 0, 0)
    blob += off.tobytes() + pad4(fam.tobytes()) + pad4(res.tobytes())
    blob += np.array(read_lens, np.int32).tobytes()
    blob += b"".join(f.encode().ljust(8, b"\0") for f in fams)
    for rl in read_lens:
        for f in fams:
            blob += struct.pack("<dddii", *pars[(rl, f)], 0)
    blob += coeff.tobytes() + weight.tobytes() + names_blob
    with gzip.GzipFile(out, "wb", mtime=0) as fh:
        fh.write(blob)
    print("wrote %s: %d subjects, %d residues, %d bytes raw" % (out, len(seqs), len(res), len(blob)))

if __name__ == "__main__":
    main(*sys.argv[1:])

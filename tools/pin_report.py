#!/usr/bin/env python3
"""How closely the oracle (and therefore the CUDA path, which equals it bit for bit) reproduces the UNMODIFIED reference
on the reference's own inputs: RAPsearch2's m8 lines and the reference's classification / AGS.  Needs /root/reference and
baseline/_ref.  Prints the table of DESIGN.md section 3.   python tools/pin_report.py [quick]"""
import collections, os, subprocess, sys, tempfile, time, warnings
warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
import numpy as np
from microbe_census import microbe_census as mc
from microbecensus_b200 import microbe_census as mcb
from microbecensus_b200.engine import ReadBatch, SearchResult, format_m8, HIT_FIELDS
from microbecensus_b200.markers import Markers, report_floor
from oracle_lib import Oracle, OC_HIT_FIELDS
REF = "/root/reference"
RAP = os.path.join(REF, "microbe_census/bin/rapsearch_Linux_2.15")
DB = os.path.join(ROOT, "baseline/_ref/microbe_census/data/rapdb_2.15")
m = Markers(); orc = Oracle(m)
col = [OC_HIT_FIELDS.index(k) for k in HIT_FIELDS]

def reads_of(path, L, n=None):
    seqs = [r.seq for r in mc.parse_seqs(mc.open_file(path)) if len(r.seq) >= L]
    return [s[:L] for s in (seqs[:n] if n else seqs)]

def row(label, seqs, L):
    with tempfile.TemporaryDirectory() as tmp:
        fa = os.path.join(tmp, "r.fa")
        with open(fa, "w") as fh:
            for i, s in enumerate(seqs):
                fh.write(">%d\n%s\n" % (i, s))
        base = os.path.join(tmp, "o")
        subprocess.check_call("%s -q %s -d %s -o %s -z 8 -e 1 -t n -p f -b 0" % (RAP, fa, DB, base), shell=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
        ref = [l.rstrip("\n") for l in open(base + ".m8") if l[0] != "#"]
        paths = mc.get_relative_paths({}); paths["tempfile"] = base
        args = {"read_length": L, "verbose": False, "sampled_reads": len(seqs)}
        best = mc.classify_reads(args, paths)
        ags_ref = mc.estimate_average_genome_size(args, paths, mc.aggregate_hits(args, paths, best))
    batch = ReadBatch.from_strings(seqs)
    hits, _ = orc.search(batch, L, report_floor(L), cap=8_000_000)
    ours = set(format_m8(hits[:, col], m, L))
    key = lambda l: tuple(l.split("\t")[:2])
    ours_keys = set(key(l) for l in ours)
    single = [l for l in ref if len(l.split("\t")[10].split(".")[-1]) <= 2]    # sum-statistics lines print six digits
    found = sum(key(l) in ours_keys for l in single)
    same = sum(l in ours for l in single)
    oc = orc.classify(hits, L, m, batch.n)

    class Raw: pass
    raw = Raw()
    for k in ("too_short", "low_qual", "dups", "reads_with_hits", "n_hsp", "n_seed_hits", "n_gapped", "gapped_cells"): setattr(raw, k, 0)
    raw.sampled_reads = len(seqs); raw.reads_classified = oc["classified"]
    raw.fam_hits = oc["fam_hits"]; raw.fam_aln = oc["fam_aln"]; raw.aln_by_len = oc["aln_by_len"].ravel()
    ags = mcb.estimate_average_genome_size({"read_length": L, "sampled_reads": len(seqs), "verbose": False}, None, SearchResult(raw, m, L).agg_hits())
    gene2fam = mc.read_dic(paths["fams"], header=False, dtype="char")
    ref_cls = {int(k): v[0] for k, v in best.items()}
    our_cls = {i: m.fam_names[m.fam[s]] for i, s in enumerate(oc["best_subject"]) if s >= 0}
    common = [k for k in ref_cls if k in our_cls]
    same_fam = sum(ref_cls[k] == our_cls[k] for k in common)
    print("| %s | %d | %d / %d | %.2f %% | %d / %d (%d common, %d same family) | %.1f / %.1f (%.2f %%) |" % (
        label, len(seqs), found, len(single), 100.0 * same / max(found, 1), len(ref_cls), len(our_cls), len(common), same_fam,
        ags_ref, ags, 100.0 * abs(ags - ags_ref) / ags_ref), flush=True)

print("| input (reference's own files) | reads | RS2 single-HSP pairs found / printed | of those, identical lines (all 12 fields as text) | classified reads ref / ours | AGS ref / ours |")
print("|---|---|---|---|---|---|")
quick = len(sys.argv) > 1
meta = os.path.join(REF, "tests/data/metagenome.fa.gz")
fq = os.path.join(REF, "microbe_census/example/example.fq.gz")
fa = os.path.join(REF, "microbe_census/example/example.fa.gz")
row("`example/example.fq.gz`, 100 bp", reads_of(fq, 100), 100)
row("`example/example.fa.gz`, 150 bp", reads_of(fa, 150), 150)
row("`example/example.fa.gz`, 500 bp", reads_of(fa, 500), 500)
row("`tests/data/metagenome.fa.gz`, 50 bp, first 20k", reads_of(meta, 50, 20000), 50)
if not quick:
    row("`tests/data/metagenome.fa.gz`, 100 bp", reads_of(meta, 100), 100)

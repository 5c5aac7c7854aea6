#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference in the build container.

Needs /root/reference and the writable copy under baseline/_ref with the regenerated
data/rapdb_2.15.  For every fixture it writes
  <name>.fa.gz / <name>.fq.gz   the reads (subsets of the reference's own example/test inputs)
  <name>.L<len>.m8.gz           what rapsearch_Linux_2.15 prints for them (command line of mc.py:375)
  <name>.L<len>.json            sampled/hit/classified counts, per-read classification, agg_hits and AGS
                                from the reference's classify_reads/aggregate_hits/estimate (mc.py:432-512)
and for the FASTQ fixture the counters of process_seqfile (mc.py:328-367) under several option sets.
"""
import collections, gzip, json, os, random, subprocess, sys, tempfile, warnings

warnings.filterwarnings("ignore")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
from microbe_census import microbe_census as mc  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
RAP = os.path.join(REF, "microbe_census", "bin", "rapsearch_Linux_2.15")
DB = os.path.join(ROOT, "baseline", "_ref", "microbe_census", "data", "rapdb_2.15")


def rapsearch(fasta, out):
    subprocess.check_call("%s -q %s -d %s -o %s -z 8 -e 1 -t n -p f -b 0" % (RAP, fasta, DB, out), shell=True,
                          stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return out + ".m8"


def reference_stage(m8, L, sampled):
    paths = mc.get_relative_paths({})
    paths["tempfile"] = m8[:-3]
    args = {"read_length": L, "verbose": False, "sampled_reads": sampled}
    best = mc.classify_reads(args, paths)
    agg = mc.aggregate_hits(args, paths, best)
    ags = mc.estimate_average_genome_size(args, paths, agg)
    lines = [l for l in open(m8) if l[0] != "#"]
    gene2fam = mc.read_dic(paths["fams"], header=False, dtype="char")
    return {"read_length": L, "sampled_reads": sampled, "m8_lines": len(lines),
            "reads_with_hits": len(set(l.split("\t")[0] for l in lines)),
            "classified": {k: {"fam": v[0], "aln": v[1], "score": v[3]} for k, v in best.items()},
            "agg_hits": agg, "ags": ags}


def write_set(name, seqs, L, tmp):
    fa = os.path.join(tmp, name + ".fa")
    with open(fa, "w") as fh:
        for i, s in enumerate(seqs):
            fh.write(">%d\n%s\n" % (i, s[:L]))
    m8 = rapsearch(fa, os.path.join(tmp, "%s.L%d" % (name, L)))
    exp = reference_stage(m8, L, len(seqs))
    with gzip.GzipFile(os.path.join(GOLD, "%s.L%d.m8.gz" % (name, L)), "wb", mtime=0) as fh:
        fh.write(open(m8, "rb").read())
    json.dump(exp, open(os.path.join(GOLD, "%s.L%d.json" % (name, L)), "w"), indent=1, sort_keys=True)
    print(name, L, "reads", len(seqs), "lines", exp["m8_lines"], "classified", len(exp["classified"]), "AGS", exp["ags"])


def make_ties():
    """G5: reads of example.fa.gz whose alignments to some subjects can be grown from several seeds into alignments of
    the same score and ends but different gap placement / identity (reads 645, 1077, 1202 at 500 bp): RAPsearch2 reports
    the one from the leftmost seed.  `python tools/make_golden.py ties` writes only this fixture."""
    recs = [r.seq for r in mc.parse_seqs(mc.open_file(os.path.join(REF, "microbe_census", "example", "example.fa.gz")))]
    pick = [640, 645, 646, 1077, 1202, 1203]
    seqs = [recs[i] for i in pick]
    with tempfile.TemporaryDirectory() as tmp:
        with gzip.GzipFile(os.path.join(GOLD, "ties.fa.gz"), "wb", mtime=0) as fh:
            fh.write("".join(">e%d\n%s\n" % (i, s) for i, s in zip(pick, seqs)).encode())
        write_set("ties", seqs, 500, tmp)


def make_full():
    """G6: the reference's own inputs, whole, with what the unmodified reference makes of them (no m8: classification,
    per-family sums, counters, AGS).  `python tools/make_golden.py full`.
      full/metagenome.fa.gz  tests/data/metagenome.fa.gz, API defaults (tests/test_microbe_census.py:15-25)
      full/example.fq.gz     BASELINE config 1: CLI defaults (-n 2000000)
      full/example.fa.gz     -l 150 and -l 500
    and G7 `cap`: the reads of the metagenome with the most m8 lines (RAPsearch2 prints at most 500 per read) plus the
    classified reads the first fixtures left out (more than 60 lines), with RAPsearch2's own lines."""
    import shutil
    full = os.path.join(GOLD, "full")
    os.makedirs(full, exist_ok=True)
    srcs = {"metagenome.fa.gz": os.path.join(REF, "tests", "data", "metagenome.fa.gz"),
            "example.fq.gz": os.path.join(REF, "microbe_census", "example", "example.fq.gz"),
            "example.fa.gz": os.path.join(REF, "microbe_census", "example", "example.fa.gz")}
    for name, src in srcs.items():
        shutil.copyfile(src, os.path.join(full, name))
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for key, name, opts in (("metagenome", "metagenome.fa.gz", {}), ("example_fq", "example.fq.gz", {"nreads": 2000000}),
                                ("example_fa_150", "example.fa.gz", {"read_length": 150}), ("example_fa_500", "example.fa.gz", {"read_length": 500})):
            args = {"seqfiles": [os.path.join(full, name)], "verbose": False}
            args.update(opts)
            paths = mc.get_relative_paths(args)
            mc.impute_missing_args(args)
            mc.process_seqfile(args, paths)
            mc.search_seqs(args, paths)
            lines = [l for l in open(paths["tempfile"] + ".m8") if l[0] != "#"]
            best = mc.classify_reads(args, paths)
            agg = mc.aggregate_hits(args, paths, best)
            if key == "metagenome":
                shutil.copyfile(paths["tempfile"] + ".m8", os.path.join(tmp, "meta_full.m8"))
            mc.clean_up(paths)
            ags = mc.estimate_average_genome_size(args, paths, agg)
            total = mc.count_bases(args)
            out[key] = {"file": name, "opts": opts, "read_length": args["read_length"], "sampled_reads": args["sampled_reads"],
                        "quality_offset": args.get("quality_offset"), "m8_lines": len(lines),
                        "reads_with_hits": len(set(l.split("\t")[0] for l in lines)), "classified": {k: v[0] for k, v in best.items()},
                        "agg_hits": agg, "ags": ags, "total_bases": total}
            print(key, "sampled", args["sampled_reads"], "hits", out[key]["reads_with_hits"], "classified", len(best), "AGS", ags, "bases", total)
        json.dump(out, open(os.path.join(full, "expected.json"), "w"), indent=1, sort_keys=True)
        # ---- G7
        recs = [r.seq for r in mc.parse_seqs(mc.open_file(srcs["metagenome.fa.gz"]))]
        m8 = os.path.join(tmp, "meta_full.m8")
        nl = collections.Counter(l.split("\t")[0] for l in open(m8) if l[0] != "#")
        most = [int(k) for k, v in nl.most_common(40)]
        left_out = [int(k) for k in out["metagenome"]["classified"] if nl[k] > 60]
        pick = sorted(set(most + left_out))
        seqs = [recs[i] for i in pick]
        with gzip.GzipFile(os.path.join(GOLD, "cap.fa.gz"), "wb", mtime=0) as fh:
            fh.write("".join(">r%d\n%s\n" % (i, s) for i, s in zip(pick, seqs)).encode())
        write_set("cap", seqs, 100, tmp)
        print("cap: lines per read", sorted(nl[str(i)] for i in pick)[-10:], "left-out classified", len(left_out))


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "ties":
        return make_ties()
    if len(sys.argv) > 1 and sys.argv[1] == "full":
        return make_full()
    os.makedirs(GOLD, exist_ok=True)
    rnd = random.Random(20260101)
    with tempfile.TemporaryDirectory() as tmp:
        # ---- G1 / G4: the reference's test metagenome (tests/data/metagenome.fa.gz), 100 bp and 50 bp
        recs = [r.seq for r in mc.parse_seqs(mc.open_file(os.path.join(REF, "tests", "data", "metagenome.fa.gz")))]
        for name, L, pool in (("meta", 100, recs), ("meta", 50, recs[:20000])):
            fa = os.path.join(tmp, "all%d.fa" % L)
            with open(fa, "w") as fh:
                for i, s in enumerate(pool):
                    fh.write(">%d\n%s\n" % (i, s[:L]))
            m8 = rapsearch(fa, os.path.join(tmp, "all%d" % L))
            exp = reference_stage(m8, L, len(pool))
            nl = collections.Counter(l.split("\t")[0] for l in open(m8) if l[0] != "#")
            classified = [int(k) for k in exp["classified"] if nl[k] <= 60]
            hit = [int(k) for k in nl if nl[k] <= 60 and k not in exp["classified"]]
            rnd.shuffle(hit)
            nohit = [i for i in range(len(pool)) if str(i) not in nl]
            rnd.shuffle(nohit)
            pick = sorted(set(classified + hit[:150] + nohit[:250]))
            seqs = [pool[i] for i in pick]
            if L == 100:
                with gzip.GzipFile(os.path.join(GOLD, "meta.fa.gz"), "wb", mtime=0) as fh:
                    fh.write("".join(">r%d\n%s\n" % (i, s) for i, s in zip(pick, seqs)).encode())
                write_set("meta", seqs, 100, tmp)
            else:
                with gzip.GzipFile(os.path.join(GOLD, "meta50.fa.gz"), "wb", mtime=0) as fh:
                    fh.write("".join(">r%d\n%s\n" % (i, s) for i, s in zip(pick, seqs)).encode())
                write_set("meta50", seqs, 50, tmp)
        # ---- G3: 500 bp FASTA example, searched at 500 and 150 bp
        recs = [r.seq for r in mc.parse_seqs(mc.open_file(os.path.join(REF, "microbe_census", "example", "example.fa.gz")))][:400]
        with gzip.GzipFile(os.path.join(GOLD, "long.fa.gz"), "wb", mtime=0) as fh:
            fh.write("".join(">s%d some description\n%s\n%s\n" % (i, s[:250], s[250:]) for i, s in enumerate(recs)).encode())
        write_set("long", recs, 500, tmp)
        write_set("long", recs, 150, tmp)
        write_set("long", recs, 250, tmp)
        # ---- G2: FASTQ example with qualities: QC counters under several option sets + default search
        fq = [(r.id, r.seq, r.quality) for r in mc.parse_seqs(mc.open_file(os.path.join(REF, "microbe_census", "example", "example.fq.gz")))][:2500]
        fq_path = os.path.join(GOLD, "short.fq.gz")
        with gzip.GzipFile(fq_path, "wb", mtime=0) as fh:
            fh.write("".join("@%s\n%s\n+\n%s\n" % r for r in fq).encode())
        qc_cases = []
        for opts in ({}, {"min_quality": 10}, {"mean_quality": 30}, {"max_unknown": 0}, {"nreads": 700},
                     {"nreads": 500, "mean_quality": 28, "min_quality": 5}, {"read_length": 70, "mean_quality": 25},
                     {"read_length": 70, "nreads": 1234}):
            args = {"seqfiles": [fq_path], "verbose": False}
            args.update(opts)
            mc.impute_missing_args(args)
            paths = {"tempfile": os.path.join(tmp, "qc.tmp")}
            mc.process_seqfile(args, paths)
            kept = [l.strip() for l in open(paths["tempfile"]) if l[0] != ">"]
            import io, contextlib
            args["verbose"] = True
            buf = io.StringIO()
            with contextlib.redirect_stdout(buf):
                mc.process_seqfile(args, paths)
            nums = [int(l.split()[0]) for l in buf.getvalue().splitlines() if l.startswith("\t")]
            qc_cases.append({"opts": opts, "read_length": args["read_length"], "quality_offset": args["quality_offset"],
                             "too_short": nums[0], "low_qual": nums[1], "dups": nums[2], "sampled": nums[3],
                             "first_kept": kept[:3], "last_kept": kept[-1]})
        json.dump(qc_cases, open(os.path.join(GOLD, "short.qc.json"), "w"), indent=1, sort_keys=True)
        kept = [s for (_, s, _) in fq if len(s) >= 100]
        write_set("short", kept, 100, tmp)
    make_ties()
    print("done")


if __name__ == "__main__":
    main()

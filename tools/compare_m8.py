#!/usr/bin/env python3
"""Compare an oracle/GPU m8-style dump against a RAPsearch2 .m8 (calibration aid).

usage: compare_m8.py ref.m8 ours.m8 L [ref_root]
Reports pair-level overlap, field agreement on common pairs, and the per-family classification
each file produces through the reference's own classify_reads/aggregate_hits logic
(microbe_census.py:432-472), re-implemented inline over parsed records.
"""
import sys, collections
sys.path.insert(0, '/root/repo/baseline/_ref')
import warnings; warnings.filterwarnings('ignore')
from microbe_census import microbe_census as mc

def load(path):
    d = collections.OrderedDict()
    for line in open(path):
        if line[0] == '#': continue
        f = line.rstrip('\n').split('\t')
        rec = dict(query=f[0], target=f[1], pid=float(f[2]), aln=float(f[3]), mis=float(f[4]), gaps=float(f[5]),
                   qstart=float(f[6]), qend=float(f[7]), tstart=float(f[8]), tend=float(f[9]), evalue=float(f[10]), score=float(f[11]))
        d.setdefault((f[0], f[1]), rec)   # first line per pair
    return d

def classify(recs, L, paths):
    optpars = mc.find_opt_pars(paths['params'], L)
    gene2fam = mc.read_dic(paths['fams'], header=False, dtype='char')
    gene2len = mc.read_dic(paths['genelen'], header=False, dtype='float')
    best = collections.OrderedDict()
    for r in recs:
        r = dict(r); r['query_len'] = L; r['target_fam'] = gene2fam[r['target']]; r['target_len'] = gene2len[r['target']]
        if mc.alignment_filter(r, optpars): continue
        if r['query'] not in best or best[r['query']][3] < r['score']:
            best[r['query']] = [r['target_fam'], r['aln'], r['aln'] / r['target_len'], r['score'], r['target']]
    return best

def main():
    ref, ours, L = sys.argv[1], sys.argv[2], int(sys.argv[3])
    floor = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    paths = mc.get_relative_paths({})
    A, B = load(ref), load(ours)
    if floor:
        A = collections.OrderedDict((k, v) for k, v in A.items() if v['score'] >= floor)
        B = collections.OrderedDict((k, v) for k, v in B.items() if v['score'] >= floor)
    ka, kb = set(A), set(B)
    print("pairs: ref %d ours %d common %d ref-only %d ours-only %d" % (len(ka), len(kb), len(ka & kb), len(ka - kb), len(kb - ka)))
    ra, rb = set(k[0] for k in ka), set(k[0] for k in kb)
    print("reads with hits: ref %d ours %d common %d ref-only %d ours-only %d" % (len(ra), len(rb), len(ra & rb), len(ra - rb), len(rb - ra)))
    same_score = same_all = 0; lower = higher = 0
    for k in ka & kb:
        a, b = A[k], B[k]
        if a['score'] == b['score']:
            same_score += 1
            if all(a[x] == b[x] for x in ('aln', 'qstart', 'qend', 'tstart', 'tend', 'mis', 'gaps')) and abs(a['pid'] - b['pid']) < 1e-3: same_all += 1
        elif b['score'] > a['score']: higher += 1
        else: lower += 1
    n = max(1, len(ka & kb))
    print("common pairs: same score %d (%.1f%%), all fields same %d (%.1f%%), ours higher %d, ours lower %d" % (same_score, 100. * same_score / n, same_all, 100. * same_all / n, higher, lower))
    ca, cb = classify(A.values(), L, paths), classify(B.values(), L, paths)
    qa, qb = set(ca), set(cb)
    print("classified reads: ref %d ours %d common %d ref-only %d ours-only %d" % (len(qa), len(qb), len(qa & qb), len(qa - qb), len(qb - qa)))
    difffam = [q for q in qa & qb if ca[q][0] != cb[q][0]]
    diffval = [q for q in qa & qb if ca[q][0] == cb[q][0] and (ca[q][1] != cb[q][1] or ca[q][2] != cb[q][2])]
    print("common classified: family differs %d, aln/cov differs %d" % (len(difffam), len(diffval)))
    def agg(c):
        args = {'read_length': L}
        return mc.aggregate_hits(args, paths, collections.OrderedDict((k, v[:4]) for k, v in c.items()))
    ga, gb = agg(ca), agg(cb)
    for fam in sorted(set(ga) | set(gb)):
        x, y = ga.get(fam, 0.0), gb.get(fam, 0.0)
        flag = '' if abs(x - y) < 1e-9 else '   <--'
        print("  %s ref %.4f ours %.4f%s" % (fam, x, y, flag))
    nreads = int(sys.argv[5]) if len(sys.argv) > 5 else None
    if nreads:
        for name, g in (('ref', ga), ('ours', gb)):
            args = {'read_length': L, 'sampled_reads': nreads, 'verbose': False}
            print("AGS %s: %.3f" % (name, mc.estimate_average_genome_size(args, paths, g)))
    if '-v' in sys.argv:
        print("ref-only classified:", sorted(qa - qb, key=int)[:40])
        print("ours-only classified:", sorted(qb - qa, key=int)[:40])
        print("fam differs:", difffam[:20]); print("val differs:", diffval[:20])

if __name__ == '__main__':
    main()

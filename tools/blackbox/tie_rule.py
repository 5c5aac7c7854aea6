#!/usr/bin/env python3
"""Which of several equal-score alignments of a (query, subject) pair does RAPsearch2 v2.15 print?  Runs the binary and the
oracle (OC_DEBUG_ALN=1: every extended seed with its alignment) on reads of a reference input and scores candidate rules.
usage: tie_rule.py <fasta/fastq(.gz)> <read length> <number of reads>.  Result (round 1): longest alignment, then leftmost seed."""
import sys, os, re, subprocess, warnings, collections
warnings.filterwarnings("ignore")
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests'); sys.path.insert(0,'/root/repo/baseline/_ref')
from microbe_census import microbe_census as mc
from microbecensus_b200.engine import dna_coords
from microbecensus_b200.markers import Markers
src, L, N = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
RAP = "/root/reference/microbe_census/bin/rapsearch_Linux_2.15"
DB = "/root/repo/baseline/_ref/microbe_census/data/rapdb_2.15"
recs = [r.seq for r in mc.parse_seqs(mc.open_file(src))]
seqs=[s[:L] for s in recs if len(s)>=L][:N]
m = Markers(); nameidx={n:i for i,n in enumerate(m.names)}
tmp="/tmp/tierule"; os.makedirs(tmp,exist_ok=True)
fa=os.path.join(tmp,"r.fa")
with open(fa,"w") as fh:
    for i,s in enumerate(seqs): fh.write(">%d\n%s\n"%(i,s))
subprocess.check_call("%s -q %s -d %s -o %s -z 8 -e 1 -t n -p f -b 0"%(RAP,fa,DB,os.path.join(tmp,"o")),shell=True,stdout=subprocess.DEVNULL,stderr=subprocess.DEVNULL)
ref=collections.defaultdict(list)
for l in open(os.path.join(tmp,"o.m8")):
    if l[0]=="#": continue
    f=l.rstrip("\n").split("\t"); ref[(int(f[0]),f[1])].append(f)
code = """
import sys,warnings
warnings.filterwarnings('ignore')
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
from microbecensus_b200.engine import ReadBatch
from microbecensus_b200.markers import Markers, report_floor
from oracle_lib import Oracle
m=Markers(); o=Oracle(m)
seqs=[l.strip() for l in open(sys.argv[1]) if l[0] != '>']
o.search(ReadBatch.from_strings(seqs), int(sys.argv[2]), report_floor(int(sys.argv[2])), cap=8000000)
"""
with open(os.path.join(tmp,"err.txt"),"w") as ef:
    subprocess.run([sys.executable,"-c",code,fa,str(L)],stderr=ef,env=dict(os.environ,OC_DEBUG_ALN="1"))
cands=collections.defaultdict(set)
pat=re.compile(r"HIT read (\d+) subj (\d+) frame (\d+) score (\d+) q (\d+)-(\d+) t (\d+)-(\d+) aln (\d+) ident (\d+) gapo (\d+) seed qb (\d+) sb (\d+) len (\d+)")
for line in open(os.path.join(tmp,"err.txt")):
    mm=pat.match(line)
    if mm:
        v=list(map(int,mm.groups())); cands[(v[0],v[1])].add(tuple(v[2:]))
stats=collections.Counter(); shown=0
rules={'max_ident':lambda c:(-c[7],), 'max_aln':lambda c:(-c[6],), 'min_seed_qb':lambda c:(c[9],), 'max_seed_qb':lambda c:(-c[9],), 'max_aln_then_ident':lambda c:(-c[6],-c[7]), 'max_ident_then_aln':lambda c:(-c[7],-c[6]),'min_aln':lambda c:(c[6],),'max_seedlen':lambda c:(-c[11],), 'min_uq':lambda c:(c[9],), 'max_aln_then_min_qb':lambda c:(-c[6],c[9]), 'max_aln_then_max_qb':lambda c:(-c[6],-c[9])}
for (r,name),lines in ref.items():
    if len(lines)!=1 or name not in nameidx: continue
    cs=cands.get((r,nameidx[name]))
    if not cs: continue
    top=max(c[1] for c in cs); best=[c for c in cs if c[1]==top]
    alns=set((c[0],c[2],c[3],c[4],c[5],c[6],c[7]) for c in best)
    if len(alns)<2: continue
    f=lines[0]
    refq=(int(f[6]),int(f[7])); reft=(int(f[8]),int(f[9])); ra=int(f[3]); ri=round(float(f[2])*ra/100)
    chosen=[c for c in best if dna_coords(L,c[0],c[2],c[3])==refq and (c[4],c[5])==reft and c[6]==ra and c[7]==ri]
    if not chosen: stats['none']+=1; continue
    ch=chosen[0]; stats['cases']+=1
    for nme,k in rules.items():
        w=sorted(best,key=k)[0]
        if (w[0],w[2],w[3],w[4],w[5],w[6],w[7])==(ch[0],ch[2],ch[3],ch[4],ch[5],ch[6],ch[7]): stats[nme]+=1
    if shown<8:
        shown+=1; print("read",r,name,"ref chose",ch,"among",sorted(best))
print(stats)

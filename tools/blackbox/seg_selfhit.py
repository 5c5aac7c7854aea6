import ctypes, random, subprocess, gzip, numpy as np
lib=ctypes.CDLL('/root/repo/oracle/libmcoracle.so')
AA="ARNDCQEGHILKMFPSTWYV"
back={'A':'GCT','R':'CGT','N':'AAT','D':'GAT','C':'TGT','Q':'CAA','E':'GAA','G':'GGT','H':'CAT','I':'ATT','L':'CTG','K':'AAA','M':'ATG','F':'TTT','P':'CCG','S':'TCT','T':'ACT','W':'TGG','Y':'TAT','V':'GTT'}
names=[];seqs=[]
for line in open('uniq.faa'):
    if line[0]=='>': names.append(line[1:].split()[0])
    else: seqs.append(line.strip())
random.seed(5)
def mask(s):
    buf=(ctypes.c_uint8*200)(*[AA.index(c) if c in AA else 20 for c in s]); m=(ctypes.c_uint8*200)()
    lib.oc_seg_mask(buf,len(s),m); return ''.join('x' if m[i] else '-' for i in range(len(s)))
out=open('segexp.fa','w'); meta={}
n=0; tries=0
while n<400 and tries<200000:
    tries+=1
    si=random.randrange(len(seqs)); s=seqs[si]
    if len(s)<40: continue
    p=random.randrange(len(s)-33); w=s[p:p+33]
    if 'X' in w: continue
    mk=mask(w); c=mk.count('x')
    if c<6 or c>24: continue
    dna=''.join(back[a] for a in w)+'A'
    out.write('>%d\n%s\n'%(n,dna)); meta[str(n)]=(names[si],p,w,mk); n+=1
out.close()
import pickle; pickle.dump(meta,open('segexp.pkl','wb'))
print(n,tries)

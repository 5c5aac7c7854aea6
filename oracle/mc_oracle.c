/*
 * mc_oracle.c -- CPU oracle for the MicrobeCensus translated marker search path.
 * TEST INFRASTRUCTURE ONLY (see mc_oracle.h).  Plain C, single-threaded, written for clarity:
 * full DP matrices with an explicit traceback, sort+unique for seeds, no SIMD.
 *
 * Each function cites what it restates:  mc.py = /root/reference/microbe_census/microbe_census.py,
 * RS2 = RAPsearch2 v2.15 behaviour as pinned in SURVEY.md 3.3a (binary only, no source in the tree).
 */
#include "mc_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <math.h>

/* ------------------------------------------------------------------ scoring tables ---- */
/* RS2: 20x20 table at symbol `blosum62` of rapsearch_Linux_2.15 (.data 0x6749e0), residue
 * order ARNDCQEGHILKMFPSTWYV; '.' (stop / N-codon) scores -5 against everything. */
static const int8_t B62[20][20] = {
 { 4,-1,-2,-2, 0,-1,-1, 0,-2,-1,-1,-1,-1,-2,-1, 1, 0,-3,-2, 0},
 {-1, 5, 0,-2,-3, 1, 0,-2, 0,-3,-2, 2,-1,-3,-2,-1,-1,-3,-2,-3},
 {-2, 0, 6, 1,-3, 0, 0, 0, 1,-3,-3, 0,-2,-3,-2, 1, 0,-4,-2,-3},
 {-2,-2, 1, 6,-3, 0, 2,-1,-1,-3,-4,-1,-3,-3,-1, 0,-1,-4,-3,-3},
 { 0,-3,-3,-3, 9,-3,-4,-3,-3,-1,-1,-3,-1,-2,-3,-1,-1,-2,-2,-1},
 {-1, 1, 0, 0,-3, 5, 2,-2, 0,-3,-2, 1, 0,-3,-1, 0,-1,-2,-1,-2},
 {-1, 0, 0, 2,-4, 2, 5,-2, 0,-3,-3, 1,-2,-3,-1, 0,-1,-3,-2,-2},
 { 0,-2, 0,-1,-3,-2,-2, 6,-2,-4,-4,-2,-3,-3,-2, 0,-2,-2,-3,-3},
 {-2, 0, 1,-1,-3, 0, 0,-2, 8,-3,-3,-1,-2,-1,-2,-1,-2,-2, 2,-3},
 {-1,-3,-3,-3,-1,-3,-3,-4,-3, 4, 2,-3, 1, 0,-3,-2,-1,-3,-1, 3},
 {-1,-2,-3,-4,-1,-2,-3,-4,-3, 2, 4,-2, 2, 0,-3,-2,-1,-2,-1, 1},
 {-1, 2, 0,-1,-3, 1, 1,-2,-1,-3,-2, 5,-1,-3,-1, 0,-1,-3,-2,-2},
 {-1,-1,-2,-3,-1, 0,-2,-3,-2, 1, 2,-1, 5, 0,-2,-1,-1,-1,-1, 1},
 {-2,-3,-3,-3,-2,-3,-3,-3,-1, 0, 0,-3, 0, 6,-4,-2,-2, 1, 3,-1},
 {-1,-2,-2,-1,-3,-1,-1,-2,-2,-3,-3,-1,-2,-4, 7,-1,-1,-4,-3,-2},
 { 1,-1, 1, 0,-1, 0, 0, 0,-1,-2,-2, 0,-1,-2,-1, 4, 1,-3,-2,-2},
 { 0,-1, 0,-1,-1,-1,-1,-2,-2,-1,-1,-1,-1,-2,-1, 1, 5,-2,-2, 0},
 {-3,-3,-4,-4,-2,-2,-3,-2,-2,-3,-2,-3,-1, 1,-4,-3,-2,11, 2,-3},
 {-2,-2,-2,-3,-2,-1,-2,-3, 2,-1,-1,-2,-1, 3,-3,-2,-2, 2, 7,-1},
 { 0,-3,-3,-3,-1,-2,-2,-3,-3, 3, 1,-2, 1,-1,-2,-2, 0,-3,-1, 4}};

int oc_blosum(int a, int b) {
    if (a >= 20 || b >= 20 || a < 0 || b < 0) return -5;
    return B62[a][b];
}

/* RS2: byte table `murphy10` (.data 0x67f2c0): A | KR | EDNQ | C | G | H | ILVM | FYW | P | ST */
static const uint8_t M10[21] = {0,1,2,2,3,2,2,4,5,6,6,1,6,7,8,9,9,7,7,6,10};
int oc_murphy10(int a) { return (a >= 0 && a < 20) ? M10[a] : 10; }

#define GAP_OPEN 11   /* RS2: a gap of k columns costs 11 + k */
#define GAP_EXT   1

/* ------------------------------------------------------------------ translation ------- */
/* RS2: table `aa` (.data 0x688c00), codon index = 16*b0+4*b1+b2 with T=0,C=1,A=2,G=3;
 * stops and codons holding anything but upper-case ACGT give '.'. */
static const char CODON[65] = "FFLLSSSSYY..CC.WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG";
static const char AAORDER[21] = "ARNDCQEGHILKMFPSTWYV";

static int base_tcag(uint8_t c) {
    switch (c) { case 'T': return 0; case 'C': return 1; case 'A': return 2; case 'G': return 3; }
    return -1;
}
static int aa_code(char c) {
    const char *p = strchr(AAORDER, c);
    return (p && c) ? (int)(p - AAORDER) : OC_AA_STOP;
}

int oc_translate(const uint8_t *read, int L, int frame, uint8_t *aa) {
    int o = frame % 3, m = (L - o) / 3;
    for (int k = 0; k < m; ++k) {
        int b[3];
        for (int x = 0; x < 3; ++x) {
            int p = o + 3 * k + x;
            if (frame < 3) b[x] = base_tcag(read[p]);
            else { int c = base_tcag(read[L - 1 - p]); b[x] = c < 0 ? -1 : (c ^ 2); } /* T<->A, C<->G */
        }
        aa[k] = (b[0] < 0 || b[1] < 0 || b[2] < 0) ? OC_AA_STOP
                                                   : (uint8_t)aa_code(CODON[16 * b[0] + 4 * b[1] + b[2]]);
    }
    return m;
}

/* ------------------------------------------------------------------ SEG --------------- */
/* RS2: class Seg of the binary is Wootton & Federhen's seg.c (segseq/seqent/findlo/findhi/trim/
 * getprob/lnass/lnperm symbols; Seg::initialize stores window, locut 2.2, hicut 2.5, maxtrim 100);
 * BuildQHash instantiates it with window 12 for frames longer than 11 aa and HARD-masks the frame
 * (masked residues become 'x' in the sequence that is hashed and aligned).  '.' is not an
 * alphabet letter: it is left out of the composition.
 * The port inside RAPsearch2 never runs seg.c's getparams(): Seg::initialize (0x439650) leaves
 * downset = 0 and upset = 1 instead of 5 and 7, and segseq/seqent use those members.  So H[i] is the
 * entropy of the window STARTING at i (the last window repeated for the tail positions) and a raw
 * segment is the run of low-entropy window starts [loi, hii], not widened by the window length.
 * Verified black-box: self-hits of partly low-complexity marker windows are cut exactly where this
 * restatement puts the mask (tools/blackbox/seg_selfhit.py). */
#define SEG_WINDOW 12
#define SEG_DOWNSET 0
#define SEG_UPSET 1
#define SEG_LOCUT 2.2
#define SEG_HICUT 2.5
#define SEG_MAXTRIM 100

static double LNFAC[OC_MAX_FRAME + 32];
static double ENT_TERM[SEG_WINDOW + 1][SEG_WINDOW + 1]; /* [total][c] = -(c/total) log2(c/total) */
static int seg_ready = 0;

/* The tables are part of the spec: the CUDA path receives the same doubles from the host. */
void oc_seg_tables(double *lnfac /*200*/, double *ent_term /*13*13*/) {
    for (int i = 0; i < 200; ++i) lnfac[i] = lgamma((double)i + 1.0);
    lnfac[0] = 0.0; lnfac[1] = 0.0;
    for (int t = 0; t <= SEG_WINDOW; ++t)
        for (int c = 0; c <= SEG_WINDOW; ++c)
            ent_term[t * 13 + c] = (c == 0 || t == 0 || c > t) ? 0.0
                                   : -((double)c / (double)t) * (log((double)c / (double)t) / log(2.0));
}
static void seg_init(void) {
    if (seg_ready) return;
    double lf[200], et[169];
    oc_seg_tables(lf, et);
    for (int i = 0; i < OC_MAX_FRAME + 32; ++i) LNFAC[i] = lf[i];
    memcpy(ENT_TERM, et, sizeof et);
    seg_ready = 1;
}

/* sorted (descending) composition of s[0..len) ; returns number of counted letters */
static int state_vector(const uint8_t *s, int len, int *sv /*21*/) {
    int comp[20] = {0}, tot = 0;
    for (int i = 0; i < len; ++i) if (s[i] < 20) { comp[s[i]]++; tot++; }
    /* counting sort, descending */
    int k = 0;
    for (int c = len; c >= 1; --c) for (int a = 0; a < 20; ++a) if (comp[a] == c) sv[k++] = c;
    while (k < 21) sv[k++] = 0;
    return tot;
}
static double seg_entropy(const int *sv, int tot) {
    double e = 0.0;
    if (tot == 0) return 0.0;
    for (int i = 0; sv[i] != 0; ++i) e += ENT_TERM[tot][sv[i]];
    return e;
}
static double seg_lnperm(const int *sv, int tot) {
    double ans = LNFAC[tot];
    for (int i = 0; sv[i] != 0; ++i) ans -= LNFAC[sv[i]];
    return ans;
}
static double seg_lnass(const int *sv) {
    double ans = LNFAC[20];
    if (sv[0] == 0) return ans;
    int total = 20, cls = 1, svim1 = sv[0], svi, i = 0;
    for (;;) {
        if (++i == 20) { ans -= LNFAC[cls]; break; }
        svi = sv[i];
        if (svi == svim1) { cls++; continue; }
        total -= cls;
        ans -= LNFAC[cls];
        if (svi == 0) { ans -= LNFAC[total]; break; }
        cls = 1; svim1 = svi;
    }
    return ans;
}
static double LN20TOT[OC_MAX_FRAME + 32];
static double seg_getprob(const int *sv, int tot_len) {
    return seg_lnass(sv) + seg_lnperm(sv, tot_len) - LN20TOT[tot_len];
}

static void seg_trim(const uint8_t *s, int slen, int *leftend, int *rightend) {
    int lend = 0, rend = slen - 1, minlen = 1;
    if (slen - SEG_MAXTRIM > minlen) minlen = slen - SEG_MAXTRIM;
    double minprob = 1.0;
    int sv[21];
    for (int len = slen; len > minlen; --len) {
        for (int i = 0; i + len <= slen; ++i) {
            int tot = state_vector(s + i, len, sv);
            /* seg.c passes the window length, lnperm walks the counted letters */
            (void)tot;
            double prob = seg_getprob(sv, len);
            if (prob < minprob) { minprob = prob; lend = i; rend = len + i - 1; }
        }
    }
    *leftend += lend;
    *rightend -= (slen - rend - 1);
}

static void seg_segseq(const uint8_t *s, int slen, int offset, uint8_t *mask) {
    if (SEG_WINDOW > slen) return;
    double H[OC_MAX_FRAME + 1];
    int sv[21];
    int first = SEG_DOWNSET, last = slen - SEG_UPSET;
    for (int i = 0; i < slen; ++i) H[i] = -1.0;
    for (int i = first; i <= last; ++i) {
        int w0 = i - SEG_DOWNSET;                       /* seqent: shiftwin1 refuses to run off the end */
        if (w0 > slen - SEG_WINDOW) w0 = slen - SEG_WINDOW;
        int tot = state_vector(s + w0, SEG_WINDOW, sv);
        H[i] = seg_entropy(sv, tot);
    }
    int lowlim = first;
    for (int i = first; i <= last; ++i) {
        if (H[i] <= SEG_LOCUT && H[i] != -1.0) {
            int j, loi, hii;
            for (j = i; j >= lowlim; --j) { if (H[j] == -1.0) break; if (H[j] > SEG_HICUT) break; }
            loi = j + 1;
            for (j = i; j <= last; ++j) { if (H[j] == -1.0) break; if (H[j] > SEG_HICUT) break; }
            hii = j - 1;
            int leftend = loi - SEG_DOWNSET, rightend = hii + SEG_UPSET - 1;
            seg_trim(s + leftend, rightend - leftend + 1, &leftend, &rightend);
            if (i + SEG_UPSET - 1 < leftend) {
                int lend = loi - SEG_DOWNSET, rend = leftend - 1;
                seg_segseq(s + lend, rend - lend + 1, offset + lend, mask);
            }
            for (j = leftend; j <= rightend; ++j) mask[offset + j] = 1;
            i = hii < rightend + SEG_DOWNSET ? hii : rightend + SEG_DOWNSET;
            lowlim = i + 1;
        }
    }
}

void oc_seg_mask(const uint8_t *aa, int m, uint8_t *mask) {
    seg_init();
    if (LN20TOT[1] == 0.0) for (int i = 0; i < OC_MAX_FRAME + 32; ++i) LN20TOT[i] = (double)i * log(20.0);
    memset(mask, 0, (size_t)m);
    seg_segseq(aa, m, 0, mask);
}

/* RS2 BuildQHash (0x40d27a-0x40d2c5): residues SEG flags are overwritten with 'x' in the frame that is
 * hashed AND aligned; 'x' scores -5 like '.' (black-box: self-hits lose exactly self+5 per masked
 * residue), so a masked residue simply becomes OC_AA_STOP. */
int oc_frame(const uint8_t *read, int L, int frame, int use_seg, uint8_t *aa) {
    uint8_t mask[OC_MAX_FRAME];
    int m = oc_translate(read, L, frame, aa);
    if (use_seg) {
        oc_seg_mask(aa, m, mask);
        for (int i = 0; i < m; ++i) if (mask[i]) aa[i] = OC_AA_STOP;
    }
    return m;
}

/* ------------------------------------------------------------------ seed index -------- */
/* RS2 seeding, decoded from CHashSearch::Searching (0x415050) and ExtendSeq2Set (0x413b90):
 *  - the database is hashed by murphy10 6-mers (base-10 code, 10^6 buckets); a subject of n residues
 *    contributes positions 0..n-7 only (the shipped rapdb_2.15.info counts sum to sum(n-6));
 *  - at query position i the EXACT seed length is 6 when the 6-mer occurs at most `median` (= 1, first
 *    word of rapdb_2.15.info) times in the database, otherwise 6 + k where k <= 3 is the number of
 *    following letters needed to bring count * prod(letter frequency) down to the median
 *    (0x415ec0-0x415f71); the length actually used is max(that, previous used length - 1) when the
 *    previous position found a database match, and the position is skipped when it does not fit;
 *  - exact seeds are taken left-maximal (0x4140c0-0x414113);
 *  - ONE-SUBSTITUTION seeds: the 6-mer with its 4th, 5th or 6th letter replaced by each of the nine
 *    other letters, followed by the next 4 query letters (length 10, multipliers {10,1,100} pushed in
 *    CHashSearch::Process 0x41b5ff-0x41b748); no left-maximality test for these.
 * The ten letter frequencies are the doubles stored at the end of rapdb_2.15.info (passed in by the
 * caller; the marker blob carries them). */
#define N_PAT 4
static const int PAT_LEN[N_PAT] = {10, 10, 10, 10};
static const int PAT_WILD[N_PAT] = {3, 4, 5, 6};
/* CDbPckg median of the 10^6 bucket sizes: 75 % of the buckets are empty, so it is 0 and every
 * occupied 6-mer takes the "frequent word" branch, i.e. exact seeds are 9 letters long. */
#define DB_MEDIAN 0.0

struct oc_index {
    oc_db db;
    int64_t n[N_PAT];
    uint64_t *ent[N_PAT];    /* one-substitution words: (code << 32) | global residue position, sorted */
    int64_t n6;
    uint64_t *ent6;          /* exact 6-mers: (code << 32) | global residue position, sorted */
    int32_t *cnt6;           /* 10^6 bucket sizes */
    int64_t *beg6;           /* 10^6 + 1 bucket starts */
    double freq[10];         /* murphy10 letter frequencies of the database */
    int32_t *subj_of;        /* subject index of every residue position */
    uint8_t *red;            /* reduced letter of every residue (10 = invalid) */
};

static int cmp_u64(const void *a, const void *b) {
    uint64_t x = *(const uint64_t *)a, y = *(const uint64_t *)b;
    return x < y ? -1 : x > y;
}

/* word code of pattern p at r[0..]; -1 when a non-wildcard letter is invalid */
static int64_t word_code(const uint8_t *r, int p) {
    int64_t c = 0;
    for (int k = 0; k < PAT_LEN[p]; ++k) {
        if (k == PAT_WILD[p]) continue;
        if (r[k] >= 10) return -1;
        c = c * 10 + r[k];
    }
    return c;
}
static int32_t code6(const uint8_t *r) {
    int32_t c = 0;
    for (int k = 0; k < 6; ++k) { if (r[k] >= 10) return -1; c = c * 10 + r[k]; }
    return c;
}

oc_index *oc_index_build(const oc_db *db) {
    oc_index *ix = (oc_index *)calloc(1, sizeof *ix);
    ix->db = *db;
    int64_t nres = db->off[db->n_subj];
    ix->subj_of = (int32_t *)malloc(sizeof(int32_t) * (size_t)nres);
    ix->red = (uint8_t *)malloc((size_t)nres);
    int64_t lc[10] = {0}, ltot = 0;
    for (int s = 0; s < db->n_subj; ++s)
        for (int64_t g = db->off[s]; g < db->off[s + 1]; ++g) {
            ix->subj_of[g] = s;
            ix->red[g] = (uint8_t)oc_murphy10(db->res[g]);
            if (ix->red[g] < 10) { lc[ix->red[g]]++; ltot++; }
        }
    for (int k = 0; k < 10; ++k) ix->freq[k] = (double)lc[k] / (double)ltot;
    ix->ent6 = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nres);
    ix->cnt6 = (int32_t *)calloc(1000000, sizeof(int32_t));
    ix->beg6 = (int64_t *)calloc(1000001, sizeof(int64_t));
    for (int s = 0; s < db->n_subj; ++s)
        for (int64_t g = db->off[s]; g + 6 < db->off[s + 1]; ++g) {   /* positions 0..n-7 */
            int32_t c = code6(ix->red + g);
            if (c >= 0) { ix->ent6[ix->n6++] = ((uint64_t)c << 32) | (uint64_t)g; ix->cnt6[c]++; }
        }
    qsort(ix->ent6, (size_t)ix->n6, sizeof(uint64_t), cmp_u64);
    for (int c = 0; c < 1000000; ++c) ix->beg6[c + 1] = ix->beg6[c] + ix->cnt6[c];
    for (int p = 0; p < N_PAT; ++p) {
        ix->ent[p] = (uint64_t *)malloc(sizeof(uint64_t) * (size_t)nres);
        int64_t n = 0;
        for (int s = 0; s < db->n_subj; ++s)
            for (int64_t g = db->off[s]; g + PAT_LEN[p] <= db->off[s + 1]; ++g) {
                int64_t c = word_code(ix->red + g, p);
                if (c >= 0) ix->ent[p][n++] = ((uint64_t)c << 32) | (uint64_t)g;
            }
        qsort(ix->ent[p], (size_t)n, sizeof(uint64_t), cmp_u64);
        ix->n[p] = n;
    }
    return ix;
}
void oc_index_free(oc_index *ix) {
    if (!ix) return;
    for (int p = 0; p < N_PAT; ++p) free(ix->ent[p]);
    free(ix->ent6); free(ix->cnt6); free(ix->beg6);
    free(ix->subj_of); free(ix->red); free(ix);
}
const int32_t *oc_index_counts(const oc_index *ix) { return ix->cnt6; }
double *oc_index_freq(oc_index *ix) { return ix->freq; }
static int64_t lower_bound(const uint64_t *a, int64_t n, uint64_t key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (a[mid] < key) lo = mid + 1; else hi = mid; }
    return lo;
}

/* ------------------------------------------------------------------ seed stage -------- */
/* A seed is the maximal murphy10-identical stretch around a word hit (RS2 ExtendSeq2Set, 0x413fc4-
 * 0x414394: the word is grown to the right, then to the left, while the reduced letters of query and
 * subject agree); for the one-substitution words the stretch spans the substituted position. */
typedef struct { int32_t subj, frame, qb, sb, len; } seed_t;
static int cmp_seed(const void *a, const void *b) {
    const seed_t *x = (const seed_t *)a, *y = (const seed_t *)b;
    if (x->subj != y->subj) return x->subj < y->subj ? -1 : 1;
    if (x->frame != y->frame) return x->frame < y->frame ? -1 : 1;
    if (x->qb != y->qb) return x->qb < y->qb ? -1 : 1;
    if (x->sb != y->sb) return x->sb < y->sb ? -1 : 1;
    if (x->len != y->len) return x->len < y->len ? -1 : 1;
    return 0;
}

/* RS2 ExtendSeq2Set acceptance (0x414058-0x414073): BLOSUM62 sum over the stretch >= 11 (member
 * +0x403a0) and exact identities >= 4 (member +0x403a8). */
#define SEED_MIN_SCORE 11
#define SEED_MIN_IDENT 4

static int red_eq(uint8_t a, uint8_t b) { int x = oc_murphy10(a); return x < 10 && x == oc_murphy10(b); }

static int seed_grow(const uint8_t *q, int m, const uint8_t *t, int n, int *i, int *j, int *len,
                     int *score, int *ident) {
    while (*i + *len < m && *j + *len < n && red_eq(q[*i + *len], t[*j + *len])) ++*len;
    while (*i > 0 && *j > 0 && red_eq(q[*i - 1], t[*j - 1])) { --*i; --*j; ++*len; }
    int sc = 0, id = 0;
    for (int k = 0; k < *len; ++k) {
        sc += oc_blosum(q[*i + k], t[*j + k]);
        id += (q[*i + k] == t[*j + k] && q[*i + k] < 20);
    }
    *score = sc; *ident = id;
    return sc >= SEED_MIN_SCORE && id >= SEED_MIN_IDENT;
}

static void push_seed(seed_t **tk, int *nt, int *cap, int s, int f, int qi, int sj, int len) {
    if (*nt == *cap) { *cap *= 2; *tk = (seed_t *)realloc(*tk, sizeof(seed_t) * (size_t)*cap); }
    seed_t *x = &(*tk)[(*nt)++];
    x->subj = s; x->frame = f; x->qb = qi; x->sb = sj; x->len = len;
}

/* exact seed length at position i from the database frequency of its 6-mer; 0 = skip the position */
static int exact_seed_len(const oc_index *ix, const uint8_t *rq, int m, int i, int32_t h) {
    int remaining = m - i - 6;
    int max_extra = remaining >= 2 ? 3 : remaining + 1;
    double cnt = (double)ix->cnt6[h];
    if (!(cnt > DB_MEDIAN)) return 6;
    if (max_extra <= 1) return 7;
    int c = rq[i + 6];
    if (c >= 10) return 0;
    double x = cnt * ix->freq[c];
    if (DB_MEDIAN >= x) return 7;
    int extra = 1, pos = i + 7;
    for (;;) {
        ++extra;
        if (!(max_extra > extra)) break;
        c = rq[pos++];
        if (c >= 10) return 0;
        x *= ix->freq[c];
        if (DB_MEDIAN >= x) break;
    }
    return 6 + extra;
}

static int read_seeds(const oc_index *ix, const uint8_t *read, int L, int use_seg, seed_t **out) {
    int cap = 256, nt = 0;
    seed_t *tk = (seed_t *)malloc(sizeof(seed_t) * (size_t)cap);
    const oc_db *db = &ix->db;
    for (int f = 0; f < 6; ++f) {
        uint8_t aa[OC_MAX_FRAME], rq[OC_MAX_FRAME + 16];
        int m = oc_frame(read, L, f, use_seg, aa);
        for (int i = 0; i < m; ++i) rq[i] = (uint8_t)oc_murphy10(aa[i]);
        for (int i = m; i < m + 16; ++i) rq[i] = 10;
        int prev = 6;
        for (int i = 0; i + 6 <= m; ++i) {
            int32_t h = code6(rq + i);
            if (h < 0) continue;
            int len = exact_seed_len(ix, rq, m, i, h);
            if (len == 0) continue;
            if (prev - 1 > len) len = prev - 1;
            if (i + len > m) continue;
            if (ix->cnt6[h] > 0) {
                int nmatch = 0;
                for (int64_t k = ix->beg6[h]; k < ix->beg6[h + 1]; ++k) {
                    int64_t g = (int64_t)(ix->ent6[k] & 0xffffffffu);
                    int s = ix->subj_of[g];
                    int j = (int)(g - db->off[s]), n = db->off[s + 1] - db->off[s];
                    const uint8_t *rt = ix->red + db->off[s];
                    if (j + len > n) continue;
                    int ok = 1;
                    for (int e = 6; e < len; ++e) if (rq[i + e] >= 10 || rq[i + e] != rt[j + e]) { ok = 0; break; }
                    if (!ok) continue;
                    ++nmatch;
                    if (i > 0 && j > 0 && rq[i - 1] < 10 && rq[i - 1] == rt[j - 1]) continue; /* left-maximal only */
                    int qi = i, sj = j, sl = len, sc, id;
                    if (!seed_grow(aa, m, db->res + db->off[s], n, &qi, &sj, &sl, &sc, &id)) continue;
                    push_seed(&tk, &nt, &cap, s, f, qi, sj, sl);
                }
                prev = nmatch > 0 ? len : 6;
            }
            if (i + 10 > m) continue;
            for (int p = 0; p < N_PAT; ++p) {
                int64_t c = word_code(rq + i, p);
                if (c < 0) continue;
                int64_t k = lower_bound(ix->ent[p], ix->n[p], (uint64_t)c << 32);
                for (; k < ix->n[p] && (int64_t)(ix->ent[p][k] >> 32) == c; ++k) {
                    int64_t g = (int64_t)(ix->ent[p][k] & 0xffffffffu);
                    int s = ix->subj_of[g];
                    int j = (int)(g - db->off[s]), n = db->off[s + 1] - db->off[s];
                    const uint8_t *rt = ix->red + db->off[s];
                    int w = PAT_WILD[p];
                    if (j + 6 >= n) continue;                    /* the database hashes positions 0..n-7 */
                    if (rq[i + w] >= 10 || rq[i + w] == rt[j + w]) continue; /* the replaced letter differs */
                    int qi = i, sj = j, sl = 10, sc, id;
                    if (!seed_grow(aa, m, db->res + db->off[s], n, &qi, &sj, &sl, &sc, &id)) continue;
                    push_seed(&tk, &nt, &cap, s, f, qi, sj, sl);
                }
            }
        }
    }
    qsort(tk, (size_t)nt, sizeof(seed_t), cmp_seed);
    int u = 0;
    for (int k = 0; k < nt; ++k) if (u == 0 || cmp_seed(&tk[k], &tk[u - 1]) != 0) tk[u++] = tk[k];
    *out = tk;
    return u;
}

int oc_read_seeds(const oc_index *ix, const uint8_t *read, int L, int use_seg,
                  int32_t *subj, int32_t *frame, int32_t *qb, int32_t *sb, int32_t *len, int cap) {
    seed_t *tk; int n = read_seeds(ix, read, L, use_seg, &tk);
    for (int k = 0; k < n && k < cap; ++k) {
        subj[k] = tk[k].subj; frame[k] = tk[k].frame; qb[k] = tk[k].qb; sb[k] = tk[k].sb; len[k] = tk[k].len;
    }
    free(tk);
    return n;
}

/* ------------------------------------------------------------------ extension --------- */
/* RS2 AlignFwd / AlignBwd (0x406d30, 0x406e20; inlined again in AlignSeqs 0x413370): ungapped
 * X-drop walk along the diagonal.  The running score starts at the seed score; the walk stops
 * after a residue that leaves it below -20 or more than 8.9 (member +0x40388: 7 bits through
 * Bits2RawScoreUngapped) under the best so far.  Returns the gain over score0. */
/* RS2 CHashSearch::Search 0x418cee-0x418dca: the three cut-offs are bit scores pushed through the
 * Karlin-Altschul conversions raw = (bits*ln2 + ln K)/lambda with the BLOSUM62 sets decoded from
 * BlastStat::SetPar (ungapped lambda 0.318 K 0.134; gapped 11/1 lambda 0.267 K 0.041). */
#define UNGAP_XDROP ((7.0 * 0.6931471805599453 + -2.0099154790312257) / 0.318)    /*  8.94 */
#define UNGAP_FLOOR (-20)
#define GAP_TRIGGER ((25.0 * 0.6931471805599453 + -2.0099154790312257) / 0.318)   /* 48.17 */
#define GAP_XDROP ((15.0 * 0.6931471805599453 + -3.1941832122778293) / 0.267)     /* 26.98 */

static int ungapped_walk(const uint8_t *q, const uint8_t *t, int step, int nq, int nt, int score0,
                         int *ext, int *ident) {
    *ext = 0; *ident = 0;
    if (nq <= 0 || nt <= 0 || score0 < UNGAP_FLOOR) return 0;
    int best = score0, cur = score0, n = 0, id = 0;
    for (;;) {
        cur += oc_blosum(q[n * step], t[n * step]);
        id += (q[n * step] == t[n * step] && q[n * step] < 20);
        ++n;
        if (cur > best) { best = cur; *ext = n; *ident = id; }
        if (n >= nt || n >= nq) break;
        if (cur < UNGAP_FLOOR) break;
        if ((double)cur < (double)best - UNGAP_XDROP) break;
    }
    return best - score0;
}

/* RS2 AlignGapped (0x40a550): gapped X-drop extension from (0,0) with a free end, gap 11+k, drop 27.
 * Row-by-row with a live column window [cs, ce]; the window logic, the tie rules of the three
 * recurrences and the traceback flags follow the binary instruction for instruction, because the
 * pruning decides which cells exist.  q/t are walked with stride `step` (+1 forward, -1 backward:
 * the binary reverses the prefixes into temporaries, 0x4137b5-0x413a49).
 * Output: gain (<= 0: nothing appended), rows/cols consumed, identities, gap columns, gap runs. */
typedef struct { int gain, eq, et, ident, gapcols, gapopens, aln; } gext_t;

/* diagnostics for the tests of the CUDA path's fixed-size column window: per extension, the widest span of columns
 * [cs - 1, last column written] any row needed (index = that width, capped at 255), split by gain / no gain */
long oc_gap_width_hist[2][256];
static void gapped_xdrop(const uint8_t *q, const uint8_t *t, int step, int nQ, int nD, gext_t *g) {
    memset(g, 0, sizeof *g);
    int widest = 0;
    const int GI = GAP_OPEN, GE = GAP_EXT;
    int limit = (int)((GAP_XDROP - (double)GI) / (double)GE);
    if (nQ <= 0 || limit <= 1) return;
    int W = nD + 1;
    int *H = (int *)malloc(sizeof(int) * (size_t)W * 2), *F = H + W;
    char *M = (char *)calloc((size_t)(nQ + 1) * (size_t)W * 3, 1);
    char *EM = M + (size_t)(nQ + 1) * (size_t)W, *FM = EM + (size_t)(nQ + 1) * (size_t)W;
#define AT(mat, i, j) mat[(size_t)(i) * (size_t)W + (size_t)(j)]
    H[0] = 0; F[0] = -GI; AT(M, 0, 0) = '0';
    { int r = -GI;
      for (int j = 1; j <= limit && j <= nD; ++j) {
          r -= GE; H[j] = r; F[j] = r - GI;
          AT(M, 0, j) = AT(EM, 0, j) = (j == 1) ? 'E' : 'e'; AT(FM, 0, j) = 'D';
      } }
    int cs = 1, ce = limit, best = 0, bcol = 0, brow = 0;
    for (int i = 1; i <= nQ; ++i) {
        int diag = H[cs - 1];
        AT(M, i, cs - 1) = AT(FM, i, cs - 1) = (i == 1) ? 'D' : 'd';
        AT(EM, i, cs - 1) = (i == 1) ? 'E' : 'e';
        int v = H[cs - 1] - (GI + GE), f1 = F[cs - 1] - GE;
        if (v < f1) v = f1;
        F[cs - 1] = v; H[cs - 1] = v;
        int E = v - GI, hl = v, skip_tail = 0, j = cs;
        int qa = q[(i - 1) * step];
        if (!(cs > ce || cs > nD)) {
            for (;;) {
                int a = hl - (GI + GE), b = E - GE; char ef, ff;
                if (a >= b) { E = a; ef = 'E'; } else { E = b; ef = 'e'; }
                int c = H[j] - (GI + GE), d = F[j] - GE, Fv;
                if (c >= d) { Fv = c; ff = 'D'; } else { Fv = d; ff = 'd'; }
                int h = diag + oc_blosum(qa, t[(j - 1) * step]); char mode = 's';
                if (E > h) { h = E; mode = ef; }
                if (h < Fv) { h = Fv; mode = ff; }
                AT(M, i, j) = mode; AT(EM, i, j) = ef; AT(FM, i, j) = ff;
                diag = H[j]; H[j] = h; F[j] = Fv; hl = h;
                if (h > best) { best = h; bcol = j; brow = i; }
                else if ((double)best - GAP_XDROP > (double)h && j > bcol) {
                    if (j >= ce) { ce = j; break; }          /* 0x40b48f: fall into the tail loop */
                    ce = j; skip_tail = 1; break;            /* 0x40a9f5: straight to the next row */
                }
                ++j;
                if (j > nD || j > ce) break;
            }
        }
        { int lastw = j < nD ? j : nD; if (lastw - (cs - 1) + 1 > widest) widest = lastw - (cs - 1) + 1; }
        if (!skip_tail) {
            int cs_row = cs;
            for (int jj = ce + 1; jj <= nD; ++jj) {           /* 0x40ac45: run on by horizontal gaps */
                if (jj - (cs_row - 1) + 1 > widest) widest = jj - (cs_row - 1) + 1;
                int a = hl - (GI + GE), b = E - GE; char fl;
                if (a > b) { E = a; fl = 'E'; } else { E = b; fl = 'e'; }
                AT(M, i, jj) = AT(EM, i, jj) = fl;
                H[jj] = E; F[jj] = E - GI; hl = E;
                if (E > best) { best = E; bcol = jj; brow = i; }
                else if ((double)best - GAP_XDROP > (double)E) { ce = jj; break; }
            }
            if (cs <= bcol) {                                  /* 0x40ad0c: drop dead cells on the left */
                double thr = (double)best - GAP_XDROP;
                if (thr > (double)H[bcol]) cs = bcol;
                else for (int c = bcol - 1; c >= cs; --c) if (thr > (double)H[c]) { cs = c; break; }
            }
        }
        if (!(cs < ce)) break;
    }
    __atomic_fetch_add(&oc_gap_width_hist[best > 0][widest > 255 ? 255 : widest], 1, __ATOMIC_RELAXED);
    if (best > 0) {
        if (AT(M, brow, bcol) != 's') { fprintf(stderr, "oracle: gapped traceback does not end on a match\n"); abort(); }
        int i = brow, j = bcol; char c = 's', prev = 0;
        g->gain = best; g->eq = brow; g->et = bcol;
        while (c != '0') {
            if (i < 0 || j < 0) { fprintf(stderr, "oracle: gapped traceback ran off the matrix\n"); abort(); }
            g->aln++;
            if (c == 's') {
                if (q[(i - 1) * step] == t[(j - 1) * step] && q[(i - 1) * step] < 20) g->ident++;
                --i; --j; prev = 's'; c = AT(M, i, j);
            } else if (c == 'd' || c == 'D') {
                g->gapcols++; if (prev != 'd') g->gapopens++;
                prev = 'd'; --i; c = (c == 'D') ? AT(M, i, j) : AT(FM, i, j);
            } else {
                g->gapcols++; if (prev != 'e') g->gapopens++;
                prev = 'e'; --j; c = (c == 'E') ? AT(M, i, j) : AT(EM, i, j);
            }
        }
    }
#undef AT
    free(H); free(M);
}

/* RS2 AlignSeqs (0x413370): seed -> ungapped X-drop both ways (both walks start from the seed score) ->
 * if the ungapped total reaches 48.2 raw, gapped X-drop extensions from both HSP ends when more than
 * two residues are left on both sequences.  Fills h (score, aln, ident, mism, gapo, ranges). */
/* OC_DEBUG_ALN=1: every extended seed on stderr (tools/blackbox/tie_rule.py reads them) */
static long g_dbg_read = -1; static int g_dbg_subj = -1, g_dbg_frame = -1;
static void extend_seed(const uint8_t *q, int m, const uint8_t *t, int n, int qb, int sb, int len, oc_hit *h) {
    int score0 = 0, id0 = 0;
    for (int k = 0; k < len; ++k) { score0 += oc_blosum(q[qb + k], t[sb + k]); id0 += (q[qb + k] == t[sb + k] && q[qb + k] < 20); }
    int fe, fid, be, bid;
    int gf = ungapped_walk(q + qb + len, t + sb + len, 1, m - qb - len, n - sb - len, score0, &fe, &fid);
    int gb = ungapped_walk(q + qb - 1, t + sb - 1, -1, qb, sb, score0, &be, &bid);
    h->score = score0 + gf + gb; h->ident = id0 + fid + bid;
    h->q0 = qb - be; h->q1 = qb + len + fe - 1; h->t0 = sb - be; h->t1 = sb + len + fe - 1;
    h->aln = len + fe + be; h->gapo = 0;
    h->diag = h->q0;      /* start of the ungapped HSP: stands for the order in which RAPsearch2 finds the seeds */
    int gapcols = 0;
    if ((double)h->score >= GAP_TRIGGER) {
        /* The product bounds the subject side of a gapped extension to query length + 63 columns
         * (fixed-size DP rows on the device); RAPsearch2 has no such bound, but no alignment of the
         * reference comes near 63 net gap columns (maximum observed: 17). */
        int ql = m - (h->q1 + 1), tl = n - (h->t1 + 1);
        if (ql > 2 && tl > 2) {
            if (tl > ql + OC_GAP_SLACK) tl = ql + OC_GAP_SLACK;
            gext_t g; gapped_xdrop(q + h->q1 + 1, t + h->t1 + 1, 1, ql, tl, &g);
            if (g.gain > 0) { h->score += g.gain; h->ident += g.ident; h->q1 += g.eq; h->t1 += g.et; h->aln += g.aln; h->gapo += g.gapopens; gapcols += g.gapcols; }
        }
        ql = h->q0; tl = h->t0;
        if (ql > 2 && tl > 2) {
            if (tl > ql + OC_GAP_SLACK) tl = ql + OC_GAP_SLACK;
            gext_t g; gapped_xdrop(q + h->q0 - 1, t + h->t0 - 1, -1, ql, tl, &g);
            if (g.gain > 0) { h->score += g.gain; h->ident += g.ident; h->q0 -= g.eq; h->t0 -= g.et; h->aln += g.aln; h->gapo += g.gapopens; gapcols += g.gapcols; }
        }
    }
    h->mism = h->aln - h->ident - gapcols;
    if (getenv("OC_DEBUG_ALN"))
        fprintf(stderr, "HIT read %ld subj %d frame %d score %d q %d-%d t %d-%d aln %d ident %d gapo %d seed qb %d sb %d len %d\n", g_dbg_read,
                g_dbg_subj, g_dbg_frame, h->score, h->q0, h->q1, h->t0, h->t1, h->aln, h->ident, h->gapo, qb, sb, len);
}

/* exported for tests: extension of one seed */
void oc_extend_seed(const uint8_t *q, int m, const uint8_t *t, int n, int qb, int sb, int len, oc_hit *h) {
    extend_seed(q, m, t, n, qb, sb, len, h);
}

/* ------------------------------------------------------------------ statistics -------- */
/* RS2: bits = (0.267 S + ln(1/0.041)) / ln 2, printed with two decimals (BLOSUM62 11/1 gapped
 * Karlin-Altschul set decoded from BlastStat::SetPar). mc.py:393 parses the printed text. */
double oc_bits(int raw) {
    double b = (0.267 * (double)raw + log(1.0 / 0.041)) / log(2.0);
    char buf[64];
    snprintf(buf, sizeof buf, "%.2f", b);
    return strtod(buf, NULL);
}
int oc_min_raw_for_bits(double cutoff) {
    int s = 1;
    while (oc_bits(s) < cutoff) ++s;
    return s;
}

/* SURVEY 3.3a coordinate mapping: aa range (0-based inclusive) on a frame -> 1-based DNA coords */
void oc_dna_coords(int L, int frame, int q0, int q1, int *qs, int *qe) {
    int a0 = q0 + 1, a1 = q1 + 1;
    if (frame < 3) { *qs = 3 * (a0 - 1) + frame + 1; *qe = 3 * a1 + frame; }
    else { int o = frame - 3; *qs = L - o - 3 * (a0 - 1); *qe = L - o - 3 * a1 + 1; }
}

/* mc.py:400-418, operation for operation in IEEE double */
double oc_alignment_coverage(double query_len_bp, double qstart, double qend,
                             double tstart, double tend, double aln, double target_len) {
    double query_len = query_len_bp / 3;
    double qs = qstart < qend ? qstart : qend, qe = qstart < qend ? qend : qstart;
    double fm = fmod(qs, 3.0);
    double frame = (fm == 1.0 || fm == 2.0) ? fm : 3.0;
    double query_start = (qs + 3 - frame) / 3;
    double query_stop = (qe + 1 - frame) / 3;
    double a = tstart + 1, b = tend + 1;
    double target_start = a < b ? a : b, target_stop = a < b ? b : a;
    double x = (query_start - 1 < target_start - 1) ? query_start - 1 : target_start - 1;
    double y = aln;
    double z = (query_len - query_stop < target_len - target_stop) ? query_len - query_stop
                                                                   : target_len - target_stop;
    double maxaln = x + y + z;
    return aln / maxaln;
}

/* mc.py:420-430 on the fields RAPsearch2 would have printed */
int oc_alignment_filter(const oc_hit *h, int L, int subj_len, const oc_cutoff *c) {
    int qs, qe;
    oc_dna_coords(L, h->frame, h->q0, h->q1, &qs, &qe);
    double cov = oc_alignment_coverage((double)L, (double)qs, (double)qe, (double)h->t0, (double)h->t1,
                                       (double)h->aln, (double)subj_len);
    if (cov < c->min_cov) return 1;
    if (oc_bits(h->score) < c->min_score) return 1;
    char buf[64];
    snprintf(buf, sizeof buf, "%g", 100.0 * (double)h->ident / (double)h->aln);
    if (strtod(buf, NULL) > c->max_aaid) return 1;
    return 0;
}

/* ------------------------------------------------------------------ search of one read - */
/* All distinct HSPs of one read with score >= min_raw, ordered by (subject, score desc, frame, q0).
 * RAPsearch2 prints one line per HSP (several per subject when they are distinct; mc.py:432-453 looks
 * at every line on its own), so duplicates are dropped and everything else is kept. */
static int cmp_hit(const void *a, const void *b) {
    const oc_hit *x = (const oc_hit *)a, *y = (const oc_hit *)b;
    if (x->subject != y->subject) return x->subject < y->subject ? -1 : 1;
    if (x->score != y->score) return x->score > y->score ? -1 : 1;
    /* Different seeds can grow into alignments of one (query, subject) pair with the same score: another end (a tail
     * of net score zero), or the same ends with another gap placement and identity count.  Which one RAPsearch2
     * prints was read off its output (tools/blackbox/tie_rule.py: 71 such pairs on the reference's own inputs at 100,
     * 150 and 500 bp): the longest alignment, then the one grown from the leftmost seed -- 70 of 71; "highest
     * identity" explains 40, "leftmost seed" alone 41. */
    if (x->aln != y->aln) return x->aln > y->aln ? -1 : 1;
    if (x->diag != y->diag) return x->diag < y->diag ? -1 : 1;
    if (x->frame != y->frame) return x->frame < y->frame ? -1 : 1;
    if (x->q0 != y->q0) return x->q0 < y->q0 ? -1 : 1;
    if (x->q1 != y->q1) return x->q1 < y->q1 ? -1 : 1;
    if (x->t0 != y->t0) return x->t0 < y->t0 ? -1 : 1;
    if (x->t1 != y->t1) return x->t1 < y->t1 ? -1 : 1;
    if (x->ident != y->ident) return x->ident < y->ident ? -1 : 1;
    return 0;
}

int oc_search_read(const oc_index *ix, const uint8_t *read, int L, int W, int use_seg,
                   int min_raw, oc_hit *out, int cap, int64_t *n_tasks, int64_t *n_cells) {
    (void)W; (void)n_cells;
    seed_t *tk;
    int nt = read_seeds(ix, read, L, use_seg, &tk), nh = 0;
    const oc_db *db = &ix->db;
    uint8_t aa[6][OC_MAX_FRAME]; int m[6];
    for (int f = 0; f < 6; ++f) m[f] = oc_frame(read, L, f, use_seg, aa[f]);
    if (n_tasks) *n_tasks += nt;
    for (int k = 0; k < nt && nh < cap; ++k) {
        int s = tk[k].subj, f = tk[k].frame;
        int n = db->off[s + 1] - db->off[s];
        const uint8_t *t = db->res + db->off[s];
        oc_hit h; memset(&h, 0, sizeof h);
        g_dbg_subj = s; g_dbg_frame = f;
        extend_seed(aa[f], m[f], t, n, tk[k].qb, tk[k].sb, tk[k].len, &h);
        if (h.score < min_raw) continue;
        h.read = 0; h.subject = s; h.frame = f;
        out[nh++] = h;
    }
    free(tk);
    qsort(out, (size_t)nh, sizeof(oc_hit), cmp_hit);
    /* Several seeds usually reach the same or an overlapping HSP.  RAPsearch2 prints a second line for
     * a (query, subject) pair only for a disjoint HSP (24 such pairs in 72,457 on the test metagenome);
     * restated as: in score order, an HSP is kept when its query range (on the read) and its subject
     * range are both disjoint from every HSP already kept for that subject. */
    int u = 0, first = 0;
    for (int k = 0; k < nh; ++k) {
        if (u == 0 || out[u - 1].subject != out[k].subject) first = u;
        int keep = 1, qs, qe, ks, ke;
        oc_dna_coords(L, out[k].frame, out[k].q0, out[k].q1, &qs, &qe);
        if (qs > qe) { int t = qs; qs = qe; qe = t; }
        for (int j = first; j < u && keep; ++j) {
            oc_dna_coords(L, out[j].frame, out[j].q0, out[j].q1, &ks, &ke);
            if (ks > ke) { int t = ks; ks = ke; ke = t; }
            if (!(qe < ks || ke < qs) || !(out[k].t1 < out[j].t0 || out[j].t1 < out[k].t0)) keep = 0;
        }
        if (keep) out[u++] = out[k];
    }
    /* RAPsearch2 prints at most 500 lines per query (-v default), best E-value first.  Restated as: the 500
     * highest raw scores; among lines tied at the cut score the ones first in (subject, ...) order stay
     * (RAPsearch2's own order among equal E-values is unspecified). */
    if (u > OC_MAX_LINES) {
        int hist[2048];
        memset(hist, 0, sizeof hist);
        for (int k = 0; k < u; ++k) hist[out[k].score > 2047 ? 2047 : out[k].score]++;
        int T = 2047, above = 0;
        while (T > 0 && above + hist[T] < OC_MAX_LINES) { above += hist[T]; --T; }
        int allow = OC_MAX_LINES - above, w = 0;
        for (int k = 0; k < u; ++k) {
            int sc = out[k].score > 2047 ? 2047 : out[k].score;
            if (sc > T || (sc == T && allow-- > 0)) out[w++] = out[k];
        }
        u = w;
    }
    return u;
}

/* ------------------------------------------------------------------ classification ---- */
/* mc.py:432-472.  Hits arrive grouped by read (ascending), within a read by subject (ascending);
 * RAPsearch2's own m8 order is by E-value with unspecified tie order, so ties on score are broken
 * here by the lower subject index.  Aggregates are integers: hits, sum(aln) and sum(aln) split by
 * subject length (so that sum(aln/target_len) is formed on the host as sum_len T[len]/len). */
int64_t oc_classify(const oc_db *db, const oc_hit *hits, int64_t n_hits, int L, const oc_cutoff *cut,
                    int64_t *fam_hits, int64_t *fam_aln, int64_t *aln_by_len, int32_t *best_subject,
                    int64_t n_reads) {
    int64_t classified = 0;
    memset(fam_hits, 0, sizeof(int64_t) * 30);
    memset(fam_aln, 0, sizeof(int64_t) * 30);
    memset(aln_by_len, 0, sizeof(int64_t) * 30 * 1280);
    if (best_subject) for (int64_t r = 0; r < n_reads; ++r) best_subject[r] = -1;
    int64_t k = 0;
    while (k < n_hits) {
        int64_t e = k; const oc_hit *best = NULL;
        while (e < n_hits && hits[e].read == hits[k].read) {
            const oc_hit *h = &hits[e++];
            int fam = db->fam[h->subject];
            int slen = db->off[h->subject + 1] - db->off[h->subject];
            if (oc_alignment_filter(h, L, slen, &cut[fam])) continue;
            if (!best || best->score < h->score) best = h;
        }
        if (best) {
            int fam = db->fam[best->subject];
            int slen = db->off[best->subject + 1] - db->off[best->subject];
            fam_hits[fam] += 1; fam_aln[fam] += best->aln; aln_by_len[fam * 1280 + slen] += best->aln;
            if (best_subject && hits[k].read < n_reads) best_subject[hits[k].read] = best->subject;
            ++classified;
        }
        k = e;
    }
    return classified;
}

/* ------------------------------------------------------------------ QC ---------------- */
/* mc.py:342 (too short, on the untrimmed length) and mc.py:265-279 (on seq[:L], qual[:L]).
 * The three float comparisons of the reference are exact in integers:
 *   100*nN/len > u  <=> 100*nN > u*len ;  mean(Q) < m <=> sum(Q) < m*len  (|difference| >= 1/len). */
int oc_read_qc(const uint8_t *seq, const uint8_t *qual, int len, int L,
               int quality_offset, int min_quality, int mean_quality, int max_unknown) {
    if (len < L) return 1;
    int nN = 0;
    for (int i = 0; i < L; ++i) nN += (seq[i] == 'N');
    if (100 * nN > max_unknown * L) return 2;
    if (qual) {
        long sum = 0; int mn = 1 << 30;
        for (int i = 0; i < L; ++i) { int qv = (int)qual[i] - quality_offset; sum += qv; if (qv < mn) mn = qv; }
        if (sum < (long)mean_quality * L) return 2;
        if (mn < min_quality) return 2;
    }
    return 0;
}

/* -d: the reference keeps a set of whole untrimmed strings and probes seq and revcomp(seq)
 * (mc.py:345, 355).  Restated as a 128-bit polynomial fingerprint, canonical over strands. */
#define FP_B1 0x9E3779B97F4A7C15ull
#define FP_B2 0xC2B2AE3D27D4EB4Full
static int fp_code(uint8_t c) {
    switch (c) { case 'A': return 1; case 'C': return 2; case 'G': return 3; case 'T': return 4; case 'N': return 5; }
    return 6 + (c & 0x7f);
}
static int fp_comp(int c) { return (c >= 1 && c <= 4) ? 5 - c : c; }
void oc_fingerprint(const uint8_t *seq, int len, uint64_t fp[2]) {
    uint64_t f1 = 0, f2 = 0, r1 = 0, r2 = 0, p1 = 1, p2 = 1;
    for (int i = 0; i < len; ++i) {
        uint64_t c = (uint64_t)fp_code(seq[i]), rc = (uint64_t)fp_comp(fp_code(seq[i]));
        f1 = f1 * FP_B1 + c; f2 = f2 * FP_B2 + c;          /* sum c_i B^(len-1-i) */
        r1 += rc * p1; r2 += rc * p2; p1 *= FP_B1; p2 *= FP_B2; /* sum comp(c_i) B^i  = hash of revcomp */
    }
    f1 += (uint64_t)len * 0xD6E8FEB86659FD93ull; r1 += (uint64_t)len * 0xD6E8FEB86659FD93ull;
    if (f1 < r1 || (f1 == r1 && f2 <= r2)) { fp[0] = f1; fp[1] = f2; } else { fp[0] = r1; fp[1] = r2; }
}

/* ------------------------------------------------------------------ batch helpers ----- */
/* search reads [0,n) given as concatenated ASCII + offsets (only the first L bases of each are used);
 * hits get read = index; returns the number of hits written (stops at cap) */
int64_t oc_search_batch(const oc_index *ix, const uint8_t *bases, const int64_t *offs, int64_t n, int L,
                        int use_seg, int min_raw, oc_hit *out, int64_t cap, int64_t *n_seeds) {
    int64_t nh = 0;
    enum { PER = 65536 };
    oc_hit *buf = (oc_hit *)malloc(sizeof(oc_hit) * PER);
    for (int64_t r = 0; r < n; ++r) {
        if (offs[r + 1] - offs[r] < L) continue;
        g_dbg_read = (long)r;
        int k = oc_search_read(ix, bases + offs[r], L, 0, use_seg, min_raw, buf, PER, n_seeds, NULL);
        for (int i = 0; i < k && nh < cap; ++i) { buf[i].read = (int32_t)r; out[nh++] = buf[i]; }
    }
    free(buf);
    return nh;
}

/* mc.py:328-367 over a batch: codes per read (0 keep, 1 too short, 2 low quality, 3 duplicate, 4 = beyond
 * the -n cut) ; returns the number of sampled reads; counters[0..2] = too_short, low_qual, dups counted up
 * to the read that filled the quota, exactly as the reference's loop does. */
int64_t oc_process_reads(const uint8_t *bases, const uint8_t *quals, const int64_t *offs, int64_t n, int L,
                         int quality_offset, int min_quality, int mean_quality, int max_unknown,
                         int64_t nreads /* <0: all */, uint8_t *code, int64_t *counters) {
    return oc_process_reads_d(bases, quals, offs, n, L, quality_offset, min_quality, mean_quality, max_unknown, 0,
                              nreads, code, counters);
}

/* same with -d: a read that is long enough is skipped as a duplicate when the fingerprint of its untrimmed
 * string (either strand) equals that of a read already kept (mc.py:345, 355); the test comes before QC. */
int64_t oc_process_reads_d(const uint8_t *bases, const uint8_t *quals, const int64_t *offs, int64_t n, int L,
                           int quality_offset, int min_quality, int mean_quality, int max_unknown, int filter_dups,
                           int64_t nreads /* <0: all */, uint8_t *code, int64_t *counters) {
    int64_t sampled = 0;
    counters[0] = counters[1] = counters[2] = 0;
    size_t cap = 1; while (cap < (size_t)(2 * n + 16)) cap <<= 1;
    uint64_t *set = filter_dups ? (uint64_t *)calloc(cap * 2, sizeof(uint64_t)) : NULL;   /* (0,0) = empty */
    int64_t r = 0;
    for (; r < n; ++r) {
        const uint8_t *q = quals ? quals + offs[r] : NULL;
        int len = (int)(offs[r + 1] - offs[r]);
        int c = oc_read_qc(bases + offs[r], q, len, L, quality_offset, min_quality, mean_quality, max_unknown);
        uint64_t fp[2] = {0, 0}; size_t slot = 0;
        if (filter_dups && c != 1) {
            oc_fingerprint(bases + offs[r], len, fp);
            if (fp[0] == 0 && fp[1] == 0) fp[1] = 1;
            slot = (size_t)(fp[0] * 0x9E3779B97F4A7C15ull >> 20) & (cap - 1);
            int dup = 0;
            while (set[2 * slot] || set[2 * slot + 1]) {
                if (set[2 * slot] == fp[0] && set[2 * slot + 1] == fp[1]) { dup = 1; break; }
                slot = (slot + 1) & (cap - 1);
            }
            if (dup) c = 3;
        }
        code[r] = (uint8_t)c;
        if (c == 1) counters[0]++;
        else if (c == 2) counters[1]++;
        else if (c == 3) counters[2]++;
        else {
            if (filter_dups) { set[2 * slot] = fp[0]; set[2 * slot + 1] = fp[1]; }
            ++sampled;
            if (nreads >= 0 && sampled == nreads) { ++r; break; }
        }
    }
    for (; r < n; ++r) code[r] = 4;
    free(set);
    return sampled;
}

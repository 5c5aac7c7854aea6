/*
 * mc_oracle.h -- CPU restatement (TEST INFRASTRUCTURE, not product code) of the
 * MicrobeCensus hot path: read QC/sampling -> translated marker search -> per-read
 * classification -> per-family aggregation.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this.
 * The product path (microbecensus_b200/csrc) never links or calls it.
 *
 * Reference anchors (all under /root/reference):
 *   microbe_census/microbe_census.py:265-279  quality_filter
 *   microbe_census/microbe_census.py:328-367  process_seqfile (filter order, -n prefix rule, -d)
 *   microbe_census/microbe_census.py:369-389  search_seqs (RAPsearch2 v2.15 command line)
 *   microbe_census/microbe_census.py:400-430  alignment_coverage / alignment_filter
 *   microbe_census/microbe_census.py:432-472  classify_reads / aggregate_hits
 * The search itself lives in a third-party prebuilt binary (RAPsearch2 v2.15, Zhao, Tang & Ye
 * 2012; microbe_census/bin/rapsearch_Linux_2.15, source NOT in the reference tree).  Its
 * published algorithm (6-frame translation, SEG on the query, murphy10 reduced-alphabet
 * seeds, BLOSUM62 11/1 extension, Karlin-Altschul bit scores) is restated here and pinned
 * against outputs of that binary run in the build container (tests/golden, see
 * tools/make_golden.py).
 */
#ifndef MC_ORACLE_H
#define MC_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OC_AA_STOP 20      /* '.' : stop codon, codon with a non-ACGT base, DB 'X' */
#define OC_MAX_FRAME 168   /* 500 bp / 3 rounded up */
#define OC_MAX_LINES 500   /* RAPsearch2 -v default: lines printed per query */
#define OC_GAP_SLACK 63    /* subject columns beyond the query length in a gapped extension */

/* one alignment record (best HSP of one read x subject pair) */
typedef struct {
    int32_t read;      /* read index within the call */
    int32_t subject;   /* subject index in the marker DB */
    int32_t frame;     /* 0..2 forward offset, 3..5 reverse-complement offset */
    int32_t diag;      /* first aa (on the frame) of the ungapped HSP the alignment grew from: tie-break, see cmp_hit */
    int32_t score;     /* raw Smith-Waterman score */
    int32_t aln;       /* alignment columns (incl. gap columns) */
    int32_t ident;     /* identical residue pairs */
    int32_t mism;      /* mismatching residue pairs */
    int32_t gapo;      /* gap openings */
    int32_t q0, q1;    /* 0-based inclusive aa range on the frame */
    int32_t t0, t1;    /* 0-based inclusive aa range on the subject */
} oc_hit;

typedef struct {
    int32_t n_subj;
    const int32_t *off;    /* n_subj+1 residue offsets */
    const uint8_t *res;    /* residues, codes 0..19 in ARNDCQEGHILKMFPSTWYV order, 20 = X */
    const uint8_t *fam;    /* family index 0..29 per subject */
} oc_db;

typedef struct oc_index oc_index;

/* scoring tables */
int  oc_blosum(int a, int b);               /* 21x21 incl. OC_AA_STOP = -5 */
int  oc_murphy10(int a);                    /* reduced letter 0..9, 10 for OC_AA_STOP */

/* six-frame translation of an ASCII read trimmed to L; returns aa count */
int  oc_translate(const uint8_t *read, int L, int frame, uint8_t *aa);

/* SEG (window 12, locut 2.2, hicut 2.5, maxtrim 100): mask[i]=1 for low-complexity residues */
void oc_seg_mask(const uint8_t *aa, int m, uint8_t *mask);

/* translated + SEG-hard-masked frame (masked residues become OC_AA_STOP); returns aa count */
int  oc_frame(const uint8_t *read, int L, int frame, int use_seg, uint8_t *aa);

/* seed index over the marker DB */
oc_index *oc_index_build(const oc_db *db);
void      oc_index_free(oc_index *ix);

/* extension of one seed (ungapped X-drop, then gapped X-drop from both HSP ends): fills score, aln,
 * ident, mism, gapo, q0..q1, t0..t1 of h */
void oc_extend_seed(const uint8_t *q, int m, const uint8_t *t, int n, int qb, int sb, int len, oc_hit *h);

/* raw score -> bit score as RAPsearch2 prints it (2 decimals) */
double oc_bits(int raw);
/* smallest raw score whose printed bit score is >= cutoff */
int    oc_min_raw_for_bits(double cutoff);

/* seed stage only: distinct seeds (subject, frame, query begin, subject begin, length) of one read,
 * sorted; returns count */
int  oc_read_seeds(const oc_index *ix, const uint8_t *read, int L, int use_seg,
                   int32_t *subj, int32_t *frame, int32_t *qb, int32_t *sb, int32_t *len, int cap);

/* full search of one read: distinct HSPs with score >= min_raw, sorted by (subject, score desc);
 * W and n_cells are unused (kept for ABI stability); returns n hits */
int  oc_search_read(const oc_index *ix, const uint8_t *read, int L, int W, int use_seg,
                    int min_raw, oc_hit *out, int cap, int64_t *n_tasks, int64_t *n_cells);

/* mc.py:400-418 */
double oc_alignment_coverage(double query_len_bp, double qstart, double qend,
                             double tstart, double tend, double aln, double target_len);
/* DNA coordinates of an aa range on a frame (SURVEY 3.3a coordinate mapping) */
void oc_dna_coords(int L, int frame, int q0, int q1, int *qs, int *qe);

/* cutoffs of one family at the current read length */
typedef struct { double min_cov, max_aaid, min_score; int32_t stat; int32_t pad; } oc_cutoff;

/* mc.py:420-430: 1 = filtered out */
int  oc_alignment_filter(const oc_hit *h, int L, int subj_len, const oc_cutoff *c);

/* classify a sorted-by-read hit list: per-family hits / sum(aln) / aln-by-length table
 * (table is 30 x 1280 int64, index fam*1280+subject_len) ; returns number of classified reads.
 * best_subject[r] (optional, n_reads) receives the chosen subject or -1. */
int64_t oc_classify(const oc_db *db, const oc_hit *hits, int64_t n_hits, int L,
                    const oc_cutoff *cut /*30*/, int64_t *fam_hits /*30*/, int64_t *fam_aln /*30*/,
                    int64_t *aln_by_len /*30*1280*/, int32_t *best_subject, int64_t n_reads);

/* mc.py:265-279 + 342-356: per-read verdict. codes: 0 keep, 1 too_short, 2 low_qual (dup handled by caller) */
int  oc_read_qc(const uint8_t *seq, const uint8_t *qual /*nullable*/, int len, int L,
                int quality_offset, int min_quality, int mean_quality, int max_unknown);

/* batch helpers used by tests/ and bench.py's cpu_baseline leg */
int64_t oc_search_batch(const oc_index *ix, const uint8_t *bases, const int64_t *offs, int64_t n, int L,
                        int use_seg, int min_raw, oc_hit *out, int64_t cap, int64_t *n_seeds);
int64_t oc_process_reads(const uint8_t *bases, const uint8_t *quals, const int64_t *offs, int64_t n, int L,
                         int quality_offset, int min_quality, int mean_quality, int max_unknown,
                         int64_t nreads, uint8_t *code, int64_t *counters);

int64_t oc_process_reads_d(const uint8_t *bases, const uint8_t *quals, const int64_t *offs, int64_t n, int L,
                           int quality_offset, int min_quality, int mean_quality, int max_unknown, int filter_dups,
                           int64_t nreads, uint8_t *code, int64_t *counters);

/* 128-bit canonical fingerprint of the UNTRIMMED read (min over strand), for -d */
void oc_fingerprint(const uint8_t *seq, int len, uint64_t fp[2]);

#ifdef __cplusplus
}
#endif
#endif

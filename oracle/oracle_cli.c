/*
 * oracle_cli.c -- command-line driver around mc_oracle.c (TEST INFRASTRUCTURE).
 * Reads a marker blob (tools/build_marker_db.py, un-gzipped) and a FASTA of already trimmed
 * reads (what process_seqfile writes, mc.py:352), runs the oracle search and prints one
 * RAPsearch2-style m8 line per (read, subject) so that the output can be diffed against the
 * reference binary's .m8 (mc.py:375 command line).  pthreads over reads (reads are independent,
 * like rapsearch -z); results are printed in read order.
 */
#include "mc_oracle.h"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <time.h>
#include <unistd.h>

typedef struct { char *buf; size_t len; } blob_t;
static blob_t slurp(const char *path) {
    blob_t b = {0, 0};
    FILE *f = fopen(path, "rb");
    if (!f) { perror(path); exit(2); }
    fseek(f, 0, SEEK_END); b.len = (size_t)ftell(f); fseek(f, 0, SEEK_SET);
    b.buf = (char *)malloc(b.len + 1);
    if (fread(b.buf, 1, b.len, f) != b.len) { perror("read"); exit(2); }
    b.buf[b.len] = 0; fclose(f);
    return b;
}

enum { CAP = 32768, CHUNK = 64 };
typedef struct {
    oc_index *ix; char **seq; size_t nreads; int L, W, use_seg, min_raw;
    oc_hit **res; int *nres; size_t next; int64_t tasks, cells; pthread_mutex_t mu;
} work_t;
static void *worker(void *arg) {
    work_t *w = (work_t *)arg;
    oc_hit *buf = (oc_hit *)malloc(sizeof(oc_hit) * CAP);
    int64_t tasks = 0, cells = 0;
    for (;;) {
        pthread_mutex_lock(&w->mu);
        size_t b = w->next; w->next += CHUNK;
        pthread_mutex_unlock(&w->mu);
        if (b >= w->nreads) break;
        size_t e = b + CHUNK < w->nreads ? b + CHUNK : w->nreads;
        for (size_t r = b; r < e; ++r) {
            if ((int)strlen(w->seq[r]) < w->L) continue;
            int n = oc_search_read(w->ix, (const uint8_t *)w->seq[r], w->L, w->W, w->use_seg, w->min_raw, buf, CAP, &tasks, &cells);
            if (n) { w->res[r] = (oc_hit *)malloc(sizeof(oc_hit) * (size_t)n); memcpy(w->res[r], buf, sizeof(oc_hit) * (size_t)n); }
            w->nres[r] = n;
        }
    }
    pthread_mutex_lock(&w->mu); w->tasks += tasks; w->cells += cells; pthread_mutex_unlock(&w->mu);
    free(buf);
    return NULL;
}

int main(int argc, char **argv) {
    if (argc < 6) {
        fprintf(stderr, "usage: %s markers.mcxdb reads.fa L W use_seg [min_raw] [out.m8]\n", argv[0]);
        return 2;
    }
    blob_t db_blob = slurp(argv[1]);
    int L = atoi(argv[3]), W = atoi(argv[4]), use_seg = atoi(argv[5]);
    int min_raw = argc > 6 ? atoi(argv[6]) : 1;
    FILE *out = argc > 7 ? fopen(argv[7], "w") : stdout;
    const int32_t *hdr = (const int32_t *)(db_blob.buf + 8);
    int n_subj = hdr[0], n_res = hdr[1], n_fam = hdr[2], n_len = hdr[3], names_bytes = hdr[4];
    const char *p = db_blob.buf + 8 + 32;
    oc_db db; db.n_subj = n_subj;
    db.off = (const int32_t *)p; p += 4 * (size_t)(n_subj + 1);
    db.fam = (const uint8_t *)p; p += (size_t)((n_subj + 3) & ~3);
    db.res = (const uint8_t *)p; p += (size_t)((n_res + 3) & ~3);
    p += 4 * (size_t)n_len + 8 * (size_t)n_fam + 32 * (size_t)n_len * n_fam + 16 * (size_t)n_len * n_fam;
    char *names_blob = (char *)malloc((size_t)names_bytes + 1);
    memcpy(names_blob, p, (size_t)names_bytes); names_blob[names_bytes] = 0;
    char **names = (char **)malloc(sizeof(char *) * (size_t)n_subj);
    { char *s = names_blob; for (int i = 0; i < n_subj; ++i) { names[i] = s; char *e = strchr(s, '\n'); if (e) { *e = 0; s = e + 1; } } }

    /* reads */
    blob_t fa = slurp(argv[2]);
    size_t nreads = 0, capr = 1 << 16;
    char **seq = (char **)malloc(sizeof(char *) * capr), **rid = (char **)malloc(sizeof(char *) * capr);
    for (char *s = fa.buf; s && *s;) {
        char *e = strchr(s, '\n'); if (e) *e = 0;
        if (*s == '>') {
            if (nreads == capr) { capr *= 2; seq = (char **)realloc(seq, sizeof(char *) * capr); rid = (char **)realloc(rid, sizeof(char *) * capr); }
            rid[nreads] = s + 1; seq[nreads] = e ? e + 1 : s; ++nreads;
        }
        s = e ? e + 1 : NULL;
    }
    oc_index *ix = oc_index_build(&db);
    oc_hit **res = (oc_hit **)calloc(nreads, sizeof(oc_hit *));
    int *nres = (int *)calloc(nreads, sizeof(int));
    int nthreads = getenv("ORACLE_THREADS") ? atoi(getenv("ORACLE_THREADS")) : (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (nthreads < 1) nthreads = 1;
    work_t wk = {ix, seq, nreads, L, W, use_seg, min_raw, res, nres, 0, 0, 0, PTHREAD_MUTEX_INITIALIZER};
    struct timespec ts0, ts1; clock_gettime(CLOCK_MONOTONIC, &ts0);
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, worker, &wk);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    clock_gettime(CLOCK_MONOTONIC, &ts1);
    double t0 = 0, t1 = (double)(ts1.tv_sec - ts0.tv_sec) + 1e-9 * (double)(ts1.tv_nsec - ts0.tv_nsec);
    int64_t tasks = wk.tasks, cells = wk.cells;
    long lines = 0, reads_hit = 0;
    for (size_t r = 0; r < nreads; ++r) {
        if (nres[r]) ++reads_hit;
        for (int k = 0; k < nres[r]; ++k) {
            oc_hit *h = &res[r][k]; int qs, qe;
            oc_dna_coords(L, h->frame, h->q0, h->q1, &qs, &qe);
            fprintf(out, "%s\t%s\t%g\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%d\t%.2f\t%d\t%d\t%d\n", rid[r], names[h->subject],
                    100.0 * h->ident / h->aln, h->aln, h->mism, h->gapo, qs, qe, h->t0, h->t1, 0, oc_bits(h->score),
                    h->score, h->frame, h->diag);
            ++lines;
        }
    }
    fprintf(stderr, "oracle: %zu reads, %ld with hits, %ld lines, %lld tasks, %lld cells, %.2f s (%d threads)\n",
            nreads, reads_hit, lines, (long long)tasks, (long long)cells, t1 - t0, nthreads);
    if (out != stdout) fclose(out);
    return 0;
}

/*
 * mcxio.h -- C ABI of libmcxio.so: streaming FASTA/FASTQ reader that feeds libmcx (host side, no CUDA).
 *
 * What it replaces in the reference (/root/reference/microbe_census/microbe_census.py = mc.py):
 *
 *   mcxio_open / mcxio_open_mem   open_file() (mc.py:47-59): plain text or gzip by magic bytes; bz2 input is
 *                                 inflated by the caller and handed over with mcxio_open_mem
 *   mcxio_next_batch              parse_seqs() (mc.py:294-325), the readfq generator: same record boundaries for
 *                                 multi-line FASTA, multi-line FASTQ, '+' lines, qualities longer than the
 *                                 sequence, truncated files, universal newlines ('\n', '\r\n', '\r') and the
 *                                 l[:-1] quirk that drops the last character of a final line without a newline.
 *                                 Records come back packed the way mcx_push_reads takes them
 *                                 (bases / quals / offsets), not as Python objects.
 *   mcxio_skip_rest               count_bases() / read_seqfile() (mc.py:540-584): total length of ALL records of
 *                                 the file, folded into the same pass instead of a second one
 *
 * Decompression (zlib inflate or read()) runs in a producer thread, parsing in the caller's thread.
 * All functions return 0 or a negative MCXIO_E* code; nothing throws across the ABI.
 */
#ifndef MCXIO_H
#define MCXIO_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCXIO_OK       0
#define MCXIO_EINVAL  -1
#define MCXIO_EIO     -2   /* open/read error */
#define MCXIO_EFORMAT -3   /* corrupt gzip stream / unsupported compression */
#define MCXIO_ENOMEM  -4

typedef struct mcxio_file mcxio_file;

/* Records parsed by one mcxio_next_batch call.  The pointers belong to the reader and stay valid until the next
 * call on the same reader. */
typedef struct {
    const uint8_t *bases;    /* concatenated sequences */
    const uint8_t *quals;    /* same layout; quality characters cut to the sequence length, '~' where a record
                                had none; NULL when no record of the batch had qualities */
    const int64_t *offsets;  /* n + 1 */
    int64_t n;               /* records in this batch */
    int64_t records_total;   /* records parsed from the file so far */
    int64_t bases_total;     /* sum of their sequence lengths */
    int32_t eof;             /* 1 when the parser has reached the end of the file */
} mcxio_batch;

/* Records in the layout libmcx keeps them in HBM (include/mcx.h, mcx_push_reads_packed): bit-planes of 2-bit bases +
 * mask, lengths, quality bytes.  The three buffers come from the allocator given to mcxio_set_allocator (page-locked
 * memory from mcx_host_alloc makes the push an asynchronous DMA; default malloc) and belong to the caller, who
 * releases them with mcxio_free_packed. */
typedef struct {
    uint32_t *packed;        /* n_words words */
    uint32_t *lengths;       /* n */
    uint8_t  *quals;         /* n_bases bytes, or NULL when no record of the batch had qualities */
    int64_t n, n_words, n_bases;
    int64_t records_total, bases_total;
    int32_t eof;
    int32_t last_without_quality;   /* the file ended inside a FASTQ record: its last record has no qualities (mc.py:323) */
    int64_t reparsed;        /* diagnostic: windows in which the parallel reader had to fall back to one thread */
} mcxio_packed;

int  mcxio_open(mcxio_file **out, const char *path);
int  mcxio_open_mem(mcxio_file **out, const uint8_t *data, int64_t n);   /* caller keeps `data` alive */
/* parse up to max_records further records (max_records < 0: to the end of the file) */
int  mcxio_next_batch(mcxio_file *f, int64_t max_records, mcxio_batch *out);
/* The same records, about target_records of them (< 0: the rest of the file; a plain file may return somewhat more or
 * fewer -- it is cut by bytes), parsed by `threads` threads where the input allows (plain files: pieces of the file
 * in parallel, each checked against the sequential state machine's own position; gzip / memory: one parser) and packed
 * by `threads` threads.  Do not mix with mcxio_next_batch / mcxio_skip_rest on a plain file. */
int  mcxio_next_packed(mcxio_file *f, int64_t target_records, int threads, mcxio_packed *out);
/* advance by the records the same mcxio_next_packed call would have returned, without storing them (*n_skipped = how
 * many): ranks of a sharded run walk the file together and keep every world-th batch */
int  mcxio_skip_packed(mcxio_file *f, int64_t target_records, int threads, int64_t *n_skipped);
int  mcxio_state(mcxio_file *f, int64_t *records_total, int64_t *bases_total, int32_t *eof);
void mcxio_free_packed(mcxio_packed *b);
int  mcxio_set_allocator(int (*alloc)(void **, size_t), void (*free_)(void *));
/* parse to the end of the file without storing anything; *records / *bases receive the file totals */
int  mcxio_skip_rest(mcxio_file *f, int64_t *records, int64_t *bases);
void mcxio_close(mcxio_file *f);
const char *mcxio_last_error(mcxio_file *f);

#ifdef __cplusplus
}
#endif
#endif

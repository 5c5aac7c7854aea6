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

int  mcxio_open(mcxio_file **out, const char *path);
int  mcxio_open_mem(mcxio_file **out, const uint8_t *data, int64_t n);   /* caller keeps `data` alive */
/* parse up to max_records further records (max_records < 0: to the end of the file) */
int  mcxio_next_batch(mcxio_file *f, int64_t max_records, mcxio_batch *out);
/* parse to the end of the file without storing anything; *records / *bases receive the file totals */
int  mcxio_skip_rest(mcxio_file *f, int64_t *records, int64_t *bases);
void mcxio_close(mcxio_file *f);
const char *mcxio_last_error(mcxio_file *f);

#ifdef __cplusplus
}
#endif
#endif

// mcxio.cpp -- libmcxio.so: streaming FASTA/FASTQ reader for the MicrobeCensus hot path (host side).
//
// Restates parse_seqs() of the reference (mc.py:294-325, the readfq generator) as a line-driven state machine over
// a byte stream, and open_file() (mc.py:47-59) as a producer thread that read()s or inflates (zlib, multi-member
// gzip) 8 MB chunks ahead of the parser.  Output is the packed layout mcx_push_reads takes.  See include/mcxio.h.
#include "../../include/mcxio.h"

#include <zlib.h>
#include <dlfcn.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#if defined(__SSE2__)
#include <emmintrin.h>
#endif

#include <algorithm>
#include <atomic>
#include <climits>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr size_t CHUNK = 8u << 20;
constexpr size_t QUEUE_DEPTH = 4;

struct Chunk { std::vector<uint8_t> data; size_t n = 0; };

// producer: fills chunks with decompressed bytes
struct Source {
    FILE *fp = nullptr;
    bool gz = false;
    std::thread th;
    std::mutex mu;
    std::condition_variable cv_full, cv_free;
    std::deque<Chunk> full, free_;
    bool done = false, stop = false;
    int err = 0;
    std::string msg;

    void fail(int code, const std::string &m) {
        std::lock_guard<std::mutex> g(mu);
        err = code; msg = m; done = true;
        cv_full.notify_all();
    }
    bool get_free(Chunk &c) {
        std::unique_lock<std::mutex> g(mu);
        cv_free.wait(g, [&] { return stop || !free_.empty() || full.size() < QUEUE_DEPTH; });
        if (stop) return false;
        if (!free_.empty()) { c = std::move(free_.front()); free_.pop_front(); }
        else c = Chunk();
        if (c.data.size() < CHUNK) c.data.resize(CHUNK);
        c.n = 0;
        return true;
    }
    void put_full(Chunk &&c) {
        std::lock_guard<std::mutex> g(mu);
        full.push_back(std::move(c));
        cv_full.notify_one();
    }
    void run_plain() {
        for (;;) {
            Chunk c;
            if (!get_free(c)) return;
            const size_t n = fread(c.data.data(), 1, CHUNK, fp);
            if (n == 0) {
                if (ferror(fp)) { fail(MCXIO_EIO, "read error"); return; }
                break;
            }
            c.n = n;
            put_full(std::move(c));
        }
        std::lock_guard<std::mutex> g(mu);
        done = true;
        cv_full.notify_all();
    }
    void run_gzip() {
        z_stream zs;
        memset(&zs, 0, sizeof zs);
        if (inflateInit2(&zs, 15 + 32) != Z_OK) { fail(MCXIO_ENOMEM, "inflateInit2 failed"); return; }
        std::vector<uint8_t> in(1u << 20);
        bool in_eof = false, member_open = false;
        Chunk c;
        if (!get_free(c)) { inflateEnd(&zs); return; }
        zs.next_out = c.data.data(); zs.avail_out = (uInt)CHUNK;
        for (;;) {
            if (zs.avail_in == 0 && !in_eof) {
                const size_t n = fread(in.data(), 1, in.size(), fp);
                if (n == 0) {
                    if (ferror(fp)) { inflateEnd(&zs); fail(MCXIO_EIO, "read error"); return; }
                    in_eof = true;
                }
                zs.next_in = in.data(); zs.avail_in = (uInt)n;
            }
            if (zs.avail_in == 0 && in_eof) {
                if (member_open) { inflateEnd(&zs); fail(MCXIO_EFORMAT, "gzip stream ends inside a member"); return; }
                break;
            }
            member_open = true;
            const int rc = inflate(&zs, Z_NO_FLUSH);
            if (rc == Z_STREAM_END) {
                member_open = false;
                // gzip files may hold several members back to back (gzip.open reads them all); trailing zero
                // padding is tolerated the way Python's gzip tolerates it
                while (zs.avail_in > 0 && *zs.next_in == 0) { ++zs.next_in; --zs.avail_in; }
                if (inflateReset(&zs) != Z_OK) { inflateEnd(&zs); fail(MCXIO_EFORMAT, "inflateReset failed"); return; }
            } else if (rc != Z_OK && rc != Z_BUF_ERROR) {
                const std::string m = std::string("corrupt gzip stream: ") + (zs.msg ? zs.msg : "inflate error");
                inflateEnd(&zs);
                fail(MCXIO_EFORMAT, m);
                return;
            }
            if (zs.avail_out == 0) {
                c.n = CHUNK;
                put_full(std::move(c));
                if (!get_free(c)) { inflateEnd(&zs); return; }
                zs.next_out = c.data.data(); zs.avail_out = (uInt)CHUNK;
            }
        }
        c.n = CHUNK - zs.avail_out;
        if (c.n) put_full(std::move(c));
        inflateEnd(&zs);
        std::lock_guard<std::mutex> g(mu);
        done = true;
        cv_full.notify_all();
    }
    void start() { th = std::thread([this] { gz ? run_gzip() : run_plain(); }); }
    // consumer side: next chunk or false at the end of the stream (or on error: err is set)
    bool next(Chunk &c) {
        std::unique_lock<std::mutex> g(mu);
        cv_full.wait(g, [&] { return !full.empty() || done; });
        if (full.empty()) return false;
        c = std::move(full.front());
        full.pop_front();
        cv_free.notify_one();
        return true;
    }
    void recycle(Chunk &&c) {
        std::lock_guard<std::mutex> g(mu);
        if (free_.size() < QUEUE_DEPTH) free_.push_back(std::move(c));
        cv_free.notify_one();
    }
    ~Source() {
        {
            std::lock_guard<std::mutex> g(mu);
            stop = true;
            cv_free.notify_all();
        }
        if (th.joinable()) th.join();
        if (fp) fclose(fp);
    }
};

struct Line { const uint8_t *p; size_t len; bool term; };   // term: the line had a terminator

}  // namespace

struct mcxio_file {
    Source *src = nullptr;            // null for in-memory input
    // window over the decompressed stream
    std::vector<uint8_t> win;         // file input: unconsumed tail + appended chunks
    const uint8_t *base = nullptr;    // start of the window storage
    size_t pos = 0, end = 0;
    bool src_eof = false;
    // parser state (readfq's `last`)
    bool have_last = false;
    uint8_t last0 = 0;
    size_t last_pos = 0;              // where the line held in `last` started (in-memory parsing: lets a piece stop before it)
    size_t stop_pos = (size_t)-1;     // in-memory parsing of a piece: no record is started at or behind this position
    bool hit_end = false;             // ... and the piece ran out of bytes before reaching it
    bool last_without_quality = false; // a FASTQ record cut short by the end of the file came back without qualities
    // plain (uncompressed) files can also be read piecewise by several threads (mcxio_next_packed)
    int plain_fd = -1;
    const uint8_t *plain_map = nullptr;
    struct PiecePool *pool = nullptr; // parser states of the pieces, kept between windows (their vectors keep their capacity)
    size_t plain_size = 0, plain_pos = 0;
    bool plain_fastq = false;
    int threads_hint = 1;
    bool plain_mode = false;          // the file is being read through mcxio_next_packed's piecewise path
    int64_t reparsed = 0;             // windows in which a guessed record start was wrong and one thread parsed again
    bool eof = false;
    int64_t records_total = 0, bases_total = 0;
    // batch storage
    std::vector<uint8_t> bases, quals;
    std::vector<int64_t> offs;
    bool any_qual = false;
    std::string err;

    bool refill() {                   // append the next chunk; false at the end of the stream
        if (!src) { src_eof = true; return false; }
        Chunk c;
        if (!src->next(c)) { src_eof = true; return false; }
        if (pos > 0) {
            memmove(win.data(), win.data() + pos, end - pos);
            end -= pos; pos = 0;
        }
        if (win.size() < end + c.n) win.resize(end + c.n + CHUNK);
        memcpy(win.data() + end, c.data.data(), c.n);
        end += c.n;
        base = win.data();
        src->recycle(std::move(c));
        return true;
    }

    // next line in Python's universal-newline text mode: '\n', '\r\n' and a lone '\r' all end a line
    bool get_line(Line &l) {
        for (;;) {
            const uint8_t *s = base + pos, *e = base + end;
            const uint8_t *q = s < e ? (const uint8_t *)memchr(s, '\n', (size_t)(e - s)) : nullptr;
            const uint8_t *seg_end = q ? q : e;
            const uint8_t *r = s < seg_end ? (const uint8_t *)memchr(s, '\r', (size_t)(seg_end - s)) : nullptr;
            if (r) {
                if (r + 1 < e || src_eof) {
                    l.p = s; l.len = (size_t)(r - s); l.term = true;
                    pos = (size_t)(r - base) + ((r + 1 < e && r[1] == '\n') ? 2 : 1);
                    return true;
                }
            } else if (q) {
                l.p = s; l.len = (size_t)(q - s); l.term = true;
                pos = (size_t)(q - base) + 1;
                return true;
            } else if (src_eof) {
                if (s == e) return false;
                l.p = s; l.len = (size_t)(e - s); l.term = false;
                pos = end;
                return true;
            }
            if (!refill() && src && src->err) return false;
        }
    }
};

static int skip_plain_rest(mcxio_file *f);
static void free_pool(mcxio_file *f);

namespace {

thread_local std::string g_err;

// l[:-1] of the reference: a line that ended without a terminator loses its last character
inline size_t stripped(const Line &l) { return l.term ? l.len : (l.len ? l.len - 1 : 0); }

// One pass of the readfq state machine (mc.py:294-325); store = false only counts.
int parse(mcxio_file *f, int64_t max_records, bool store, int64_t *n_out) {
    int64_t n = 0;
    Line l;
    while (!f->eof && (max_records < 0 || n < max_records)) {
        if (!f->have_last) {                                  // search for the start of the next record
            for (;;) {
                const size_t at = f->pos;
                if (!f->get_line(l)) break;
                const uint8_t c = l.len ? l.p[0] : (uint8_t)'\n';
                if (c == '>' || c == '@') { f->have_last = stripped(l) > 0; f->last0 = c; f->last_pos = at; break; }
            }
        }
        if (!f->have_last) { f->eof = true; break; }
        if (f->last_pos >= f->stop_pos) {                     // (pieces of the parallel reader) the next piece starts here
            f->pos = f->last_pos; f->have_last = false;
            break;
        }
        f->have_last = false;
        const size_t seq_start = f->bases.size();
        size_t seqlen = 0;
        for (;;) {                                            // read the sequence
            const size_t at = f->pos;
            if (!f->get_line(l)) break;
            const uint8_t c = l.len ? l.p[0] : (uint8_t)'\n';
            if (c == '@' || c == '+' || c == '>') { f->have_last = stripped(l) > 0; f->last0 = c; f->last_pos = at; break; }
            const size_t k = stripped(l);
            if (store && k) f->bases.insert(f->bases.end(), l.p, l.p + k);
            seqlen += k;
        }
        bool with_qual = false;
        if (f->have_last && f->last0 == '+') {                // a FASTQ record: read the quality
            size_t leng = 0;
            bool enough = false;
            const size_t qstart = f->quals.size();
            if (store && f->quals.size() < seq_start) f->quals.resize(seq_start, (uint8_t)'~');
            while (f->get_line(l)) {
                const size_t k = stripped(l);
                if (store && k && leng < seqlen) {            // only the first len(seq) characters are ever used
                    const size_t take = k < seqlen - leng ? k : seqlen - leng;
                    f->quals.insert(f->quals.end(), l.p, l.p + take);
                }
                leng += k;
                if (leng >= seqlen) { f->have_last = false; enough = true; break; }
            }
            if (enough) { with_qual = true; f->any_qual |= store; }
            else {                                            // end of file before enough quality: yielded as a
                if (store) f->quals.resize(qstart);           // FASTA record, and the generator stops
                f->eof = true;
                f->last_without_quality = true;
            }
        } else if (!f->have_last) {
            f->eof = true;                                    // "if not last: break" after the yield
        }
        if (f->src && f->src->err) break;
        if (store) {
            if (!with_qual && f->any_qual) f->quals.resize(seq_start + seqlen, (uint8_t)'~');
            f->offs.push_back((int64_t)(seq_start + seqlen));
        }
        ++n;
        ++f->records_total;
        f->bases_total += (int64_t)seqlen;
    }
    if (f->src && f->src->err) { f->err = f->src->msg; return f->src->err; }
    *n_out = n;
    return MCXIO_OK;
}

}  // namespace

extern "C" int mcxio_open(mcxio_file **out, const char *path) {
    if (!out || !path) { g_err = "mcxio_open: null argument"; return MCXIO_EINVAL; }
    FILE *fp = fopen(path, "rb");
    if (!fp) { g_err = std::string("mcxio_open: cannot open ") + path; return MCXIO_EIO; }
    unsigned char magic[3] = {0, 0, 0};
    const size_t got = fread(magic, 1, 3, fp);
    rewind(fp);
    if (got >= 3 && magic[0] == 'B' && magic[1] == 'Z' && magic[2] == 'h') {
        fclose(fp);
        g_err = "mcxio_open: bzip2 input must be decompressed by the caller (mcxio_open_mem)";
        return MCXIO_EFORMAT;
    }
    mcxio_file *f = new mcxio_file();
    f->src = new Source();
    f->src->fp = fp;
    f->src->gz = got >= 2 && magic[0] == 0x1f && magic[1] == 0x8b;
    f->offs.push_back(0);
    {   // size the batch buffers once instead of doubling them while parsing (plain text: at most the file size;
        // gzip: about four times the compressed size), capped so that a huge file does not reserve it all
        fseek(fp, 0, SEEK_END);
        const long fsz = ftell(fp);
        rewind(fp);
        if (fsz > 0) {
            const size_t want = std::min<size_t>((size_t)fsz * (f->src->gz ? 4 : 1), (size_t)2 << 30);
            try { f->bases.reserve(want / 2 + 64); f->quals.reserve(want / 2 + 64); f->offs.reserve(want / 200 + 16); } catch (const std::bad_alloc &) {}
        }
    }
    if (!f->src->gz) {
        f->plain_fd = open(path, O_RDONLY);
        struct stat sb;
        if (f->plain_fd >= 0 && fstat(f->plain_fd, &sb) == 0) f->plain_size = (size_t)sb.st_size;
        else if (f->plain_fd >= 0) { close(f->plain_fd); f->plain_fd = -1; }
        if (f->plain_fd >= 0 && f->plain_size > 0) {          // pieces are parsed straight from the page cache
            void *m = mmap(nullptr, f->plain_size, PROT_READ, MAP_PRIVATE, f->plain_fd, 0);
            if (m == MAP_FAILED) { close(f->plain_fd); f->plain_fd = -1; }
            else { f->plain_map = (const uint8_t *)m; madvise(m, f->plain_size, MADV_SEQUENTIAL); }
        } else if (f->plain_fd >= 0) { close(f->plain_fd); f->plain_fd = -1; }
        f->plain_fastq = got >= 1 && magic[0] == '@';
    }
    f->src->start();
    *out = f;
    return MCXIO_OK;
}

extern "C" int mcxio_open_mem(mcxio_file **out, const uint8_t *data, int64_t n) {
    if (!out || n < 0 || (n > 0 && !data)) { g_err = "mcxio_open_mem: bad argument"; return MCXIO_EINVAL; }
    mcxio_file *f = new mcxio_file();
    f->base = data; f->pos = 0; f->end = (size_t)n; f->src_eof = true;
    f->offs.push_back(0);
    *out = f;
    return MCXIO_OK;
}

extern "C" int mcxio_next_batch(mcxio_file *f, int64_t max_records, mcxio_batch *out) {
    if (!f || !out) { g_err = "mcxio_next_batch: null argument"; return MCXIO_EINVAL; }
    f->bases.clear(); f->quals.clear(); f->offs.clear(); f->offs.push_back(0);
    f->any_qual = false;
    int64_t n = 0;
    int rc;
    try {
        rc = parse(f, max_records, true, &n);
    } catch (const std::bad_alloc &) {
        f->err = "out of memory";
        return MCXIO_ENOMEM;
    }
    if (rc != MCXIO_OK) return rc;
    if (f->any_qual && f->quals.size() < f->bases.size()) f->quals.resize(f->bases.size(), (uint8_t)'~');
    if (f->any_qual && f->quals.capacity() == 0) f->quals.reserve(16);   // non-null even when every sequence is empty
    out->bases = f->bases.data();
    out->quals = f->any_qual ? f->quals.data() : nullptr;
    out->offsets = f->offs.data();
    out->n = n;
    out->records_total = f->records_total;
    out->bases_total = f->bases_total;
    out->eof = f->eof ? 1 : 0;
    return MCXIO_OK;
}

extern "C" int mcxio_skip_rest(mcxio_file *f, int64_t *records, int64_t *bases) {
    if (!f) { g_err = "mcxio_skip_rest: null argument"; return MCXIO_EINVAL; }
    int64_t n = 0;
    if (f->plain_mode) {                       // the file is read piecewise: count the rest the same way
        const int rcp = skip_plain_rest(f);
        if (rcp != MCXIO_OK) return rcp;
        if (records) *records = f->records_total;
        if (bases) *bases = f->bases_total;
        return MCXIO_OK;
    }
    const int rc = parse(f, -1, false, &n);
    if (rc != MCXIO_OK) return rc;
    if (records) *records = f->records_total;
    if (bases) *bases = f->bases_total;
    return MCXIO_OK;
}

extern "C" void mcxio_close(mcxio_file *f) {
    if (!f) return;
    if (f->plain_map) munmap((void *)f->plain_map, f->plain_size);
    if (f->plain_fd >= 0) close(f->plain_fd);
    free_pool(f);
    delete f->src;
    delete f;
}

extern "C" const char *mcxio_last_error(mcxio_file *f) { return f ? f->err.c_str() : g_err.c_str(); }

// ------------------------------------------------------------------------------------------------
// Packed batches, parsed and packed by several threads ("multithreaded C++ parser/packer").
//
// Output = the layout the reads have in HBM (include/mcx.h, mcx_push_reads_packed): per read 3 ceil(len / 32) words of
// bit-planes (lo, hi, mask), the lengths, and the quality bytes; written into buffers from the caller's allocator
// (page-locked memory from mcx_host_alloc, so that the push is an asynchronous DMA).
//
// Plain files: the window of text that holds the next ~target_records records is cut into one piece per thread.  Each
// thread guesses where the first record of its piece starts (a line that begins a record: '>' for FASTA; '@' followed
// two lines later by a '+' line for FASTQ), runs the SAME readfq state machine as the sequential reader from there, and
// stops at the first record that starts at or behind the end of its piece.  The guesses are then checked: piece k + 1
// must begin exactly where the parser of piece k stopped.  If a guess was wrong (multi-line FASTQ with '@' quality
// lines, a '+' inside FASTA, ...) everything from that piece on is parsed again by one thread from the true position,
// so the records are always those of the sequential state machine (tests/test_seqio.py compares the two paths).
// gzip / bzip2 / in-memory input: one inflating producer, the sequential parser, then the packing in parallel.
// ------------------------------------------------------------------------------------------------

namespace {

typedef int (*alloc_fn)(void **, size_t);
typedef void (*free_fn)(void *);
alloc_fn g_alloc = nullptr;
free_fn g_free = nullptr;

// Output buffers are recycled: page-locking a gigabyte costs more than parsing it, so buffers handed back through
// mcxio_free_packed wait in a small cache for the next batch of similar size.
struct CachedBuf { void *p; size_t cap; };
std::mutex g_cache_mu;
std::vector<CachedBuf> g_cache;            // free buffers
std::vector<CachedBuf> g_live;             // buffers handed out (to know their capacity when they come back)
void *out_alloc(size_t bytes) {
    bytes = bytes ? bytes : 1;
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        size_t best = g_cache.size();
        for (size_t k = 0; k < g_cache.size(); ++k)
            if (g_cache[k].cap >= bytes && g_cache[k].cap <= 2 * bytes + (1u << 20) && (best == g_cache.size() || g_cache[k].cap < g_cache[best].cap)) best = k;
        if (best < g_cache.size()) {
            CachedBuf b = g_cache[best];
            g_cache.erase(g_cache.begin() + (long)best);
            g_live.push_back(b);
            return b.p;
        }
    }
    const size_t cap = bytes + bytes / 8 + 4096;
    void *p = nullptr;
    if (g_alloc) { if (g_alloc(&p, cap) != 0) return nullptr; }
    else p = malloc(cap);
    if (p) { std::lock_guard<std::mutex> g(g_cache_mu); g_live.push_back(CachedBuf{p, cap}); }
    return p;
}
void raw_free(void *p) { if (g_free) g_free(p); else free(p); }
void out_free(void *p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(g_cache_mu);
    size_t cap = 0;
    for (size_t k = 0; k < g_live.size(); ++k) if (g_live[k].p == p) { cap = g_live[k].cap; g_live.erase(g_live.begin() + (long)k); break; }
    if (cap == 0) { raw_free(p); return; }
    if (g_cache.size() >= 9) { raw_free(g_cache.front().p); g_cache.erase(g_cache.begin()); }
    g_cache.push_back(CachedBuf{p, cap});
}

// 32 bases -> (lo, hi, mask): T C A G = 0..3; mask = not an upper-case ACGT, under it lo = (c != 'N')
inline void pack32(const uint8_t *s, int n, uint32_t &lo, uint32_t &hi, uint32_t &mk) {
    lo = hi = mk = 0;
    int k = 0;
#if defined(__SSE2__)
    const __m128i cA = _mm_set1_epi8('A'), cC = _mm_set1_epi8('C'), cG = _mm_set1_epi8('G'), cT = _mm_set1_epi8('T'), cN = _mm_set1_epi8('N');
    for (; k + 16 <= n; k += 16) {
        const __m128i v = _mm_loadu_si128(reinterpret_cast<const __m128i *>(s + k));
        const uint32_t a = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, cA)), c = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, cC)),
                       g = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, cG)), t = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, cT)),
                       nn = (uint32_t)_mm_movemask_epi8(_mm_cmpeq_epi8(v, cN));
        const uint32_t bad = ~(a | c | g | t) & 0xffffu;
        lo |= ((c | g) | (bad & ~nn)) << k;
        hi |= (a | g) << k;
        mk |= bad << k;
    }
#endif
    for (; k < n; ++k) {
        const uint8_t ch = s[k];
        uint32_t l = 0, h = 0, m = 0;
        switch (ch) {
            case 'T': break;
            case 'C': l = 1; break;
            case 'A': h = 1; break;
            case 'G': l = 1; h = 1; break;
            case 'N': m = 1; break;
            default: m = 1; l = 1; break;
        }
        lo |= l << k; hi |= h << k; mk |= m << k;
    }
}

void pack_reads(const uint8_t *bases, const int64_t *offs, int64_t r0, int64_t r1, uint32_t *out /* at the record of r0 */, uint32_t *lengths) {
    for (int64_t r = r0; r < r1; ++r) {
        const uint8_t *s = bases + offs[r];
        const int len = (int)(offs[r + 1] - offs[r]);
        const int G = (len + 31) >> 5;
        lengths[r] = (uint32_t)len;
        for (int g = 0; g < G; ++g) pack32(s + 32 * g, std::min(32, len - 32 * g), out[g], out[G + g], out[2 * G + g]);
        out += 3 * G;
    }
}

// where does a record begin at or behind `from`?  (a guess, checked afterwards)
size_t guess_record_start(const uint8_t *b, size_t from, size_t end, bool fastq) {
    size_t p = from;
    if (p > 0) {                                    // move to the start of the next line
        while (p < end && b[p - 1] != '\n' && b[p - 1] != '\r') ++p;
        if (p < end && b[p - 1] == '\r' && b[p] == '\n') ++p;
    }
    auto next_line = [&](size_t q) {                // start of the line after the one starting at q
        while (q < end && b[q] != '\n' && b[q] != '\r') ++q;
        if (q < end && b[q] == '\r' && q + 1 < end && b[q + 1] == '\n') ++q;
        return q < end ? q + 1 : end;
    };
    for (int tries = 0; p < end && tries < 1000000; ++tries) {
        if (!fastq) { if (b[p] == '>') return p; }
        else if (b[p] == '@') {
            const size_t l2 = next_line(p), l3 = next_line(l2);
            if (l3 < end && b[l3] == '+' && l2 < end && b[l2] != '@' && b[l2] != '+' && b[l2] != '>') return p;
        }
        p = next_line(p);
    }
    return end;
}

struct Piece {
    mcxio_file st;              // parser state + the records of the piece
    size_t begin = 0, guess = 0, stop = 0, stopped_at = 0;
    bool ran_out = false;
    int64_t n = 0;
};

}  // namespace

// batch in the device layout; buffers come from the allocator set with mcxio_set_allocator and belong to the CALLER
// after the call (free them with the matching free function)
extern "C" int mcxio_set_allocator(int (*alloc)(void **, size_t), void (*free_)(void *)) {
    {   // buffers of the old allocator must go back to it
        std::lock_guard<std::mutex> g(g_cache_mu);
        for (CachedBuf &b : g_cache) raw_free(b.p);
        g_cache.clear();
    }
    g_alloc = alloc; g_free = free_;
    return MCXIO_OK;
}
extern "C" void mcxio_free_packed(mcxio_packed *b) {
    if (!b) return;
    out_free(b->packed); out_free(b->lengths); out_free(b->quals);
    b->packed = nullptr; b->lengths = nullptr; b->quals = nullptr;
}

namespace {

// records of `parts` (in order) -> one packed batch; the packing runs on `threads` threads
int emit_packed(std::vector<mcxio_file *> &parts, int threads, mcxio_packed *out) {
    int64_t n = 0, nb = 0, nw = 0;
    bool any_qual = false;
    std::vector<int64_t> rec0(parts.size() + 1, 0), word0(parts.size() + 1, 0), base0(parts.size() + 1, 0);
    for (size_t k = 0; k < parts.size(); ++k) {
        mcxio_file *f = parts[k];
        const int64_t cnt = (int64_t)f->offs.size() - 1;
        int64_t w = 0;
        for (int64_t r = 0; r < cnt; ++r) w += 3 * ((f->offs[r + 1] - f->offs[r] + 31) >> 5);
        rec0[k + 1] = rec0[k] + cnt; word0[k + 1] = word0[k] + w; base0[k + 1] = base0[k] + f->offs[cnt];
        any_qual |= f->any_qual;
    }
    n = rec0.back(); nw = word0.back(); nb = base0.back();
    out->n = n; out->n_words = nw; out->n_bases = nb;
    out->packed = (uint32_t *)out_alloc((size_t)nw * 4 + 16);
    out->lengths = (uint32_t *)out_alloc((size_t)n * 4 + 16);
    out->quals = any_qual ? (uint8_t *)out_alloc((size_t)nb + 32) : nullptr;
    if (!out->packed || !out->lengths || (any_qual && !out->quals)) { mcxio_free_packed(out); return MCXIO_ENOMEM; }
    // work units: slices of the parts of about equal size
    struct Unit { size_t part; int64_t r0, r1, w0; };
    std::vector<Unit> units;
    const int64_t per = std::max<int64_t>(4096, n / std::max(1, threads * 4));
    for (size_t k = 0; k < parts.size(); ++k) {
        mcxio_file *f = parts[k];
        const int64_t cnt = (int64_t)f->offs.size() - 1;
        int64_t w = word0[k];
        for (int64_t r = 0; r < cnt; r += per) {
            const int64_t r1 = std::min(cnt, r + per);
            units.push_back(Unit{k, r, r1, w});
            for (int64_t q = r; q < r1; ++q) w += 3 * ((f->offs[q + 1] - f->offs[q] + 31) >> 5);
        }
    }
    std::atomic<size_t> next{0};
    auto work = [&]() {
        for (;;) {
            const size_t u = next.fetch_add(1);
            if (u >= units.size()) return;
            const Unit &U = units[u];
            mcxio_file *f = parts[U.part];
            pack_reads(f->bases.data(), f->offs.data(), U.r0, U.r1, out->packed + U.w0, out->lengths + rec0[U.part]);
            if (any_qual) {
                const int64_t b0 = f->offs[U.r0], b1 = f->offs[U.r1];
                uint8_t *dst = out->quals + base0[U.part];
                if (f->any_qual) {
                    const size_t have = f->quals.size();
                    for (int64_t b = b0; b < b1; ++b) dst[b] = (size_t)b < have ? f->quals[(size_t)b] : (uint8_t)'~';
                } else memset(dst + b0, '~', (size_t)(b1 - b0));
            }
        }
    };
    const int nt = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, units.size()));
    std::vector<std::thread> th;
    for (int t = 1; t < nt; ++t) th.emplace_back(work);
    work();
    for (auto &t : th) t.join();
    return MCXIO_OK;
}

}  // namespace

namespace {

// One window of a plain file parsed by `threads` threads (see the block comment above).  On return pc[0 .. good) and, if
// have_tail, `tail` hold the records (store) or just their counts, in file order; f's position, totals and eof are updated.
struct Window { std::vector<Piece *> pc; Piece *tail = nullptr; bool have_tail = false; int good = 0; };

}  // namespace

struct PiecePool { std::vector<Piece *> all; ~PiecePool() { for (Piece *p : all) delete p; } };
static void free_pool(mcxio_file *f) { delete f->pool; f->pool = nullptr; }

namespace {

void reset_piece(Piece &P, const uint8_t *base, size_t pos, size_t end, size_t stop_pos) {
    mcxio_file &S = P.st;
    S.base = base; S.pos = pos; S.end = end; S.src_eof = true; S.stop_pos = stop_pos;
    S.have_last = false; S.eof = false; S.any_qual = false; S.last_without_quality = false;
    S.records_total = 0; S.bases_total = 0;
    S.bases.clear(); S.quals.clear(); S.offs.clear(); S.offs.push_back(0);
    P.n = 0; P.ran_out = false;
}

int plain_window(mcxio_file *f, int64_t target_records, int threads, bool store, Window &W) {
    const size_t fsize = f->plain_size;
    const uint8_t *M = f->plain_map;
    if (target_records < 0) target_records = INT64_MAX / 4096;
    const double per_rec = f->records_total > 0 ? (double)f->plain_pos / (double)f->records_total : 400.0;
    size_t want = (size_t)std::min<double>((double)(fsize - f->plain_pos), std::max(1.0, (double)target_records * per_rec * 1.02 + 65536.0));
    if (f->records_total == 0) want = std::min<size_t>(want, (size_t)64 << 20);     // first window: learn the record size
    if (const char *e = getenv("MCXIO_WINDOW_BYTES")) want = std::min<size_t>(fsize - f->plain_pos, (size_t)std::max(1, atoi(e)));
    const size_t w_begin = f->plain_pos, w_end = std::min(fsize, w_begin + want);
    const bool to_eof = w_end == fsize;
    size_t margin = (size_t)4 << 20, piece_min = (size_t)1 << 20;       // (the environment overrides exist for the tests)
    if (const char *e = getenv("MCXIO_MARGIN_BYTES")) margin = (size_t)std::max(1, atoi(e));
    if (const char *e = getenv("MCXIO_PIECE_BYTES")) piece_min = (size_t)std::max(1, atoi(e));
    const int np = (int)std::max<size_t>(1, std::min<size_t>((size_t)threads, (w_end - w_begin) / piece_min + 1));
    if (!f->pool) f->pool = new PiecePool();
    while ((int)f->pool->all.size() < np + 1) f->pool->all.push_back(new Piece());
    W.pc.assign(f->pool->all.begin(), f->pool->all.begin() + np);
    W.tail = f->pool->all[(size_t)np];
    auto run_piece = [&](int k) {
        Piece &P = *W.pc[(size_t)k];
        P.begin = w_begin + (w_end - w_begin) * (size_t)k / (size_t)np;
        P.stop = k + 1 < np ? w_begin + (w_end - w_begin) * (size_t)(k + 1) / (size_t)np : w_end;
        const bool final_piece = to_eof && k + 1 == np;
        const size_t rd_end = final_piece ? fsize : std::min(fsize, P.stop + margin);
        // absolute file offsets are positions in the mapping; a piece sees [0, rd_end)
        const size_t g = k == 0 ? P.begin : guess_record_start(M, P.begin, rd_end, f->plain_fastq);
        P.guess = g;
        reset_piece(P, M, g, rd_end, final_piece ? (size_t)-1 : P.stop);
        if (store) { P.st.bases.reserve((rd_end - P.begin) / 2 + 64); if (f->plain_fastq) P.st.quals.reserve((rd_end - P.begin) / 2 + 64); }
        int64_t n = 0;
        parse(&P.st, -1, store, &n);
        P.n = n;
        P.stopped_at = P.st.pos;
        P.ran_out = P.st.eof && !final_piece;
    };
    {
        std::vector<std::thread> th;
        for (int k = 1; k < np; ++k) th.emplace_back(run_piece, k);
        run_piece(0);
        for (auto &t : th) t.join();
    }
    // check the guesses; from the first wrong one on, one thread parses the rest of the window again
    int good = 1;
    bool redo = W.pc[0]->ran_out;
    for (int k = 1; k < np && !redo; ++k) {
        if (W.pc[(size_t)k - 1]->stopped_at == W.pc[(size_t)k]->guess && !W.pc[(size_t)k]->ran_out) ++good;
        else redo = true;
    }
    W.have_tail = false;
    if (redo) {
        ++f->reparsed;
        if (W.pc[0]->ran_out) good = 0;
        const size_t from = good == 0 ? w_begin : W.pc[(size_t)good - 1]->stopped_at;
        size_t reach = (size_t)64 << 20;
        for (;;) {                                   // widen until the record that crosses the end of the window fits
            const size_t rd_end = std::min(fsize, std::max(w_end, from) + reach);
            const bool final_piece = rd_end == fsize && to_eof;
            reset_piece(*W.tail, M, from, rd_end, final_piece ? (size_t)-1 : std::max(w_end, from));
            int64_t n = 0;
            parse(&W.tail->st, -1, store, &n);
            W.tail->n = n; W.tail->stopped_at = W.tail->st.pos;
            if (W.tail->st.eof && !final_piece && rd_end < fsize) { reach *= 4; continue; }
            break;
        }
        W.have_tail = true;
    }
    W.good = good;
    Piece &lastp = W.have_tail ? *W.tail : *W.pc[(size_t)good - 1];
    f->plain_pos = lastp.stopped_at;
    f->eof = (lastp.st.eof && (W.have_tail || to_eof)) || f->plain_pos >= fsize;
    for (int k = 0; k < good; ++k) { f->records_total += W.pc[(size_t)k]->st.records_total; f->bases_total += W.pc[(size_t)k]->st.bases_total; f->last_without_quality |= W.pc[(size_t)k]->st.last_without_quality; }
    if (W.have_tail) { f->records_total += W.tail->st.records_total; f->bases_total += W.tail->st.bases_total; f->last_without_quality |= W.tail->st.last_without_quality; }
    f->plain_mode = true;
    return MCXIO_OK;
}

}  // namespace

static int skip_plain_rest(mcxio_file *f) {
    try {
        while (!f->eof) { Window W; plain_window(f, 4000000, f->threads_hint, false, W); }
    } catch (const std::bad_alloc &) { f->err = "out of memory"; return MCXIO_ENOMEM; }
    return MCXIO_OK;
}

extern "C" int mcxio_next_packed(mcxio_file *f, int64_t target_records, int threads, mcxio_packed *out) {
    if (!f || !out) { g_err = "mcxio_next_packed: null argument"; return MCXIO_EINVAL; }
    memset(out, 0, sizeof *out);
    threads = std::max(1, std::min(threads, 256));
    f->threads_hint = threads;
    try {
        if (!f->plain_map || (threads == 1 && !f->plain_mode)) {
            // sequential parse (inflating producer ahead of it), parallel packing
            f->bases.clear(); f->quals.clear(); f->offs.clear(); f->offs.push_back(0);
            f->any_qual = false;
            int64_t n = 0;
            const int rc = parse(f, target_records, true, &n);
            if (rc != MCXIO_OK) return rc;
            std::vector<mcxio_file *> parts{f};
            const int rc2 = emit_packed(parts, threads, out);
            if (rc2 != MCXIO_OK) { f->err = "out of memory"; return rc2; }
        } else {
            Window W;
            if (!f->eof) {
                plain_window(f, target_records, threads, true, W);
                std::vector<mcxio_file *> parts;
                for (int k = 0; k < W.good; ++k) parts.push_back(&W.pc[(size_t)k]->st);
                if (W.have_tail) parts.push_back(&W.tail->st);
                const int rc2 = emit_packed(parts, threads, out);
                if (rc2 != MCXIO_OK) { f->err = "out of memory"; return rc2; }
            }
        }
    } catch (const std::bad_alloc &) {
        f->err = "out of memory";
        mcxio_free_packed(out);
        return MCXIO_ENOMEM;
    }
    out->records_total = f->records_total;
    out->bases_total = f->bases_total;
    out->eof = f->eof ? 1 : 0;
    out->last_without_quality = f->last_without_quality ? 1 : 0;
    out->reparsed = f->reparsed;
    return MCXIO_OK;
}

// advance the reader by about target_records records without storing them (a rank of a sharded run that does not own the
// next batch); cut exactly like mcxio_next_packed with the same arguments would cut, so all ranks see the same batches
extern "C" int mcxio_skip_packed(mcxio_file *f, int64_t target_records, int threads, int64_t *n_skipped) {
    if (!f) { g_err = "mcxio_skip_packed: null argument"; return MCXIO_EINVAL; }
    threads = std::max(1, std::min(threads, 256));
    f->threads_hint = threads;
    const int64_t before = f->records_total;
    try {
        if (!f->plain_map || (threads == 1 && !f->plain_mode)) {
            int64_t n = 0;
            const int rc = parse(f, target_records, false, &n);
            if (rc != MCXIO_OK) return rc;
        } else if (!f->eof) {
            Window W;
            plain_window(f, target_records, threads, false, W);
        }
    } catch (const std::bad_alloc &) { f->err = "out of memory"; return MCXIO_ENOMEM; }
    if (n_skipped) *n_skipped = f->records_total - before;
    return MCXIO_OK;
}

extern "C" int mcxio_state(mcxio_file *f, int64_t *records_total, int64_t *bases_total, int32_t *eof) {
    if (!f) { g_err = "mcxio_state: null argument"; return MCXIO_EINVAL; }
    if (records_total) *records_total = f->records_total;
    if (bases_total) *bases_total = f->bases_total;
    if (eof) *eof = f->eof ? 1 : 0;
    return MCXIO_OK;
}

// mcxio.cpp -- libmcxio.so: streaming FASTA/FASTQ reader for the MicrobeCensus hot path (host side).
//
// Restates parse_seqs() of the reference (mc.py:294-325, the readfq generator) as a line-driven state machine over
// a byte stream, and open_file() (mc.py:47-59) as a producer thread that read()s or inflates (zlib, multi-member
// gzip) 8 MB chunks ahead of the parser.  Output is the packed layout mcx_push_reads takes.  See include/mcxio.h.
#include "../../include/mcxio.h"

#include <zlib.h>

#include <algorithm>
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
            while (f->get_line(l)) {
                const uint8_t c = l.len ? l.p[0] : (uint8_t)'\n';
                if (c == '>' || c == '@') { f->have_last = stripped(l) > 0; f->last0 = c; break; }
            }
        }
        if (!f->have_last) { f->eof = true; break; }
        f->have_last = false;
        const size_t seq_start = f->bases.size();
        size_t seqlen = 0;
        while (f->get_line(l)) {                              // read the sequence
            const uint8_t c = l.len ? l.p[0] : (uint8_t)'\n';
            if (c == '@' || c == '+' || c == '>') { f->have_last = stripped(l) > 0; f->last0 = c; break; }
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
    const int rc = parse(f, -1, false, &n);
    if (rc != MCXIO_OK) return rc;
    if (records) *records = f->records_total;
    if (bases) *bases = f->bases_total;
    return MCXIO_OK;
}

extern "C" void mcxio_close(mcxio_file *f) {
    if (!f) return;
    delete f->src;
    delete f;
}

extern "C" const char *mcxio_last_error(mcxio_file *f) { return f ? f->err.c_str() : g_err.c_str(); }

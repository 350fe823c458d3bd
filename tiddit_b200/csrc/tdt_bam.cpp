// libtdt_bam.so -- BGZF/BAM scanner producing columns for the GPU coverage / signal path (include/tdt_bam.h).
// Host code only (g++ + zlib).  The file is mapped, BGZF block boundaries are found by hopping over the BSIZE
// fields, a window of blocks is inflated on a pool of threads straight into one contiguous buffer (every block
// states its inflated size in its trailer, so the destinations are known up front), and records are decoded
// from that buffer without copying (record boundaries by a sequential walk over the block_size words, the fields of
// the records on a second small pool).  Blocks go through the whole-buffer decoder of tdt_inflate.h (r02 v6); zlib
// computes the CRC of every block and inflates the ones that decoder refuses or gets wrong (TDT_BAM_ZLIB=1: all, with zlib's CRC).
#include "../../include/tdt_bam.h"
#include "tdt_inflate.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <memory>
#include <mutex>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <exception>
#include <new>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

inline double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
inline uint32_t le32(const uint8_t *p) { uint32_t v; memcpy(&v, p, 4); return v; }
inline uint16_t le16(const uint8_t *p) { uint16_t v; memcpy(&v, p, 2); return v; }

// does this CIGAR operation consume the reference?  M I D N S H P = X
const uint8_t kConsumesRef[16] = {1, 0, 1, 1, 0, 0, 0, 1, 1, 0, 0, 0, 0, 0, 0, 0};

struct Block {
    const uint8_t *payload;  // raw deflate stream
    uint32_t clen;           // its length
    uint32_t isize;          // inflated size
    uint32_t crc;
    size_t dst;              // offset in the window buffer
};

// A fixed set of threads that run one job at a time: run(f) calls f(worker) on every worker (the caller is worker 0) and
// returns when all are done.  The reader owns two: the background thread inflates the next window on one while the
// caller's thread decodes the records of the current window on the other.
class Pool {
    std::vector<std::thread> th_;
    std::mutex m_;
    std::condition_variable start_, done_;
    std::function<void(int)> job_;
    int generation_ = 0, pending_ = 0;
    bool stop_ = false;

    void worker(int idx) {
        int seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(m_);
            start_.wait(lk, [&] { return stop_ || generation_ != seen; });
            if (stop_) return;
            seen = generation_;
            std::function<void(int)> f = job_;
            lk.unlock();
            f(idx);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }

public:
    explicit Pool(int n) {
        for (int i = 1; i < n; ++i) th_.emplace_back([this, i] { worker(i); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> lk(m_);
            stop_ = true;
        }
        start_.notify_all();
        for (auto &t : th_) t.join();
    }
    int size() const { return (int)th_.size() + 1; }
    void run(const std::function<void(int)> &f) {   // f must not throw
        {
            std::lock_guard<std::mutex> lk(m_);
            job_ = f;
            pending_ = (int)th_.size();
            ++generation_;
        }
        start_.notify_all();
        f(0);
        std::unique_lock<std::mutex> lk(m_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }
};

struct InflateWorker {       // per inflate-pool worker, kept between windows
    z_stream zs;
    bool zs_ready = false;
    tdtz::Inflater fast;
    ~InflateWorker() {
        if (zs_ready) inflateEnd(&zs);
    }
};

}  // namespace

struct tdt_bam_reader {
    int fd = -1;
    const uint8_t *map = nullptr;
    size_t map_len = 0;
    size_t cpos = 0;  // next compressed block
    int threads = 1;
    std::vector<uint8_t> buf;  // inflated window being parsed
    size_t bpos = 0, bend = 0;
    size_t window_blocks = 256;  // blocks inflated per window (<= 16 MiB inflated: stays in the last-level cache until it is parsed, and
                                 // only the first window is inflated with nobody parsing; 1 M-read file, 8 cores: 1024 blocks 0.27-0.30 s,
                                 // 512: 0.22, 256: 0.18-0.23, 128: 0.19-0.24, 64: 0.26 -- the pool is started per window)
    bool eof = false;
    // the NEXT window is inflated in the background while the caller works on the current one
    std::vector<uint8_t> nbuf;    // next window, data at [kPad, kPad + nbytes)
    std::thread bg;
    bool bg_running = false;
    int bg_rc = 0;                // 1 = bytes ready, 0 = end of file, < 0 error
    size_t bg_bytes = 0;
    char bg_err[512] = "";
    std::unique_ptr<Pool> inflate_pool, parse_pool;
    std::vector<std::unique_ptr<InflateWorker>> inflate_workers;
    std::vector<int64_t> rec_at;  // offsets of the records of the batch being decoded
    double t_wait = 0, t_walk = 0, t_decode = 0;   // TDT_BAM_TIMING=1: seconds waiting for the inflater / walking / decoding
    std::string text;
    std::vector<std::string> ref_names;
    std::vector<int32_t> ref_lens;
    std::string path;
};

namespace {

constexpr size_t kPad = 1 << 20;  // headroom in front of a window for the unread tail of the previous one

// Inflates the next up-to-window_blocks blocks (from r->cpos, which it advances) into dst at offset kPad.
// Runs on the background thread: touches only cpos, dst and the bg_* fields.  Returns 1 / 0 (end of file) / < 0.
int inflate_window(tdt_bam_reader *r, std::vector<uint8_t> &dst, size_t *nbytes, char *err, size_t err_len) {
    std::vector<Block> blocks;
    size_t out = kPad;
    *nbytes = 0;
    while (blocks.size() < r->window_blocks && r->cpos < r->map_len) {
        const uint8_t *h = r->map + r->cpos;
        size_t left = r->map_len - r->cpos;
        if (left < 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4)) {
            snprintf(err, err_len, "%s: not a BGZF block at offset %zu", r->path.c_str(), r->cpos);
            return TDT_BAM_E_FORMAT;
        }
        uint32_t xlen = le16(h + 10);
        int64_t bsize = -1;
        if (left >= 12 + (size_t)xlen) {
            for (uint32_t x = 0; x + 4 <= xlen;) {
                const uint8_t *sf = h + 12 + x;
                uint32_t slen = le16(sf + 2);
                if (sf[0] == 'B' && sf[1] == 'C' && slen == 2 && x + 6 <= xlen) bsize = (int64_t)le16(sf + 4) + 1;
                x += 4 + slen;
            }
        }
        if (bsize < 0 || (size_t)bsize > left || (size_t)bsize < 12 + (size_t)xlen + 8) {
            snprintf(err, err_len, "%s: bad or truncated BGZF block at offset %zu", r->path.c_str(), r->cpos);
            return TDT_BAM_E_FORMAT;
        }
        Block b;
        b.payload = h + 12 + xlen;
        b.clen = (uint32_t)(bsize - 12 - xlen - 8);
        b.crc = le32(h + bsize - 8);
        b.isize = le32(h + bsize - 4);
        if (b.isize > 65536) {
            snprintf(err, err_len, "%s: BGZF block inflates to %u bytes", r->path.c_str(), b.isize);
            return TDT_BAM_E_FORMAT;
        }
        b.dst = out;
        out += b.isize;
        r->cpos += (size_t)bsize;
        if (b.isize) blocks.push_back(b);  // the EOF marker and other empty blocks carry nothing
    }
    if (blocks.empty()) return 0;
    if (dst.size() < out) dst.resize(out + (out >> 3));
    std::atomic<size_t> next(0);
    std::atomic<int> bad(0);
    uint8_t *base = dst.data();
    const char *zenv = getenv("TDT_BAM_ZLIB");
    const bool zlib_only = zenv && zenv[0] == '1';
    auto work = [&](int w) {
        InflateWorker &ws = *r->inflate_workers[(size_t)w];
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= blocks.size()) break;
            const Block &b = blocks[i];
            uint8_t *out = base + b.dst;
            // the trailer's CRC decides whether the fast decoder's bytes stand; zlib inflates whatever it does not
            if (!zlib_only && tdtz::inflate_raw(ws.fast, b.payload, b.clen, out, b.isize) && tdtz::crc32_of(out, b.isize) == b.crc)
                continue;
            if (!ws.zs_ready) {
                memset(&ws.zs, 0, sizeof ws.zs);
                if (inflateInit2(&ws.zs, -15) != Z_OK) { bad = 1; continue; }
                ws.zs_ready = true;
            }
            ws.zs.next_in = const_cast<Bytef *>(b.payload);
            ws.zs.avail_in = b.clen;
            ws.zs.next_out = out;
            ws.zs.avail_out = b.isize;
            int rc = inflate(&ws.zs, Z_FINISH);
            if (rc != Z_STREAM_END || ws.zs.avail_out != 0 || (uint32_t)crc32(crc32(0L, Z_NULL, 0), out, b.isize) != b.crc)
                bad = 1;
            inflateReset(&ws.zs);
        }
    };
    if (blocks.size() < 4) work(0);
    else r->inflate_pool->run(work);
    if (bad) {
        snprintf(err, err_len, "%s: a BGZF block failed to inflate (corrupt data or CRC mismatch)", r->path.c_str());
        return TDT_BAM_E_FORMAT;
    }
    *nbytes = out - kPad;
    return 1;
}

void start_prefetch(tdt_bam_reader *r) {
    r->bg_running = true;
    r->bg = std::thread([r]() {
        try {
            r->bg_rc = inflate_window(r, r->nbuf, &r->bg_bytes, r->bg_err, sizeof r->bg_err);
        } catch (const std::exception &e) {   // bad_alloc, thread creation: no exception may leave the thread
            snprintf(r->bg_err, sizeof r->bg_err, "%s: %s", r->path.c_str(), e.what());
            r->bg_rc = TDT_BAM_E_IO;
        }
    });
}

// Makes the next window current, keeping the unread tail of the old one in front of it, and starts inflating the
// window after it.  Returns 1 if bytes were added, 0 at end of file, < 0 on error.
int refill(tdt_bam_reader *r) {
    if (r->eof) return 0;
    if (!r->bg_running) start_prefetch(r);
    r->bg.join();
    r->bg_running = false;
    if (r->bg_rc < 0) return fail(r->bg_rc, "%s", r->bg_err);
    if (r->bg_rc == 0) {
        r->eof = true;
        return 0;
    }
    const size_t left = r->bend - r->bpos;
    if (left > kPad) {  // a record longer than the headroom: make room the slow way
        std::vector<uint8_t> joined(left + r->bg_bytes + kPad);
        memcpy(joined.data() + kPad, r->buf.data() + r->bpos, left);
        memcpy(joined.data() + kPad + left, r->nbuf.data() + kPad, r->bg_bytes);
        r->nbuf.swap(joined);
        r->buf.swap(r->nbuf);
        r->bpos = kPad;
        r->bend = kPad + left + r->bg_bytes;
    } else {
        if (left) memcpy(r->nbuf.data() + kPad - left, r->buf.data() + r->bpos, left);
        r->buf.swap(r->nbuf);
        r->bpos = kPad - left;
        r->bend = kPad + r->bg_bytes;
    }
    start_prefetch(r);  // into the window the caller has finished with
    return 1;
}

// at least n unread bytes in the window (used for the header only)
int ensure(tdt_bam_reader *r, size_t n) {
    while (r->bend - r->bpos < n) {
        int rc = refill(r);
        if (rc < 0) return rc;
        if (rc == 0) return fail(TDT_BAM_E_FORMAT, "%s: truncated BAM header", r->path.c_str());
    }
    return 1;
}

int read_header(tdt_bam_reader *r) {
    int rc;
    if ((rc = ensure(r, 12)) < 0) return rc;
    const uint8_t *p = r->buf.data() + r->bpos;
    if (memcmp(p, "BAM\1", 4) != 0) return fail(TDT_BAM_E_FORMAT, "%s is not a BAM file", r->path.c_str());
    uint32_t l_text = le32(p + 4);
    if ((rc = ensure(r, 12 + (size_t)l_text)) < 0) return rc;
    p = r->buf.data() + r->bpos;
    r->text.assign((const char *)p + 8, l_text);
    size_t z = r->text.find('\0');
    if (z != std::string::npos) r->text.resize(z);
    uint32_t n_ref = le32(p + 8 + l_text);
    r->bpos += 12 + (size_t)l_text;
    for (uint32_t i = 0; i < n_ref; ++i) {
        if ((rc = ensure(r, 4)) < 0) return rc;
        uint32_t l_name = le32(r->buf.data() + r->bpos);
        if ((rc = ensure(r, 8 + (size_t)l_name)) < 0) return rc;
        p = r->buf.data() + r->bpos;
        r->ref_names.emplace_back((const char *)p + 4, l_name ? strnlen((const char *)p + 4, l_name) : 0);
        r->ref_lens.push_back((int32_t)le32(p + 4 + l_name));
        r->bpos += 8 + (size_t)l_name;
    }
    return 0;
}

// "SA" among the aux fields [p, e)?  Walks the typed fields (SAM spec 4.2.4).
bool aux_has_sa(const uint8_t *p, const uint8_t *e) {
    while (p + 3 <= e) {
        bool sa = p[0] == 'S' && p[1] == 'A';
        uint8_t t = p[2];
        p += 3;
        if (sa) return true;
        switch (t) {
            case 'A': case 'c': case 'C': p += 1; break;
            case 's': case 'S': p += 2; break;
            case 'i': case 'I': case 'f': p += 4; break;
            case 'Z': case 'H':
                while (p < e && *p) ++p;
                ++p;
                break;
            case 'B': {
                if (p + 5 > e) return false;
                uint8_t st = p[0];
                uint32_t cnt = le32(p + 1);
                size_t w = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
                p += 5 + (size_t)cnt * w;
                break;
            }
            default: return false;  // unknown type: cannot walk further
        }
    }
    return false;
}

}  // namespace

extern "C" {

const char *tdt_bam_last_error(void) { return g_err; }

static int open_impl(const char *path, int threads, tdt_bam_reader **out);

int tdt_bam_open(const char *path, int threads, tdt_bam_reader **out) {
    if (!path || !out) return fail(TDT_BAM_E_ARG, "tdt_bam_open: null argument");
    *out = nullptr;
    try {   // no exception crosses the C boundary
        return open_impl(path, threads, out);
    } catch (const std::exception &e) {
        return fail(TDT_BAM_E_IO, "%s: %s", path, e.what());
    }
}

static int open_impl(const char *path, int threads, tdt_bam_reader **out) {
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(TDT_BAM_E_IO, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size <= 0) {
        close(fd);
        return fail(TDT_BAM_E_IO, "%s is empty or cannot be examined", path);
    }
    void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    if (m == MAP_FAILED) {
        close(fd);
        return fail(TDT_BAM_E_IO, "cannot map %s", path);
    }
    madvise(m, (size_t)st.st_size, MADV_SEQUENTIAL);
    tdt_bam_reader *r = new tdt_bam_reader;
    r->fd = fd;
    r->map = (const uint8_t *)m;
    r->map_len = (size_t)st.st_size;
    r->path = path;
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    r->threads = std::max(1, threads);
    int rc;
    try {
        r->inflate_pool.reset(new Pool(r->threads));
        r->parse_pool.reset(new Pool(std::min(r->threads, 8)));   // decoding record fields is latency-bound: a few threads cover it
        for (int i = 0; i < r->threads; ++i) r->inflate_workers.emplace_back(new InflateWorker);
        rc = read_header(r);
    } catch (...) {   // thread creation, bad_alloc: the reader must not leak
        tdt_bam_close(r);
        throw;
    }
    if (rc < 0) {
        tdt_bam_close(r);
        return rc;
    }
    *out = r;
    return TDT_BAM_OK;
}

void tdt_bam_close(tdt_bam_reader *r) {
    if (!r) return;
    if (r->bg_running) r->bg.join();
    if (getenv("TDT_BAM_TIMING"))
        fprintf(stderr, "tdt_bam %s: waited %.3f s for inflated windows, boundary walk %.3f s, field decode %.3f s\n",
                r->path.c_str(), r->t_wait, r->t_walk, r->t_decode);
    if (r->map) munmap(const_cast<uint8_t *>(r->map), r->map_len);
    if (r->fd >= 0) close(r->fd);
    delete r;
}

const char *tdt_bam_header_text(const tdt_bam_reader *r, int64_t *len) {
    if (len) *len = (int64_t)r->text.size();
    return r->text.c_str();
}
int32_t tdt_bam_n_ref(const tdt_bam_reader *r) { return (int32_t)r->ref_names.size(); }
const char *tdt_bam_ref_name(const tdt_bam_reader *r, int32_t i) {
    return (i >= 0 && (size_t)i < r->ref_names.size()) ? r->ref_names[i].c_str() : nullptr;
}
int32_t tdt_bam_ref_len(const tdt_bam_reader *r, int32_t i) {
    return (i >= 0 && (size_t)i < r->ref_lens.size()) ? r->ref_lens[i] : -1;
}

static int64_t read_columns_impl(tdt_bam_reader *r, int64_t max_reads, int32_t *ref_id, int32_t *pos, int32_t *end,
                                 int32_t *mate_ref, int32_t *mate_pos, int32_t *tlen, uint16_t *flag, uint8_t *mapq,
                                 uint32_t *cig_first, uint32_t *cig_last, uint8_t *has_sa, int64_t *rec_off);

int64_t tdt_bam_read_columns(tdt_bam_reader *r, int64_t max_reads, int32_t *ref_id, int32_t *pos, int32_t *end,
                             int32_t *mate_ref, int32_t *mate_pos, int32_t *tlen, uint16_t *flag, uint8_t *mapq,
                             uint32_t *cig_first, uint32_t *cig_last, uint8_t *has_sa, int64_t *rec_off) {
    if (!r || max_reads < 0) return fail(TDT_BAM_E_ARG, "tdt_bam_read_columns: bad argument");
    try {
        return read_columns_impl(r, max_reads, ref_id, pos, end, mate_ref, mate_pos, tlen, flag, mapq, cig_first, cig_last,
                                 has_sa, rec_off);
    } catch (const std::exception &e) {
        return fail(TDT_BAM_E_IO, "%s: %s", r->path.c_str(), e.what());
    }
}

static int64_t read_columns_impl(tdt_bam_reader *r, int64_t max_reads, int32_t *ref_id, int32_t *pos, int32_t *end,
                                 int32_t *mate_ref, int32_t *mate_pos, int32_t *tlen, uint16_t *flag, uint8_t *mapq,
                                 uint32_t *cig_first, uint32_t *cig_last, uint8_t *has_sa, int64_t *rec_off) {
    int64_t n = 0;
    while (n == 0 && max_reads > 0) {
        // a complete record at the read position?  otherwise bring in the next window (this is the only place
        // the buffer moves, so the offsets handed out below stay valid until the next call)
        size_t have = r->bend - r->bpos;
        bool complete = have >= 4 && have >= 4 + (size_t)le32(r->buf.data() + r->bpos);
        if (!complete) {
            const double t0 = now_s();
            int rc = refill(r);
            r->t_wait += now_s() - t0;
            if (rc < 0) return rc;
            if (rc == 0) {
                if (r->bend - r->bpos != 0)
                    return fail(TDT_BAM_E_FORMAT, "%s: truncated record at end of file", r->path.c_str());
                return 0;
            }
        }
        // 1. the record boundaries of this batch: a walk over the block_size words (sequential by nature)
        const double t_a = now_s();
        const uint8_t *base = r->buf.data();
        r->rec_at.clear();
        size_t at = r->bpos;
        while ((int64_t)r->rec_at.size() < max_reads) {
            size_t avail = r->bend - at;
            if (avail < 4) break;
            uint32_t bs = le32(base + at);
            if (bs < 32) return fail(TDT_BAM_E_FORMAT, "%s: record of %u bytes", r->path.c_str(), bs);
            if (avail < 4 + (size_t)bs) break;
            r->rec_at.push_back((int64_t)at);
            at += 4 + (size_t)bs;
            // the walk is a chain of dependent loads into memory other cores have just written (~70 ns each): short reads
            // have records of nearly one size, so the lines around "eight records ahead" are requested now
            const uint8_t *ahead = base + at + 8 * (4 + (size_t)bs);
            if (ahead + 128 < base + r->bend) {
                __builtin_prefetch(ahead - 64);
                __builtin_prefetch(ahead);
                __builtin_prefetch(ahead + 64);
            }
        }
        n = (int64_t)r->rec_at.size();
        const double t_b = now_s();
        r->t_walk += t_b - t_a;
        // 2. the fields of every record into the columns, on the parse pool when the batch is worth it
        std::atomic<int64_t> next(0);
        std::atomic<int> bad(0);
        const int64_t *rec_at = r->rec_at.data();
        auto decode = [&](int) {
            for (;;) {
                const int64_t lo = next.fetch_add(4096), hi = std::min<int64_t>(lo + 4096, n);
                if (lo >= n) break;
                for (int64_t k = lo; k < hi; ++k) {
                    const uint8_t *rec = base + rec_at[k];
                    const uint32_t bs = le32(rec);
                    const uint8_t *c = rec + 4;
                    int32_t rid = (int32_t)le32(c), p0 = (int32_t)le32(c + 4);
                    uint32_t l_name = c[8];
                    uint32_t n_cig = le16(c + 12);
                    uint16_t fl = le16(c + 14);
                    uint32_t l_seq = le32(c + 16);
                    size_t fixed = 32 + (size_t)l_name + 4 * (size_t)n_cig;
                    size_t aux_at = fixed + ((size_t)l_seq + 1) / 2 + l_seq;
                    if (aux_at > bs) {
                        bad = 1;
                        return;
                    }
                    const uint8_t *cg = c + 32 + l_name;
                    if (ref_id) ref_id[k] = rid;
                    if (pos) pos[k] = p0;
                    if (end) {
                        if ((fl & 4) || n_cig == 0) {
                            end[k] = -1;
                        } else {
                            int64_t e = p0;
                            for (uint32_t q = 0; q < n_cig; ++q) {
                                uint32_t w = le32(cg + 4 * q);
                                if (kConsumesRef[w & 15]) e += w >> 4;
                            }
                            end[k] = (int32_t)e;
                        }
                    }
                    if (mate_ref) mate_ref[k] = (int32_t)le32(c + 20);
                    if (mate_pos) mate_pos[k] = (int32_t)le32(c + 24);
                    if (tlen) tlen[k] = (int32_t)le32(c + 28);
                    if (flag) flag[k] = fl;
                    if (mapq) mapq[k] = c[9];
                    if (cig_first) cig_first[k] = n_cig ? le32(cg) : 0;
                    if (cig_last) cig_last[k] = n_cig ? le32(cg + 4 * (n_cig - 1)) : 0;
                    if (has_sa) has_sa[k] = aux_has_sa(c + aux_at, c + bs) ? 1 : 0;
                    if (rec_off) rec_off[k] = rec_at[k];
                }
            }
        };
        if (n >= 16384 && r->parse_pool && r->parse_pool->size() > 1) r->parse_pool->run(decode);
        else decode(0);
        r->t_decode += now_s() - t_b;
        if (bad) return fail(TDT_BAM_E_FORMAT, "%s: record fields exceed its size", r->path.c_str());
        r->bpos = at;
    }
    return n;
}

int tdt_bam_inflate_raw(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_len) {
    if (!in || !out || in_len < 0 || out_len < 0) return fail(TDT_BAM_E_ARG, "tdt_bam_inflate_raw: bad argument");
    tdtz::Inflater *st = new (std::nothrow) tdtz::Inflater;
    if (!st) return fail(TDT_BAM_E_IO, "out of memory");
    const bool ok = tdtz::inflate_raw(*st, in, (size_t)in_len, out, (size_t)out_len);
    delete st;
    return ok ? 1 : 0;
}

uint32_t tdt_bam_crc32(const uint8_t *data, int64_t len) { return (data && len > 0) ? tdtz::crc32_of(data, (size_t)len) : 0u; }

const uint8_t *tdt_bam_batch_data(const tdt_bam_reader *r, int64_t *len) {
    if (len) *len = (int64_t)r->bend;
    return r->buf.data();
}

}  // extern "C"

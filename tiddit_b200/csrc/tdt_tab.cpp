// tdt_tab.cpp -- libtdt_tab.so: the signal tab files of tiddit_signal / tiddit_contig_analysis as columns
// (include/tdt_tab.h).  Host code only: mmap, a line-aligned split over std::threads, field parsing, string interning.
#include "../../include/tdt_tab.h"

#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

namespace {

thread_local char g_err[512] = {0};

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

struct Str {          // a field inside the mapped file
    const char *p;
    uint32_t len;
    uint32_t hash;
};

inline uint32_t hash_bytes(const char *p, uint32_t n) {   // FNV-1a folded to 32 bits
    uint64_t h = 1469598103934665603ull;
    for (uint32_t i = 0; i < n; i++) h = (h ^ (unsigned char)p[i]) * 1099511628211ull;
    return (uint32_t)(h ^ (h >> 32));
}

// strings -> ids in order of first appearance (open addressing; the bytes are copied into one blob)
struct Interner {
    std::vector<uint64_t> slots;     // (hash << 32) | (id + 1), 0 = empty: a probe rejects on the hash without a second miss
    std::vector<int64_t> offsets{0}; // per id + 1
    std::string blob;
    size_t mask = 0, count = 0;

    void rehash(size_t cap) {
        std::vector<uint64_t> fresh(cap, 0);
        mask = cap - 1;
        for (uint64_t v : slots) {
            if (!v) continue;
            size_t s = (size_t)(v >> 32) & mask;
            while (fresh[s]) s = (s + 1) & mask;
            fresh[s] = v;
        }
        slots.swap(fresh);
    }

    void reserve(size_t extra) {   // room for `extra` more strings without rehashing on the way
        size_t want = (count + extra + 1) * 2, cap = slots.empty() ? 1024 : slots.size();
        while (cap < want) cap *= 2;
        if (cap != slots.size()) rehash(cap);
        offsets.reserve(offsets.size() + extra);
    }

    // the few-valued columns (contigs, orientations) repeat their previous value most of the time
    Str last{nullptr, 0, 0};
    int32_t last_id = -1;
    int32_t intern_cached(const Str &f) {
        if (last_id >= 0 && f.hash == last.hash && f.len == last.len && memcmp(f.p, last.p, f.len) == 0) return last_id;
        last_id = intern(f);
        last = f;
        return last_id;
    }

    int32_t intern(const Str &f) {
        if ((count + 1) * 2 > slots.size()) rehash(slots.empty() ? 1024 : slots.size() * 2);
        size_t s = f.hash & mask;
        while (true) {
            const uint64_t v = slots[s];
            if (!v) break;
            if ((uint32_t)(v >> 32) == f.hash) {
                const int32_t id = (int32_t)(v & 0xffffffffu) - 1;
                const int64_t o = offsets[id], l = offsets[id + 1] - o;
                if ((uint32_t)l == f.len && memcmp(blob.data() + o, f.p, f.len) == 0) return id;
            }
            s = (s + 1) & mask;
        }
        const int32_t id = (int32_t)count++;
        slots[s] = ((uint64_t)f.hash << 32) | (uint64_t)(uint32_t)(id + 1);
        blob.append(f.p, f.len);
        offsets.push_back((int64_t)blob.size());
        return id;
    }
};

struct Chunk {        // what one thread parsed: columns in file order
    std::vector<Str> name, chrA, chrB, oriA, oriB;
    std::vector<int64_t> num[6];
    bool irregular = false;
};

inline bool is_space(char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; }

// a text field as the line reader would see it unchanged: non-empty, no white space at either end
inline bool text_field(const char *b, const char *e, Str &out) {
    if (e <= b || is_space(b[0]) || is_space(e[-1])) return false;
    out.p = b;
    out.len = (uint32_t)(e - b);
    out.hash = hash_bytes(b, out.len);
    return true;
}

// int(field) for the plain case: optional sign, decimal digits only, fits int64 comfortably
inline bool int_field(const char *b, const char *e, int64_t &out) {
    if (e <= b) return false;
    bool neg = false;
    if (*b == '-' || *b == '+') {
        neg = *b == '-';
        b++;
    }
    if (e <= b || e - b > 18) return false;
    int64_t v = 0;
    for (; b < e; b++) {
        if (*b < '0' || *b > '9') return false;
        v = v * 10 + (*b - '0');
    }
    out = neg ? -v : v;
    return true;
}

void parse_range(const char *lo, const char *hi, int kind, Chunk &c) {
    const int need = kind == TDT_TAB_DISCORDANTS ? 9 : 11;
    const char *p = lo;
    {   // one allocation per column: count the lines first (memchr runs at memory speed)
        size_t lines = 0;
        for (const char *q = lo; q < hi;) {
            const char *nl = (const char *)memchr(q, '\n', (size_t)(hi - q));
            lines++;
            if (!nl) break;
            q = nl + 1;
        }
        c.name.reserve(lines);
        c.chrA.reserve(lines);
        c.chrB.reserve(lines);
        c.oriA.reserve(lines);
        c.oriB.reserve(lines);
        for (int i = 0; i < 6; i++) c.num[i].reserve(lines);
    }
    while (p < hi) {
        const char *eol = (const char *)memchr(p, '\n', (size_t)(hi - p));
        if (!eol) eol = hi;
        const char *fb[11], *fe[11];
        int nf = 0;
        const char *q = p;
        while (true) {
            const char *t = (const char *)memchr(q, '\t', (size_t)(eol - q));
            const char *end = t ? t : eol;
            if (nf < 11) {
                fb[nf] = q;
                fe[nf] = end;
            }
            nf++;
            if (!t) break;
            q = t + 1;
        }
        if (nf < need || (kind == TDT_TAB_DISCORDANTS && nf != 9)) {
            c.irregular = true;
            return;
        }
        Str name, ca, cb, oa, ob;
        int64_t v[6] = {0, 0, 0, 0, 0, 0};
        bool ok = text_field(fb[0], fe[0], name) && text_field(fb[1], fe[1], ca) && text_field(fb[2], fe[2], cb);
        if (kind == TDT_TAB_DISCORDANTS) {
            ok = ok && int_field(fb[3], fe[3], v[0]) && int_field(fb[4], fe[4], v[1]) && text_field(fb[5], fe[5], oa) &&
                 int_field(fb[6], fe[6], v[2]) && int_field(fb[7], fe[7], v[3]) && text_field(fb[8], fe[8], ob);
        } else {
            ok = ok && int_field(fb[3], fe[3], v[0]) && text_field(fb[4], fe[4], oa) && int_field(fb[5], fe[5], v[1]) &&
                 text_field(fb[6], fe[6], ob) && int_field(fb[7], fe[7], v[2]) && int_field(fb[8], fe[8], v[3]) &&
                 int_field(fb[9], fe[9], v[4]) && int_field(fb[10], fe[10], v[5]);
            // the last field the reader uses must not run into white space the reader's rstrip() would remove
        }
        if (!ok) {
            c.irregular = true;
            return;
        }
        c.name.push_back(name);
        c.chrA.push_back(ca);
        c.chrB.push_back(cb);
        c.oriA.push_back(oa);
        c.oriB.push_back(ob);
        for (int i = 0; i < 6; i++) c.num[i].push_back(v[i]);
        p = eol < hi ? eol + 1 : hi;
    }
}

}  // namespace

struct tdt_tab_set {
    Interner names, contigs, oris;
    std::vector<int32_t> i32[5];
    std::vector<int64_t> i64[6];
};

extern "C" {

const char *tdt_tab_last_error(void) { return g_err; }

tdt_tab_set *tdt_tab_new(void) { return new tdt_tab_set(); }

void tdt_tab_free(tdt_tab_set *set) { delete set; }

int64_t tdt_tab_parse(tdt_tab_set *set, const char *path, int kind, int threads) {
    if (!set || !path || kind < 0 || kind > 2) return fail(TDT_TAB_E_ARG, "bad argument");
    const int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(TDT_TAB_E_IO, "cannot open %s", path);
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return fail(TDT_TAB_E_IO, "cannot stat %s", path);
    }
    const size_t size = (size_t)st.st_size;
    if (size == 0) {
        close(fd);
        return 0;
    }
    const char *data = (const char *)mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (data == MAP_FAILED) return fail(TDT_TAB_E_IO, "cannot map %s", path);
    int T = threads > 0 ? threads : (int)std::thread::hardware_concurrency();
    if (T < 1) T = 1;
    if ((size_t)T > size / (1 << 16) + 1) T = (int)(size / (1 << 16) + 1);
    // line-aligned cuts
    std::vector<const char *> cut(T + 1);
    cut[0] = data;
    cut[T] = data + size;
    for (int t = 1; t < T; t++) {
        const char *p = data + size * (size_t)t / (size_t)T;
        if (p < cut[t - 1]) p = cut[t - 1];
        const char *nl = (const char *)memchr(p, '\n', (size_t)(data + size - p));
        cut[t] = nl ? nl + 1 : data + size;
    }
    const bool timing = getenv("TDT_TAB_TIMING") != nullptr;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    const double t_begin = now();
    std::vector<Chunk> chunks(T);
    {
        std::vector<std::thread> pool;
        for (int t = 1; t < T; t++) pool.emplace_back(parse_range, cut[t], cut[t + 1], kind, std::ref(chunks[t]));
        parse_range(cut[0], cut[1], kind, chunks[0]);
        for (auto &th : pool) th.join();
    }
    const double t_parsed = now();
    int64_t added = 0;
    bool irregular = false;
    for (auto &c : chunks) irregular = irregular || c.irregular;
    if (!irregular) {
        size_t total = 0;
        for (auto &c : chunks) total += c.name.size();
        set->names.reserve(total);
        const size_t base = set->i32[0].size();
        for (int j = 0; j < 5; j++) set->i32[j].resize(base + total);
        for (int j = 0; j < 6; j++) set->i64[j].reserve(base + total);
        // the few-valued columns are interned first (sequential, cached), in the order the reader meets them per
        // record: chrA, chrB / oriA, oriB; then the names, the only large table
        size_t at = base;
        for (auto &c : chunks) {
            const size_t k = c.name.size();
            for (size_t i = 0; i < k; i++, at++) {
                set->i32[1][at] = set->contigs.intern_cached(c.chrA[i]);
                set->i32[2][at] = set->contigs.intern(c.chrB[i]);
                set->i32[3][at] = set->oris.intern_cached(c.oriA[i]);
                set->i32[4][at] = set->oris.intern(c.oriB[i]);
                set->i32[0][at] = set->names.intern(c.name[i]);
            }
            for (int j = 0; j < 6; j++) set->i64[j].insert(set->i64[j].end(), c.num[j].begin(), c.num[j].end());
            added += (int64_t)k;
        }
        set->contigs.last_id = set->oris.last_id = -1;   // the cached pointers die with the mapping
    }
    munmap((void *)data, size);
    if (timing) fprintf(stderr, "tdt_tab_parse %s: %d threads, parse %.3f s, intern %.3f s\n", path, T, t_parsed - t_begin, now() - t_parsed);
    if (irregular) return fail(TDT_TAB_IRREGULAR, "%s is not a perfectly regular tab file", path);
    return added;
}

int64_t tdt_tab_n(const tdt_tab_set *set) { return set ? (int64_t)set->i32[0].size() : 0; }

const int32_t *tdt_tab_col_i32(const tdt_tab_set *set, int which) {
    return (set && which >= 0 && which < 5) ? set->i32[which].data() : nullptr;
}

const int64_t *tdt_tab_col_i64(const tdt_tab_set *set, int which) {
    return (set && which >= 0 && which < 6) ? set->i64[which].data() : nullptr;
}

int64_t tdt_tab_table(const tdt_tab_set *set, int table, const char **blob, const int64_t **offsets) {
    if (!set || table < 0 || table > 2) return fail(TDT_TAB_E_ARG, "bad argument");
    const Interner &t = table == 0 ? set->names : (table == 1 ? set->contigs : set->oris);
    if (blob) *blob = t.blob.data();
    if (offsets) *offsets = t.offsets.data();
    return (int64_t)t.count;
}

}  // extern "C"

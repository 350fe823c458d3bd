// tdt_tab.cpp -- libtdt_tab.so: the signal tab files of tiddit_signal / tiddit_contig_analysis as columns
// (include/tdt_tab.h).  Host code only: mmap, a line-aligned split over std::threads, field parsing, string interning.
// Every phase runs on all threads: the few-valued columns (contigs, orientations) are interned per chunk and the small
// chunk tables merged in file order; the read names -- one table of millions of strings -- go through a hash-sharded
// table (NameTable below) in which every thread resolves the names of its own shards.
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
    std::vector<Str> name;
    std::vector<int32_t> cA, cB, oA, oB;   // ids in the chunk's OWN small tables (merged into the set's afterwards)
    Interner contigs, oris;
    std::vector<int64_t> num[6];
    bool irregular = false;
};

// The read-name table, sharded by the top hash bits so that several threads can intern one file's names at once and
// still hand out ids in order of first appearance:
//   A (parallel)   thread t scans all names of the file in order and resolves those of ITS shards: a name already in
//                  the table keeps its id; otherwise the first occurrence in this file is entered as PENDING (slot value
//                  = its position g in the file) and later occurrences point at it;
//   B (sequential) pending first occurrences get consecutive ids in file order, their bytes are appended to the blob;
//   C (parallel)   later occurrences copy the id of their first occurrence; D: pending slots become id slots.
struct NameShard {
    std::vector<uint64_t> slots;   // (hash << 32) | (id + 1), or (hash << 32) | 0x80000000 | g while pending; 0 = empty
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
};

struct NameTable {
    static constexpr int SHARD_SHIFT = 26, SHARDS = 64;   // shard = top 6 hash bits, slot = low bits
    NameShard shard[SHARDS];
    std::vector<int64_t> offsets{0};
    std::string blob;
    size_t count = 0;

    // phase A for the shards of thread t (of T): res[g] >= 0: id of a name already in the table; -2 - g0: same string
    // as position g0 <= g of this file (g0 == g: first occurrence).  -> bytes of the first occurrences found
    size_t resolve(const std::vector<Str> &flat, std::vector<int32_t> &res, int t, int T) {
        const size_t N = flat.size();
        size_t bytes = 0;
        for (int sh = t; sh < SHARDS; sh += T) {   // room for an even share of the file plus slack; grows on demand
            NameShard &S = shard[sh];
            size_t want = (S.count + N / SHARDS + N / SHARDS / 4 + 16) * 2, cap = S.slots.empty() ? 64 : S.slots.size();
            while (cap < want) cap *= 2;
            if (cap != S.slots.size()) S.rehash(cap);
        }
        for (size_t g = 0; g < N; g++) {
            const Str &f = flat[g];
            const int sh = (int)(f.hash >> SHARD_SHIFT);
            if (sh % T != t) continue;
            NameShard &S = shard[sh];
            if ((S.count + 1) * 2 > S.slots.size()) S.rehash(S.slots.size() * 2);
            size_t s = f.hash & S.mask;
            while (true) {
                const uint64_t v = S.slots[s];
                if (!v) {
                    S.slots[s] = ((uint64_t)f.hash << 32) | 0x80000000ull | (uint64_t)g;
                    S.count++;
                    res[g] = -2 - (int32_t)g;
                    bytes += f.len;
                    break;
                }
                if ((uint32_t)(v >> 32) == f.hash) {
                    const uint32_t low = (uint32_t)v;
                    if (low & 0x80000000u) {
                        const Str &h = flat[low & 0x7fffffffu];
                        if (h.len == f.len && memcmp(h.p, f.p, f.len) == 0) {
                            res[g] = -2 - (int32_t)(low & 0x7fffffffu);
                            break;
                        }
                    } else {
                        const int32_t id = (int32_t)low - 1;
                        const int64_t o = offsets[id], l = offsets[id + 1] - o;
                        if ((uint32_t)l == f.len && memcmp(blob.data() + o, f.p, f.len) == 0) {
                            res[g] = id;
                            break;
                        }
                    }
                }
                s = (s + 1) & S.mask;
            }
        }
        return bytes;
    }

    // phase D for the shards of thread t: pending slots -> id slots
    void settle(const std::vector<int32_t> &res, int t, int T) {
        for (int sh = t; sh < SHARDS; sh += T)
            for (uint64_t &v : shard[sh].slots)
                if ((uint32_t)v & 0x80000000u)
                    v = (v & 0xffffffff00000000ull) | (uint64_t)(uint32_t)(res[(uint32_t)v & 0x7fffffffu] + 1);
    }
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
        c.cA.reserve(lines);
        c.cB.reserve(lines);
        c.oA.reserve(lines);
        c.oB.reserve(lines);
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
        // in the order the reader meets them per record: chrA, chrB / oriA, oriB (the first of each mostly repeats)
        c.cA.push_back(c.contigs.intern_cached(ca));
        c.cB.push_back(c.contigs.intern(cb));
        c.oA.push_back(c.oris.intern_cached(oa));
        c.oB.push_back(c.oris.intern(ob));
        for (int i = 0; i < 6; i++) c.num[i].push_back(v[i]);
        p = eol < hi ? eol + 1 : hi;
    }
}

}  // namespace

struct tdt_tab_set {
    NameTable names;
    Interner contigs, oris;
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
        std::vector<size_t> first(T + 1, 0);          // chunk t holds records [first[t], first[t + 1]) of the file
        for (int t = 0; t < T; t++) first[t + 1] = first[t] + chunks[t].name.size();
        total = first[T];
        if (total >= (size_t)1 << 31 || set->names.count + total >= (size_t)1 << 31) {
            munmap((void *)data, size);
            return fail(TDT_TAB_E_ARG, "%s: more than 2^31 records / names", path);
        }
        const size_t base = set->i32[0].size();
        for (int j = 0; j < 5; j++) set->i32[j].resize(base + total);
        for (int j = 0; j < 6; j++) set->i64[j].resize(base + total);
        // the chunks' small tables merged in file order: a string keeps the id of its first appearance in the set
        std::vector<std::vector<int32_t>> cmap(T), omap(T);
        for (int t = 0; t < T; t++) {
            for (int which = 0; which < 2; which++) {
                const Interner &loc = which == 0 ? chunks[t].contigs : chunks[t].oris;
                Interner &glob = which == 0 ? set->contigs : set->oris;
                std::vector<int32_t> &map = which == 0 ? cmap[t] : omap[t];
                map.resize(loc.count);
                for (size_t k = 0; k < loc.count; k++) {
                    Str f;
                    f.p = loc.blob.data() + loc.offsets[k];
                    f.len = (uint32_t)(loc.offsets[k + 1] - loc.offsets[k]);
                    f.hash = hash_bytes(f.p, f.len);
                    map[k] = glob.intern(f);
                }
            }
        }
        // names: flat view of the file, then the phases of NameTable
        std::vector<Str> flat(total);
        std::vector<int32_t> res(total);
        std::vector<size_t> new_bytes(T, 0);
        auto run = [&](auto fn) {
            std::vector<std::thread> pool;
            for (int t = 1; t < T; t++) pool.emplace_back(fn, t);
            fn(0);
            for (auto &th : pool) th.join();
        };
        run([&](int t) {   // columns of chunk t into place
            const Chunk &c = chunks[t];
            const size_t k = c.name.size(), at = base + first[t];
            if (k) memcpy(&flat[first[t]], c.name.data(), k * sizeof(Str));
            for (size_t i = 0; i < k; i++) {
                set->i32[1][at + i] = cmap[t][c.cA[i]];
                set->i32[2][at + i] = cmap[t][c.cB[i]];
                set->i32[3][at + i] = omap[t][c.oA[i]];
                set->i32[4][at + i] = omap[t][c.oB[i]];
            }
            for (int j = 0; j < 6; j++)
                if (k) memcpy(&set->i64[j][at], c.num[j].data(), k * sizeof(int64_t));
        });
        run([&](int t) { new_bytes[t] = set->names.resolve(flat, res, t, T); });
        {   // first occurrences in file order: ids, bytes
            NameTable &nt = set->names;
            size_t add = 0;
            for (size_t b : new_bytes) add += b;
            nt.blob.reserve(nt.blob.size() + add);
            for (size_t g = 0; g < total; g++) {
                if (res[g] == -2 - (int32_t)g) {
                    res[g] = (int32_t)nt.count++;
                    nt.blob.append(flat[g].p, flat[g].len);
                    nt.offsets.push_back((int64_t)nt.blob.size());
                }
            }
        }
        run([&](int t) {   // later occurrences, the name column of chunk t, the pending slots of thread t's shards
            for (size_t g = first[t]; g < first[t + 1]; g++) {
                if (res[g] < 0) res[g] = res[(size_t)(-2 - res[g])];
                set->i32[0][base + g] = res[g];
            }
        });
        run([&](int t) { set->names.settle(res, t, T); });
        added = (int64_t)total;
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
    if (table == 0) {
        if (blob) *blob = set->names.blob.data();
        if (offsets) *offsets = set->names.offsets.data();
        return (int64_t)set->names.count;
    }
    const Interner &t = table == 1 ? set->contigs : set->oris;
    if (blob) *blob = t.blob.data();
    if (offsets) *offsets = t.offsets.data();
    return (int64_t)t.count;
}

}  // extern "C"

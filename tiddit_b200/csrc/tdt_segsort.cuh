// tdt_segsort.cuh -- hand-written segmented, stable sort of (u32 key, i32 value) pairs (sm_100a).
//
// The clustering path sorts twice, and both sorts are SEGMENTED: signals by posA inside every
// (chrA,chrB) pair (tiddit_cluster.pyx:152) and x-cluster members by posB inside every x-cluster
// (DBSCAN.py:79-81).  Because the segment is implicit in the position (segment s owns
// [off[s], off[s+1]) before and after the sort) the keys stay 32 bits -- the coordinate alone -- instead of
// the 64-bit (segment, coordinate) composites a flat device-wide sort needs.  Three size classes:
//
//   tiny   (<= tiny_max = 128 elements, only when the caller has a per-element segment id; nearly all x-clusters):
//          one thread per element counts the members of its segment that precede it -- no shared memory;
//   small  (<= SS_LOCAL_MAX = 2048): CTA k gathers the small segments that START in element window [k*W, (k+1)*W)
//          into shared memory and sorts them together by (local segment index, key) with 8-bit LSD passes:
//          one global read and one global write per element;
//   large  : two chains.  Generation 3 (tdt_segsort3.cuh, the default for sorts whose value is the element index): one
//          most-significant-digit partition round + one shared-memory finish per element.  LSD chain (sorts with
//          explicit values; TDT_SEGSORT=lsd): onesweep-style passes over 4096-element tiles that never straddle a
//          segment: digit histograms of all passes up front, then per pass a stable in-tile ranking, a per-digit
//          decoupled look-back over the EARLIER TILES OF THE SAME SEGMENT and a digit-ordered write through shared memory.
//
// Stable ranking (small + large): a warp owns a contiguous run of the tile and walks it 32 elements at a time;
// lanes with equal digits find each other through a shared-memory atomicOr of their lane bits, the lowest
// lane of a group bumps the warp-private digit counter.
#pragma once
#include <stdlib.h>
#include <string.h>

#include "tdt_common.cuh"

namespace tdt {

#ifndef TDT_SS_THREADS
#define TDT_SS_THREADS 256
#endif
constexpr int SS_THREADS = TDT_SS_THREADS;  // large-segment kernels (>= 256: thread d owns digit d)
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_TILE = 4096;
constexpr int SS_CHUNKS = SS_TILE / SS_THREADS;  // 32-element chunks per warp per tile
constexpr int SS_LTHREADS = 512;           // small-segment kernel
constexpr int SS_LWARPS = SS_LTHREADS / 32;
constexpr int SS_LOCAL_MAX = 2048;
constexpr int SS_WINDOW = 1024;
constexpr int SS_LOCAL_CAP = 3072;         // >= SS_WINDOW - 1 + SS_LOCAL_MAX
constexpr int SS_LCHUNKS = SS_LOCAL_CAP / SS_LTHREADS;
constexpr int SS_MAX_PASSES = 4;
constexpr int SS_ERR_KEY_RANGE = 1;
constexpr int SS_ERR_INTERNAL = 64;  // a work list overflowed its (proven) bound; above every caller's own codes

// third generation of the large-segment chain (tdt_segsort3.cuh): MSD partition rounds + shared-memory finish
constexpr int M3_CAP = 8192;              // elements of a finish batch
constexpr int M3_THREADS = 512;
constexpr int M3_EPT = M3_CAP / M3_THREADS;
constexpr int M3_SLOT_BITS = 13;
constexpr int M3_NSLOT = 1 << M3_SLOT_BITS;
constexpr int M3_ROUNDS = 4;              // 8-bit digits of a 32-bit key
#ifndef TDT_M3_FIN_WAVES
#define TDT_M3_FIN_WAVES 8                // finish grid <= this many waves of 2 CTAs per SM (1 = persistent); B200, 30X posA
                                          // sort: 1 -> 0.58 ms (before the later tuning), 4 -> 0.405, 8 -> 0.386, 16 -> 0.396
#endif

struct M3Range {   // a contiguous element range holding every element of key range [klo, klo + (256 << shift))
    int64_t start;
    int32_t size, tile_base;
    uint32_t klo;
    int32_t shift;
    int32_t seg;       // the input segment the range belongs to
    int32_t pad;
};

struct M3Tile {    // a 4096-element tile of a queued range: everything the hist / pass kernels need to start loading
    int64_t t0;    // first element
    int32_t cnt;   // elements (the last tile of a range may be short)
    int32_t rng;   // its range
};

struct M3Batch {   // finish work item: `count` elements at `start`, keys in [klo, klo + (nslots << sh))
    int64_t start;
    int32_t count;
    uint32_t klo;
    int32_t sh, nslots, level, flags;
};

struct M3Counters {
    int32_t n_rng[M3_ROUNDS], n_tiles[M3_ROUNDS], n_batches[2];
    int32_t pad[6];
};

struct M3Layout {
    M3Counters *cnt;
    M3Range *rng[M3_ROUNDS];
    M3Tile *tile_rng[M3_ROUNDS];
    uint32_t *hist[M3_ROUNDS];   // [range][256]
    M3Batch *batch[2];           // list 0: whole segments + round-0 groups; list 1: the later rounds
    int64_t rng_max, tiles_max, batch_max;
};

struct SSLarge {
    int64_t start, size;
    int32_t tile_base, pad;
};

struct SSCounters {
    int32_t n_large, n_tiles;
    int32_t n_small;   // segments of the small class (the small-segment kernel leaves at once when there are none)
    int32_t pad[5];
};

struct SSLayout {  // carved out of the caller's temp storage
    SSCounters *cnt;
    SSLarge *large;
    int32_t *tile_seg;
    int32_t *win_lo, *win_hi;
    uint32_t *win_cnt;  // elements of the small segments starting in the window
    uint32_t *ghist;   // [n_large][SS_MAX_PASSES][256]
    uint32_t *status;  // [n_tiles][256]
    int64_t nlarge_max, tiles_max, nwin;
    size_t zero_bytes;  // leading region (counters + status) that one memset initialises
};

static inline size_t ss_align(size_t b) { return (b + 255) & ~(size_t)255; }

static inline int64_t ss_nlarge_max(int64_t n, int64_t nseg_max) {
    int64_t a = n / (SS_LOCAL_MAX + 1) + 1;
    return a < nseg_max ? a : (nseg_max > 0 ? nseg_max : 1);
}

// Work-list bounds.  A queued range holds > M3_CAP / 4 elements (buckets beyond a batch, pile-up buckets), so a round
// has at most n / (M3_CAP / 4) of them.  Batches: a whole segment (> SS_LOCAL_MAX elements) or a group of buckets; two
// neighbouring groups of a run of ordinary buckets together exceed M3_CAP, and a run ends at a dense or queued bucket
// (> M3_CAP / 4 elements) or at the end of its range: <= 6 n / M3_CAP + ranges + 1 groups per round, four rounds.
static inline int64_t m3_rng_max(int64_t n) { return n / (M3_CAP / 4) + 1; }
static inline int64_t m3_batch_max(int64_t n) { return 16 * m3_rng_max(n) + 64; }
static inline int64_t ss_tiles_max(int64_t n, int64_t nl) {
    const int64_t r = m3_rng_max(n);
    return n / SS_TILE + (nl > r ? nl : r) + 1;
}
static inline size_t m3_temp_bytes(int64_t n) {
    const int64_t r = m3_rng_max(n), tiles = ss_tiles_max(n, 0);
    return (size_t)M3_ROUNDS * (((size_t)r * sizeof(M3Range) + 255) / 256 * 256 + ((size_t)tiles * sizeof(M3Tile) + 255) / 256 * 256 +
                                ((size_t)r * 256 * 4 + 255) / 256 * 256) +
           2 * (((size_t)m3_batch_max(n) * sizeof(M3Batch) + 255) / 256 * 256) + 512;
}

static inline size_t segsort1_temp_bytes(int64_t n, int64_t nseg_max);
// what callers reserve
static inline size_t segsort_temp_bytes(int64_t n, int64_t nseg_max) { return segsort1_temp_bytes(n, nseg_max); }

static inline size_t segsort1_temp_bytes(int64_t n, int64_t nseg_max) {
    const int64_t nl = ss_nlarge_max(n, nseg_max);
    const int64_t tiles = ss_tiles_max(n, nl);
    const int64_t nwin = (n + SS_WINDOW - 1) / SS_WINDOW + 1;
    return ss_align(sizeof(SSCounters)) + ss_align((size_t)nl * sizeof(SSLarge)) + ss_align((size_t)tiles * 4) +
           3 * ss_align((size_t)nwin * 4) + ss_align((size_t)nl * SS_MAX_PASSES * 256 * 4) +
           ss_align((size_t)tiles * 256 * 4) + m3_temp_bytes(n) + 1024;
}

static inline SSLayout ss_layout(void *temp, int64_t n, int64_t nseg_max) {
    SSLayout L;
    L.nlarge_max = ss_nlarge_max(n, nseg_max);
    L.tiles_max = ss_tiles_max(n, L.nlarge_max);
    L.nwin = (n + SS_WINDOW - 1) / SS_WINDOW + 1;
    char *p = (char *)temp;
    L.cnt = (SSCounters *)p;
    p += ss_align(sizeof(SSCounters));
    L.status = (uint32_t *)p;
    p += ss_align((size_t)L.tiles_max * 256 * 4);
    L.win_cnt = (uint32_t *)p;
    p += ss_align((size_t)L.nwin * 4);
    L.win_hi = (int32_t *)p;  // (largest small segment that starts in the window) + 1; 0 = none
    p += ss_align((size_t)L.nwin * 4);
    L.win_lo = (int32_t *)p;  // INT32_MAX - (smallest small segment that starts in the window); 0 = none
    p += ss_align((size_t)L.nwin * 4);
    L.zero_bytes = (size_t)(p - (char *)temp);  // counters, status, the three window arrays: one memset
    L.large = (SSLarge *)p;
    p += ss_align((size_t)L.nlarge_max * sizeof(SSLarge));
    L.tile_seg = (int32_t *)p;
    p += ss_align((size_t)L.tiles_max * 4);
    L.ghist = (uint32_t *)p;
    return L;
}

// the generation-3 arrays live behind generation 1's (the counters share the zeroed 256-byte head with SSCounters)
static inline M3Layout m3_layout(void *temp, int64_t n, int64_t nseg_max) {
    const SSLayout S = ss_layout(temp, n, nseg_max);
    M3Layout M;
    M.cnt = (M3Counters *)((char *)temp + 64);
    M.rng_max = m3_rng_max(n);
    M.tiles_max = ss_tiles_max(n, 0);
    M.batch_max = m3_batch_max(n);
    char *p = (char *)S.ghist + ss_align((size_t)S.nlarge_max * SS_MAX_PASSES * 256 * 4);
    for (int r = 0; r < M3_ROUNDS; r++) {
        M.rng[r] = (M3Range *)p; p += ss_align((size_t)M.rng_max * sizeof(M3Range));
        M.tile_rng[r] = (M3Tile *)p; p += ss_align((size_t)M.tiles_max * sizeof(M3Tile));
        M.hist[r] = (uint32_t *)p; p += ss_align((size_t)M.rng_max * 256 * 4);
    }
    for (int l = 0; l < 2; l++) {
        M.batch[l] = (M3Batch *)p; p += ss_align((size_t)M.batch_max * sizeof(M3Batch));
    }
    return M;
}

struct SSArgs {
    const uint32_t *keys_in;
    const int32_t *vals_in;  // nullptr: value = element index
    uint32_t *keys_out;
    int32_t *vals_out;
    uint32_t *keys_tmp;
    int32_t *vals_tmp;
    const int64_t *off;    // [nseg + 1], device
    const int64_t *dims;   // device: dims[0] = n, dims[1] = nseg (actual values; the grids use upper bounds)
    const int32_t *segid;  // optional: segment of every element (enables the tiny-segment path)
    uint32_t *heads_out;   // optional: bit q set for the first element q of every non-empty segment (zeroed by the caller)
    int tiny_max;          // segments up to this size go through the tiny path (0 without segid)
    int key_bits, n_passes, bits_per_pass;
    SSLayout L;
    int msd;               // 1: large segments go through the generation-3 chain (tdt_segsort3.cuh)
    int m3_byval;          // generation 3: the value is the element index (ties broken by value, unordered partition passes)
    int m3_shift0;         // bit position of the round-0 digit
    M3Layout m3;
    int *err;
};

__device__ __forceinline__ uint32_t ss_digit(uint32_t key, int pass, int bits) {
    return (key >> (pass * bits)) & ((1u << bits) - 1u);
}

// Sorts every segment [off[s], off[s+1]) of keys_in/vals_in by key (stable) into keys_out/vals_out; defined in the
// one translation unit that sets TDT_SEGSORT_IMPL (tdt_cluster.cu), declared for the others (tdt_aggregate.cu).
int segsort_pairs(const uint32_t *keys_in, const int32_t *vals_in, uint32_t *keys_out, int32_t *vals_out,
                  uint32_t *keys_tmp, int32_t *vals_tmp, const int64_t *off, const int64_t *dims, const int32_t *segid,
                  int64_t n_max, int64_t nseg_max, int key_bits, void *temp, size_t temp_bytes, int *err,
                  cudaStream_t st);

// Callers that run several sorts concurrently (parallel branches of one call) give every branch its own forked side
// stream: segsort_set_branch(b), b in [0, SS_BRANCHES), before each segsort_pairs (thread-local, default 0).
constexpr int SS_BRANCHES = 4;
void segsort_set_branch(int b);
// The next segsort_pairs call of this thread also sets, in the zeroed bitmask `heads`, the bit of the first element of
// every non-empty segment (what the range-query kernels read); returns false when the active sort generation cannot
// (the caller then launches its own kernel).
bool segsort_want_heads(uint32_t *heads);

#ifdef TDT_SEGSORT_IMPL
static thread_local int g_ss_branch = 0;
static thread_local cudaEvent_t g_ss_join2_ev[16 * SS_BRANCHES] = {};   // join of the generation-3 side chain
static thread_local int g_ss_join2_idx = 0;
static thread_local uint32_t *g_ss_heads_out = nullptr;   // set by segsort_want_heads for the NEXT segsort_pairs call
void segsort_set_branch(int b) { g_ss_branch = b >= 0 && b < SS_BRANCHES ? b : 0; }

// ---- classification: windows of the small segments, tile ranges of the large ones -----------------
// SPREAD: threads per segment.  A sort over a few hundred large segments (posA inside the pairs) gives every segment's
// lane seven idle neighbours, so that a warp writes the tile tables of 4 ranges one after the other, not of 32.
template <int SPREAD>
__global__ void segsort_classify_kernel(SSArgs a) {
    const int64_t nseg = a.dims[1];
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t s = gid / SPREAD;
    const int lane = threadIdx.x & 31;
    int64_t q = 0, size = 0;
    if (s < nseg && (SPREAD == 1 || gid % SPREAD == 0)) {
        q = a.off[s];
        size = a.off[s + 1] - q;
    }
    {   // the caller's segment-head bitmask (a kernel of its own until r02): one atomic per group of lanes
        const bool has = a.heads_out != nullptr && size > 0;
        const uint32_t hact = __ballot_sync(0xffffffffu, has);
        if (has) {
            const uint32_t peers = __match_any_sync(hact, q >> 5);
            const uint32_t bits = __reduce_or_sync(peers, 1u << (q & 31));
            if ((peers & lanemask_lt()) == 0u) atomicOr(a.heads_out + (q >> 5), bits);
        }
    }
    const bool small = size > a.tiny_max && size <= SS_LOCAL_MAX;
    // consecutive segments mostly start in the same window: one atomic pair per group of lanes
    const uint32_t act = __ballot_sync(0xffffffffu, small);
    if (small) {
        const int64_t k = q / SS_WINDOW;
        const uint32_t peers = __match_any_sync(act, k);
        const uint32_t total = __reduce_add_sync(peers, (uint32_t)size);
        if ((peers & lanemask_lt()) == 0u) {
            atomicMax(a.L.win_lo + k, 0x7fffffff - (int32_t)s);
            atomicAdd(a.L.win_cnt + k, total);
        }
        if ((act & lanemask_lt()) == 0u) atomicAdd(&a.L.cnt->n_small, __popc(act));
        if ((peers >> lane) == 1u) atomicMax(a.L.win_hi + k, (int32_t)s + 1);
    }
    // large segments (rare): the whole warp writes the tile -> segment entries and zeroes the digit histograms of
    // each one its lanes found (this was a kernel of its own, one more launch on the critical path of every sort)
    if (a.msd) {
        // generation 3: a segment that fits one finish batch is queued as such (its keys span the whole key range);
        // larger ones become round-0 ranges (tile -> range entries and zeroed digit counts written by the whole warp)
        if (size > SS_LOCAL_MAX && size <= M3_CAP) {
            const int32_t bi = atomicAdd(&a.m3.cnt->n_batches[0], 1);
            if (bi < a.m3.batch_max) {
                M3Batch B;
                B.start = q;
                B.count = (int32_t)size;
                B.klo = 0u;
                B.sh = a.key_bits > M3_SLOT_BITS ? a.key_bits - M3_SLOT_BITS : 0;
                B.nslots = 1 << (a.key_bits > M3_SLOT_BITS ? M3_SLOT_BITS : a.key_bits);
                B.level = 0;
                B.flags = 0;
                a.m3.batch[0][bi] = B;
            } else {
                atomicMax(a.err, SS_ERR_INTERNAL);
            }
        }
        const bool ranged = size > M3_CAP;
        int32_t ri = -1, rtb = 0, rnt = 0;
        if (ranged) {
            ri = atomicAdd(&a.m3.cnt->n_rng[0], 1);
            rnt = (int32_t)((size + SS_TILE - 1) / SS_TILE);
            rtb = atomicAdd(&a.m3.cnt->n_tiles[0], rnt);
            if (ri < a.m3.rng_max && (int64_t)rtb + rnt <= a.m3.tiles_max) {
                M3Range R;
                R.start = q;
                R.size = (int32_t)size;
                R.tile_base = rtb;
                R.klo = 0u;
                R.shift = a.m3_shift0;
                R.seg = (int32_t)s;
                R.pad = 0;
                a.m3.rng[0][ri] = R;
            } else {
                atomicMax(a.err, SS_ERR_INTERNAL);
                ri = -1;
            }
        }
        u32 rtodo = __ballot_sync(0xffffffffu, ri >= 0);
        while (rtodo) {
            const int src = __ffs(rtodo) - 1;
            rtodo &= rtodo - 1;
            const int32_t i2 = __shfl_sync(0xffffffffu, ri, src), t2 = __shfl_sync(0xffffffffu, rtb, src);
            const int32_t n2 = __shfl_sync(0xffffffffu, rnt, src);
            const int64_t q2 = __shfl_sync(0xffffffffu, q, src), z2 = __shfl_sync(0xffffffffu, size, src);
            for (int32_t t = lane; t < n2; t += 32) {
                M3Tile T;
                T.t0 = q2 + (int64_t)t * SS_TILE;
                T.cnt = (int32_t)(z2 - (int64_t)t * SS_TILE < SS_TILE ? z2 - (int64_t)t * SS_TILE : SS_TILE);
                T.rng = i2;
                a.m3.tile_rng[0][t2 + t] = T;
            }
            uint32_t *h = a.m3.hist[0] + (size_t)i2 * 256;
            for (int i = lane; i < 256; i += 32) h[i] = 0u;
        }
        return;
    }
    const bool large = size > SS_LOCAL_MAX;
    int32_t idx = 0, tb = 0, nt = 0;
    if (large) {
        idx = atomicAdd(&a.L.cnt->n_large, 1);
        nt = (int32_t)((size + SS_TILE - 1) / SS_TILE);
        tb = atomicAdd(&a.L.cnt->n_tiles, nt);
        SSLarge rec;
        rec.start = q;
        rec.size = size;
        rec.tile_base = tb;
        rec.pad = 0;
        a.L.large[idx] = rec;
    }
    u32 todo = __ballot_sync(0xffffffffu, large);
    while (todo) {
        const int src = __ffs(todo) - 1;
        todo &= todo - 1;
        const int32_t i2 = __shfl_sync(0xffffffffu, idx, src), t2 = __shfl_sync(0xffffffffu, tb, src);
        const int32_t n2 = __shfl_sync(0xffffffffu, nt, src);
        for (int32_t t = lane; t < n2; t += 32) a.L.tile_seg[t2 + t] = i2;
        uint32_t *h = a.L.ghist + (size_t)i2 * SS_MAX_PASSES * 256;
        for (int i = lane; i < SS_MAX_PASSES * 256; i += 32) h[i] = 0u;
    }
}

#ifndef TDT_SS_TINY_SPLIT
#define TDT_SS_TINY_SPLIT 0
#endif
#ifndef TDT_SS_TINY_MAX
#define TDT_SS_TINY_MAX 128   // measured on B200 (30X set, posB sort): 32 -> 0.338 ms, 64 -> 0.296, 128 -> 0.275, 256 -> 0.275
#endif
// ---- tiny segments: rank by counting ------------------------------------------------------------------
// (one element per thread and iteration: 2 / 4 consecutive elements per thread with 16-byte loads measured slower,
//  sort_y 0.445 / 0.450 ms vs 0.432 ms -- fewer resident chains outweigh the batched prologue)
__global__ void __launch_bounds__(256) segsort_tiny_kernel(SSArgs a) {
    const int64_t n = a.dims[0];
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int64_t g = a.segid[j];
        const int64_t s0 = a.off[g], s1 = a.off[g + 1];
        if (s1 - s0 > a.tiny_max) continue;
        const uint32_t key = a.keys_in[j];
        if (a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
        int rank = 0;
#if TDT_SS_TINY_SPLIT
        // members before j precede it on "<=", members after it on "<": one compare per member, four loads in flight
        const uint32_t *kb = a.keys_in + s0;
        const int nb = (int)(j - s0), na = (int)(s1 - j) - 1;
        const uint32_t *ka = a.keys_in + j + 1;
        int i = 0;
        for (; i + 4 <= nb; i += 4) {
            const uint32_t x0 = kb[i], x1 = kb[i + 1], x2 = kb[i + 2], x3 = kb[i + 3];
            rank += (x0 <= key) + (x1 <= key) + (x2 <= key) + (x3 <= key);
        }
        for (; i < nb; i++) rank += kb[i] <= key;
        i = 0;
        for (; i + 4 <= na; i += 4) {
            const uint32_t x0 = ka[i], x1 = ka[i + 1], x2 = ka[i + 2], x3 = ka[i + 3];
            rank += (x0 < key) + (x1 < key) + (x2 < key) + (x3 < key);
        }
        for (; i < na; i++) rank += ka[i] < key;
#else
        for (int64_t i = s0; i < s1; i++) {
            const uint32_t other = a.keys_in[i];
            rank += (other < key) || (other == key && i < j);
        }
#endif
        a.keys_out[s0 + rank] = key;
        a.vals_out[s0 + rank] = a.vals_in ? a.vals_in[j] : (int32_t)j;
    }
}

// ---- stable ranking primitives ------------------------------------------------------------------------
// lanes holding the same digit: every lane ORs its bit into the warp's mask word of that digit
// (mm: this warp's 256 words, all zero on entry and on exit)
#ifndef TDT_SS_MATCH_BALLOT
#define TDT_SS_MATCH_BALLOT 0   // measured on B200 (r01, posA sort of the 30X set): ballots 0.74 ms, atomicOr 0.65 ms
                                // (later tree: atomicOr 0.60 ms, __match_any_sync / SASS MATCH.ANY 0.86 ms)
#endif
constexpr int SS_MM = TDT_SS_MATCH_BALLOT ? 0 : 1;  // per-warp match words only for the atomicOr variant
__device__ __forceinline__ uint32_t ss_match(uint32_t *mm, uint32_t d, bool valid, int bits) {
#if TDT_SS_MATCH_BALLOT
    // one ballot per digit bit, intersected: no shared memory (an ATOMS.OR costs ~2 cycles per lane of the LSU pipe,
    // which made the L1 pipe the limiter of the sort passes)
    uint32_t peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 8; b++) {
        if (b < bits) {
            const bool bit = (d >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, bit);
            peers &= bit ? bal : ~bal;
        }
    }
    return valid ? peers : 0u;
#else
    const int lane = threadIdx.x & 31;
    if (valid) atomicOr(&mm[d], 1u << lane);
    __syncwarp();
    const uint32_t peers = valid ? mm[d] : 0u;
    __syncwarp();
    if (valid && (peers & lanemask_lt()) == 0u) mm[d] = 0u;
    return peers;
#endif
}

// info word per element: the lanes of the warp's 32-element chunk that hold the same digit (0 = no element);
// rank among them = popc(info & lanes below), group size = popc(info), group leader = lowest set lane.
// TDT_SS_ATOMIC_RANK (default): the leader adds the group to the warp-private counter with ONE shared-memory
// atomic in each phase and, when scattering, hands the old value to its group by shuffle -- instead of a read
// by every lane plus a write by the leader (two bank-conflicted shared accesses per element and phase).
// Measured on B200 (r01, posA sort of the 30X set): 0.601 ms with, 0.606 ms without -- within noise; the same for
// the 64-bit (key, value) staging below (TDT_SS_KV64).  Fewer shared-memory wavefronts did not move the pass:
// with 24 resident warps per SM it waits on latencies (scoreboard and barrier stalls), not on a saturated pipe.
#ifndef TDT_SS_ATOMIC_RANK
#define TDT_SS_ATOMIC_RANK 1
#endif

// TDT_SS_FUSED_RANK (default, r02): the count phase already yields every element's rank inside its warp -- the group
// leader's atomicAdd on the warp-private counter RETURNS the number of equal digits the warp has seen before this chunk,
// handed to the group by shuffle -- so the scatter phase is a plain shared-memory load of the (warp, digit) base instead
// of a second round of leader atomics: two shared-memory atomics per element and pass (match + count) instead of three.
// The passes do not speed up with more resident CTAs (r01: 3, 4 or 6 per SM give the same time), i.e. they are bound
// by the shared-memory pipe these atomics go through.
#ifndef TDT_SS_FUSED_RANK
#define TDT_SS_FUSED_RANK 1
#endif
constexpr uint32_t SS_NO_ELEM = 0xffffffffu;

// count phase: warp-private digit histograms of the warp's own run of the tile.
// info[c]: FUSED_RANK: the element's rank among the warp's elements with the same digit (SS_NO_ELEM = no element);
//          otherwise the match result (lanes of the chunk holding the same digit, 0 = no element).
template <int CHUNKS, typename DigitFn>
__device__ __forceinline__ void ss_count(int count, int epw, uint32_t *wh, uint32_t *mm, uint32_t (&info)[CHUNKS],
                                         int bits, DigitFn digit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
        info[c] = TDT_SS_FUSED_RANK ? SS_NO_ELEM : 0u;
        if (c * 32 < epw) {
            const int e = warp * epw + c * 32 + lane;
            const bool valid = e < count;
            const uint32_t d = valid ? digit(e, c) : 0u;
            const uint32_t peers = ss_match(mm, d, valid, bits);
#if TDT_SS_FUSED_RANK
            const uint32_t rank = __popc(peers & lanemask_lt());
            uint32_t before = 0;
            if (valid && rank == 0u) before = atomicAdd(&wh[d], (uint32_t)__popc(peers));
            before = __shfl_sync(0xffffffffu, before, valid ? (__ffs(peers) - 1) : lane);
            if (valid) info[c] = before + rank;
#else
            if (valid) {
                info[c] = peers;
                if ((peers & lanemask_lt()) == 0u) {
#if TDT_SS_ATOMIC_RANK
                    atomicAdd(&wh[d], (uint32_t)__popc(peers));
#else
                    wh[d] += __popc(peers);
#endif
                }
            }
            __syncwarp();
#endif
        }
    }
}

// scatter phase: stable position of every element; wh[d] holds the base of (this warp, digit).
// digit(e, c) must return what it returned in the count phase.
template <int CHUNKS, typename DigitFn, typename EmitFn>
__device__ __forceinline__ void ss_scatter(int epw, uint32_t *wh, const uint32_t (&info)[CHUNKS], DigitFn digit,
                                           EmitFn emit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < CHUNKS; c++) {
        if (c * 32 < epw) {
#if TDT_SS_FUSED_RANK
            const int e = warp * epw + c * 32 + lane;
            if (info[c] != SS_NO_ELEM) emit(e, c, wh[digit(e, c)] + info[c]);
#else
            const uint32_t peers = info[c];
            const bool valid = peers != 0u;
            const int e = warp * epw + c * 32 + lane;
            const uint32_t d = valid ? digit(e, c) : 0u;
            const uint32_t rank = __popc(peers & lanemask_lt()), group = __popc(peers);
            uint32_t base = 0;
#if TDT_SS_ATOMIC_RANK
            if (valid && rank == 0u) base = atomicAdd(&wh[d], group);
            base = __shfl_sync(0xffffffffu, base, valid ? (__ffs(peers) - 1) : lane);
#else
            if (valid) base = wh[d];
            __syncwarp();
            if (valid && rank == 0u) wh[d] = base + group;
            __syncwarp();
#endif
            if (valid) emit(e, c, base + rank);
#endif
        }
    }
}

// exclusive scan of one value per thread over the first 256 threads (thread = digit); all threads of the CTA call it
template <int NWARPS>
__device__ __forceinline__ uint32_t ss_block_excl_scan_256(uint32_t v, uint32_t *scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31 && warp < 8) scratch[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < 8; w++)
        if (w < warp) wbase += scratch[w];
    __syncthreads();
    return wbase + inc - v;
}

// turns the per-warp counts wh[w][d] into running bases: (#elements with a smaller digit) + (#elements with
// digit d in earlier warps); returns this thread's digit total and exclusive digit base (threads >= 256: 0)
template <int NWARPS>
__device__ __forceinline__ void ss_digit_bases(uint32_t (*wh)[256], uint32_t *scratch, uint32_t &total,
                                               uint32_t &excl) {
    const int d = threadIdx.x;
    uint32_t run = 0;
    if (d < 256) {
#pragma unroll
        for (int w = 0; w < NWARPS; w++) {
            const uint32_t t = wh[w][d];
            wh[w][d] = run;
            run += t;
        }
    }
    total = run;
    excl = ss_block_excl_scan_256<NWARPS>(run, scratch);
    if (d < 256) {
#pragma unroll
        for (int w = 0; w < NWARPS; w++) wh[w][d] += excl;
    }
}

// ---- small segments: gather, sort in shared memory, put back ----------------------------------------------
// A CTA owns a contiguous span of element windows and walks it in BATCHES of consecutive windows whose small
// segments together fill (at most) the shared-memory capacity, so the per-batch fixed costs are amortised.
constexpr int SS_SCAN_WINDOWS = 512;  // windows examined per planning round
constexpr int SS_BATCH_WINDOWS = 60;  // a batch spans < 2^16 element positions (16-bit relative starts)
constexpr size_t SS_LOCAL_SMEM = (size_t)SS_LOCAL_CAP * (4 + 4 + 2) * 2 + (size_t)SS_LWARPS * 256 * 4 * (1 + SS_MM) + 256 * 4 +
                                 (size_t)(SS_LOCAL_CAP + 2) * 2 * 2 + SS_SCAN_WINDOWS * 4 + 64;

__global__ void __launch_bounds__(SS_LTHREADS) segsort_local_kernel(SSArgs a) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned char *p = ss_smem;
    uint32_t *K[2];
    int32_t *V[2];
    uint16_t *Lid[2];
    K[0] = (uint32_t *)p; p += SS_LOCAL_CAP * 4;
    K[1] = (uint32_t *)p; p += SS_LOCAL_CAP * 4;
    V[0] = (int32_t *)p; p += SS_LOCAL_CAP * 4;
    V[1] = (int32_t *)p; p += SS_LOCAL_CAP * 4;
    Lid[0] = (uint16_t *)p; p += SS_LOCAL_CAP * 2;
    Lid[1] = (uint16_t *)p; p += SS_LOCAL_CAP * 2;
    uint32_t(*wh)[256] = (uint32_t(*)[256])p; p += SS_LWARPS * 256 * 4;
    uint32_t(*mm)[256] = (uint32_t(*)[256])p; p += SS_MM * SS_LWARPS * 256 * 4;
    uint32_t *bin = (uint32_t *)p; p += 256 * 4;
    uint32_t *wcnt = (uint32_t *)p; p += SS_SCAN_WINDOWS * 4;
    uint16_t *sc_start = (uint16_t *)p; p += (SS_LOCAL_CAP + 2) * 2;  // compact start of every gathered segment
    uint16_t *sg_rel = (uint16_t *)p;                                 // its global start, relative to the batch
    __shared__ uint32_t s_warp[SS_LWARPS];
    __shared__ uint32_t s_carry;
    __shared__ int s_batch[2];  // next batch: [first window, one past the last window) relative to the round

    if (a.L.cnt->n_small == 0) return;   // the posA sort of whole pairs: nothing here, leave the SMs to the large chain
    const int64_t n = a.dims[0];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (SS_MM)
        for (int i = threadIdx.x; i < SS_LWARPS * 256; i += SS_LTHREADS) (&mm[0][0])[i] = 0u;

    const int64_t nwin = (n + SS_WINDOW - 1) / SS_WINDOW;
    const int64_t span = (nwin + gridDim.x - 1) / gridDim.x;
    const int64_t span_lo = (int64_t)blockIdx.x * span;
    const int64_t span_hi = span_lo + span < nwin ? span_lo + span : nwin;

    for (int64_t round = span_lo; round < span_hi; round += SS_SCAN_WINDOWS) {
        const int nround = (int)(span_hi - round < SS_SCAN_WINDOWS ? span_hi - round : SS_SCAN_WINDOWS);
        __syncthreads();
        if (threadIdx.x < nround) wcnt[threadIdx.x] = a.L.win_cnt[round + threadIdx.x];
        int next = 0;
        while (true) {
            __syncthreads();
            if (threadIdx.x == 0) {  // plan: skip empty windows, then take windows while they fit
                int w = next;
                while (w < nround && wcnt[w] == 0u) w++;
                int e = w;
                uint32_t sum = 0;
                while (e < nround && e - w < SS_BATCH_WINDOWS && (wcnt[e] == 0u || sum + wcnt[e] <= SS_LOCAL_CAP)) {
                    sum += wcnt[e];
                    e++;
                }
                while (e > w && wcnt[e - 1] == 0u) e--;  // end on a populated window
                s_batch[0] = w;
                s_batch[1] = e;
                s_carry = 0u;
            }
            __syncthreads();
            const int bw0 = s_batch[0], bw1 = s_batch[1];
            if (bw0 >= nround) break;
            next = bw1;
            const int64_t k0 = round + bw0, k1 = round + bw1;  // windows [k0, k1); both ends populated
            const int32_t lo_s = 0x7fffffff - a.L.win_lo[k0], hi_s = a.L.win_hi[k1 - 1] - 1;
            const int64_t wbase = k0 * SS_WINDOW;
            // ---- A: the small segments of the batch, their compact layout (size sum << 16 | count) ----
            for (int64_t s0 = lo_s; s0 <= hi_s; s0 += SS_LTHREADS) {
                const int64_t s = s0 + threadIdx.x;
                uint32_t size = 0;
                int64_t q = 0;
                if (s <= hi_s) {
                    q = a.off[s];
                    const int64_t sz = a.off[s + 1] - q;
                    if (sz > a.tiny_max && sz <= SS_LOCAL_MAX) size = (uint32_t)sz;
                }
                const uint32_t v = size ? ((size << 16) | 1u) : 0u;
                uint32_t inc = v;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) s_warp[warp] = inc;
                __syncthreads();
                uint32_t before = s_carry;
                for (int w = 0; w < warp; w++) before += s_warp[w];
                const uint32_t ex = before + inc - v;
                if (size) {
                    sc_start[ex & 0xffffu] = (uint16_t)(ex >> 16);
                    sg_rel[ex & 0xffffu] = (uint16_t)(q - wbase);
                }
                __syncthreads();
                if (threadIdx.x == SS_LTHREADS - 1) s_carry = before + inc;
                __syncthreads();
            }
            const uint32_t tot = s_carry;
            const int nsegs = (int)(tot & 0xffffu), count = (int)(tot >> 16);
            if (threadIdx.x == 0) sc_start[nsegs] = (uint16_t)count;
            __syncthreads();
            // ---- B: gather -------------------------------------------------------------------------------
            for (int c = threadIdx.x; c < count; c += SS_LTHREADS) {
                int lo = 0, hi = nsegs;  // segment of compact position c: largest i with sc_start[i] <= c
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if ((int)sc_start[mid] <= c) lo = mid; else hi = mid;
                }
                const int64_t g = wbase + sg_rel[lo] + (c - (int)sc_start[lo]);
                const uint32_t key = a.keys_in[g];
                if (a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
                K[0][c] = key;
                V[0][c] = a.vals_in ? a.vals_in[g] : (int32_t)g;
                Lid[0][c] = (uint16_t)lo;
            }
            // ---- C: LSD passes over the composite (segment index, key), 8 bits at a time -----------------
            const int lid_bits = nsegs > 1 ? (32 - __clz(nsegs - 1)) : 0;
            const int passes = (a.key_bits + lid_bits + 7) / 8;
            const int epw = ((count + SS_LTHREADS - 1) / SS_LTHREADS) * 32;  // elements per warp, multiple of 32
            int cur = 0;
            for (int pass = 0; pass < passes; pass++) {
                for (int i = threadIdx.x; i < SS_LWARPS * 256; i += SS_LTHREADS) (&wh[0][0])[i] = 0u;
                __syncthreads();
                const uint32_t *Kc = K[cur];
                const int32_t *Vc = V[cur];
                const uint16_t *Lc = Lid[cur];
                const int kb = a.key_bits, sh = 8 * pass;
                uint32_t info[SS_LCHUNKS];
                ss_count<SS_LCHUNKS>(count, epw, wh[warp], mm[warp], info, 8, [&](int e, int) -> uint32_t {
                    const unsigned long long comp = ((unsigned long long)Lc[e] << kb) | (unsigned long long)Kc[e];
                    return (uint32_t)(comp >> sh) & 255u;
                });
                __syncthreads();
                uint32_t total, excl;
                ss_digit_bases<SS_LWARPS>(wh, bin, total, excl);
                __syncthreads();
                uint32_t *Ka = K[cur ^ 1];
                int32_t *Va = V[cur ^ 1];
                uint16_t *La = Lid[cur ^ 1];
                ss_scatter<SS_LCHUNKS>(epw, wh[warp], info, [&](int e, int) -> uint32_t {
                    const unsigned long long comp = ((unsigned long long)Lc[e] << kb) | (unsigned long long)Kc[e];
                    return (uint32_t)(comp >> sh) & 255u;
                }, [&](int e, int, uint32_t pos) {
                    Ka[pos] = Kc[e];
                    Va[pos] = Vc[e];
                    La[pos] = Lc[e];
                });
                __syncthreads();
                cur ^= 1;
            }
            // ---- D: put back: sorted by (segment, key), so compact position c is rank c - start in its segment
            for (int c = threadIdx.x; c < count; c += SS_LTHREADS) {
                const int lid = Lid[cur][c];
                const int64_t g = wbase + sg_rel[lid] + (c - (int)sc_start[lid]);
                a.keys_out[g] = K[cur][c];
                a.vals_out[g] = V[cur][c];
            }
        }
    }
}

// ---- large segments --------------------------------------------------------------------------------
__global__ void __launch_bounds__(SS_THREADS) segsort_hist_kernel(SSArgs a) {
    __shared__ uint32_t h[SS_MAX_PASSES][256];
    const int n_tiles = a.L.cnt->n_tiles;
    // the grid is capped at the resident-CTA count (ss_grid): a CTA walks tiles blockIdx.x, +gridDim.x, ...
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < SS_MAX_PASSES * 256; i += SS_THREADS) (&h[0][0])[i] = 0u;
        __syncthreads();
        const int seg = a.L.tile_seg[tile];
        const SSLarge L = a.L.large[seg];
        const int64_t t0 = L.start + (int64_t)(tile - L.tile_base) * SS_TILE;
        const int64_t rem = L.start + L.size - t0;
        const int cnt = rem < SS_TILE ? (int)rem : SS_TILE;
        for (int e = threadIdx.x; e < cnt; e += SS_THREADS) {
            const uint32_t key = a.keys_in[t0 + e];
            if (a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
            for (int p = 0; p < a.n_passes; p++) atomicAdd(&h[p][ss_digit(key, p, a.bits_per_pass)], 1u);
        }
        __syncthreads();
        uint32_t *g = a.L.ghist + (size_t)seg * SS_MAX_PASSES * 256;
        for (int i = threadIdx.x; i < a.n_passes * 256; i += SS_THREADS) {
            const uint32_t v = (&h[0][0])[i];
            if (v) atomicAdd(g + i, v);
        }
        __syncthreads();
    }
}

// exclusive scan over the 256 digits of every (large segment, pass): one warp each
__global__ void segsort_scan_hist_kernel(SSArgs a) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int64_t seg = wid / SS_MAX_PASSES;
    const int pass = (int)(wid % SS_MAX_PASSES);
    if (seg >= a.L.cnt->n_large || pass >= a.n_passes) return;
    uint32_t *g = a.L.ghist + ((size_t)seg * SS_MAX_PASSES + pass) * 256 + lane * 8;
    uint32_t v[8], t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = g[i];
        t += v[i];
    }
    uint32_t inc = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    uint32_t ex = inc - t;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        g[i] = ex;
        ex += v[i];
    }
}

// look-back word: [31:30] flag (1 = tile total, 2 = inclusive prefix), [29:26] epoch (pass + 1), [25:0] count
__device__ __forceinline__ uint32_t ss_pack(uint32_t flag, uint32_t epoch, uint32_t count) {
    return (flag << 30) | (epoch << 26) | count;
}
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

#ifndef TDT_SS_KV64
#define TDT_SS_KV64 1
#endif
constexpr size_t SS_PASS_SMEM = (size_t)SS_TILE * 4 * 2 + (size_t)SS_WARPS * 256 * 4 * (1 + SS_MM) + 256 * 4 + 256 * 8 + 64;

// One pass over one tile.  Tiles are taken in blockIdx order (the chained scan waits only on tiles with a smaller
// index, which the hardware has already scheduled -- the same forward-progress assumption as CUB's DeviceScan).
#ifndef TDT_SS_LB
#define TDT_SS_LB 4
#endif
#ifndef TDT_SS_PASS_MINBLOCKS
#define TDT_SS_PASS_MINBLOCKS 3
#endif
__global__ void __launch_bounds__(SS_THREADS, TDT_SS_PASS_MINBLOCKS) segsort_pass_kernel(SSArgs a, int pass, const uint32_t *src_k,
                                                                  const int32_t *src_v, uint32_t *dst_k,
                                                                  int32_t *dst_v) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned char *p = ss_smem;
#if TDT_SS_KV64
    uint2 *KV = (uint2 *)p; p += SS_TILE * 8;   // (key, value) staged as one 64-bit word: one scattered store, not two
#else
    uint32_t *K2 = (uint32_t *)p; p += SS_TILE * 4;
    int32_t *V2 = (int32_t *)p; p += SS_TILE * 4;
#endif
    uint32_t(*wh)[256] = (uint32_t(*)[256])p; p += SS_WARPS * 256 * 4;
    uint32_t(*mm)[256] = (uint32_t(*)[256])p; p += SS_MM * SS_WARPS * 256 * 4;
    uint32_t *bin = (uint32_t *)p; p += 256 * 4;
    int64_t *gbase = (int64_t *)p;

    const int n_tiles = a.L.cnt->n_tiles;
    // Persistent CTAs: the grid is capped at the number of CTAs the device keeps resident (ss_grid), CTA b takes tiles
    // b, b + gridDim.x, ...  A tile only waits on tiles with a smaller index; those belong to CTAs that are resident
    // too and are at the same or an earlier round, so the chained scan always makes progress -- and a sort with few
    // (or no) large segments no longer pays for thousands of empty CTAs.
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    for (int i = threadIdx.x; i < SS_WARPS * 256; i += SS_THREADS) {
        (&wh[0][0])[i] = 0u;
        if (SS_MM) (&mm[0][0])[i] = 0u;
    }
    const int seg = a.L.tile_seg[tile];
    const SSLarge L = a.L.large[seg];
    const int lt = tile - L.tile_base;
    const int64_t t0 = L.start + (int64_t)lt * SS_TILE;
    const int64_t rem = L.start + L.size - t0;
    const int cnt = rem < SS_TILE ? (int)rem : SS_TILE;
    const int bits = a.bits_per_pass;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int epw = SS_TILE / SS_WARPS;  // every warp owns 512 consecutive elements of the tile

    uint32_t key[SS_CHUNKS];
    int32_t val[SS_CHUNKS];
#pragma unroll
    for (int c = 0; c < SS_CHUNKS; c++) {
        const int e = warp * epw + c * 32 + lane;
        key[c] = 0u;
        val[c] = 0;
        if (e < cnt) {
            key[c] = src_k[t0 + e];
            val[c] = src_v ? src_v[t0 + e] : (int32_t)(t0 + e);
        }
    }
    __syncthreads();
    uint32_t info[SS_CHUNKS];
    ss_count<SS_CHUNKS>(cnt, epw, wh[warp], mm[warp], info, bits,
                        [&](int, int c) -> uint32_t { return ss_digit(key[c], pass, bits); });
    __syncthreads();
    uint32_t total, excl;
    ss_digit_bases<SS_WARPS>(wh, bin, total, excl);
    // per-digit chained scan over the earlier tiles of this segment (thread = digit), in two halves: the tile's
    // digit totals are published first, the walk over the predecessors happens AFTER the shared-memory scatter,
    // when more of them hold a complete prefix (the walk was ~30 % of the pass when it came first)
    const uint32_t epoch = (uint32_t)pass + 1u;
    uint32_t *row = a.L.status + (size_t)tile * 256 + threadIdx.x;
    const bool live = threadIdx.x < (1u << bits);  // digits that exist in this pass
    uint32_t gh = 0;
    if (live) {
        st_volatile_u32(row, ss_pack(lt == 0 ? 2u : 1u, epoch, total));
        gh = a.L.ghist[((size_t)seg * SS_MAX_PASSES + pass) * 256 + threadIdx.x];
    }
    __syncthreads();  // wh holds every warp's digit bases
    ss_scatter<SS_CHUNKS>(epw, wh[warp], info, [&](int, int c) -> uint32_t { return ss_digit(key[c], pass, bits); },
                          [&](int, int c, uint32_t pos) {
#if TDT_SS_KV64
                              KV[pos] = make_uint2(key[c], (uint32_t)val[c]);
#else
                              K2[pos] = key[c];
                              V2[pos] = val[c];
#endif
                          });
    {
        uint32_t before = 0;
        if (lt != 0 && live) {
            constexpr int LB = TDT_SS_LB;  // predecessors inspected per round trip
            const uint32_t *prow = row - 256;
            int left = lt;  // tiles of this segment before this one
            bool done = false;
            while (!done) {
                uint32_t sv[LB];
#pragma unroll
                for (int i = 0; i < LB; i++) sv[i] = i < left ? ld_volatile_u32(prow - (size_t)i * 256) : 0u;
#pragma unroll
                for (int i = 0; i < LB; i++) {
                    if (!done && i < left) {
                        uint32_t s = sv[i];
                        while ((s >> 30) == 0u || ((s >> 26) & 15u) != epoch) s = ld_volatile_u32(prow - (size_t)i * 256);
                        before += s & 0x3ffffffu;
                        done = (s >> 30) == 2u;
                    }
                }
                prow -= (size_t)LB * 256;
                left -= LB;
            }
            st_volatile_u32(row, ss_pack(2u, epoch, before + total));
        }
        if (threadIdx.x < 256) gbase[threadIdx.x] = L.start + (int64_t)gh + (int64_t)before - (int64_t)excl;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += SS_THREADS) {
#if TDT_SS_KV64
        const uint2 kv = KV[i];
        const uint32_t k = kv.x;
        const int32_t v = (int32_t)kv.y;
#else
        const uint32_t k = K2[i];
        const int32_t v = V2[i];
#endif
        const int64_t g = gbase[ss_digit(k, pass, bits)] + i;
        dst_k[g] = k;
        dst_v[g] = v;
    }
    __syncthreads();   // K2 / V2 / gbase are rewritten by the next tile
    }
}

}  // namespace tdt
#include "tdt_segsort3.cuh"
namespace tdt {

// Which large-segment chain.  Default: generation 3 (MSD rounds + shared-memory finish, tdt_segsort3.cuh) for sorts
// whose value is the element index (vals_in == nullptr: the posA sort of all pairs, the aggregation's sub-sorts) --
// measured on B200 (30X set): 0.61 -> 0.39 ms for posA -- and the four stable LSD passes otherwise: generation 3's
// stable mode pays the warp-private ranking in its pass and scattered stores in its finish, and the sorts with explicit
// values have few large segments, most of them pile-ups that need the later rounds (forced for every sort, final tree:
// posB inside x-runs 0.354 vs 0.275 ms, grouping by candidate 0.570 vs 0.342 ms).
// TDT_SEGSORT=lsd / msd in the environment (read at every call) forces one chain for every sort (A/B runs, tests).
static inline bool ss_use_msd(bool value_is_index) {
    const char *e = getenv("TDT_SEGSORT");
    if (e && e[0] == 'l') return false;
    if (e && e[0] == 'm') return true;
    return value_is_index;
}

#ifdef TDT_M3_DEBUG
#define TDT_M3_DEBUG_DUMP(stream, what)                                                                                 \
    do {                                                                                                                \
        unsigned long long h[16];                                                                                       \
        M3Counters hc;                                                                                                  \
        cudaStreamSynchronize(stream);                                                                                  \
        cudaMemcpyFromSymbol(h, g_m3_dbg, sizeof(h));                                                                   \
        cudaMemcpy(&hc, a.m3.cnt, sizeof(hc), cudaMemcpyDeviceToHost);                                                  \
        fprintf(stderr, "m3 %s byval=%d: ranges %d/%d/%d/%d tiles %d/%d/%d/%d batches %d/%d | finished %llu (hot %llu "  \
                "normal %llu) cycles hot %llu normal %llu max %llu elements %llu sumsq %llu levels %llu/%llu/%llu/%llu/%llu\n", \
                what, (int)byval, hc.n_rng[0], hc.n_rng[1], hc.n_rng[2], hc.n_rng[3], hc.n_tiles[0], hc.n_tiles[1],     \
                hc.n_tiles[2], hc.n_tiles[3], hc.n_batches[0], hc.n_batches[1], h[0], h[1], h[2], h[3], h[4], h[5],      \
                h[6], h[7], h[8], h[9], h[10], h[11], h[12]);                                                           \
        memset(h, 0, sizeof(h));                                                                                        \
        cudaMemcpyToSymbol(g_m3_dbg, h, sizeof(h));                                                                     \
    } while (0)
#else
#define TDT_M3_DEBUG_DUMP(stream, what) do { } while (0)
#endif

// ---- host launcher ---------------------------------------------------------------------------------
// Sorts every segment [off[s], off[s+1]) of keys_in/vals_in by key (stable) into keys_out/vals_out.
// n_max / nseg_max: host-side upper bounds that size the grids; the actual n / nseg are read on the device
// from dims[0] / dims[1].  keys_tmp / vals_tmp: scratch of n_max elements (only touched for large segments).
// segid (optional): segment index of every element; enables the counting path for segments of <= 32 elements.
int segsort1_pairs(const uint32_t *keys_in, const int32_t *vals_in, uint32_t *keys_out, int32_t *vals_out,
                         uint32_t *keys_tmp, int32_t *vals_tmp, const int64_t *off, const int64_t *dims,
                         const int32_t *segid, int64_t n_max, int64_t nseg_max, int key_bits, void *temp,
                         size_t temp_bytes, int *err, cudaStream_t st) {
    uint32_t *heads_out = g_ss_heads_out;
    g_ss_heads_out = nullptr;
    if (n_max <= 0 || nseg_max <= 0) return TDT_OK;
    if (temp_bytes < segsort1_temp_bytes(n_max, nseg_max))
        return fail(TDT_E_WORKSPACE, "segmented sort needs %zu bytes of temporary storage, %zu reserved",
                    segsort1_temp_bytes(n_max, nseg_max), temp_bytes);
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    SSArgs a;
    a.keys_in = keys_in;
    a.vals_in = vals_in;
    a.keys_out = keys_out;
    a.vals_out = vals_out;
    a.keys_tmp = keys_tmp;
    a.vals_tmp = vals_tmp;
    a.off = off;
    a.dims = dims;
    a.segid = segid;
    a.heads_out = heads_out;
    a.tiny_max = segid ? TDT_SS_TINY_MAX : 0;
    a.key_bits = key_bits;
    a.n_passes = (key_bits + 7) / 8;
    a.bits_per_pass = (key_bits + a.n_passes - 1) / a.n_passes;
    a.L = ss_layout(temp, n_max, nseg_max);
    a.msd = ss_use_msd(vals_in == nullptr) ? 1 : 0;
    a.m3_shift0 = key_bits > 8 ? key_bits - 8 : 0;
    a.m3_byval = vals_in == nullptr ? 1 : 0;
    a.m3 = m3_layout(temp, n_max, nseg_max);
    a.err = err;
    TDT_CUDA(cudaMemsetAsync(temp, 0, a.L.zero_bytes, st));
    if (nseg_max <= 8192) TDT_LAUNCH(segsort_classify_kernel<8>, (unsigned)((nseg_max * 8 + 255) / 256), 256, 0, st, a);
    else TDT_LAUNCH(segsort_classify_kernel<1>, (unsigned)((nseg_max + 255) / 256), 256, 0, st, a);
    static thread_local bool configured = false;
    if (!configured) {
        TDT_CUDA(cudaFuncSetAttribute(m3_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_PASS_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(m3_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)M3_UPASS_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(m3_finish_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)M3_FIN_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(m3_finish_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)M3_FIN_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(segsort_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SS_LOCAL_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(segsort_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SS_PASS_SMEM));
        configured = true;
    }
    // The tiny / small-segment kernels and the large-segment chain (hist, scan, passes) touch disjoint segments, and
    // neither fills the machine on its own (both wait on latencies): they run as two branches -- a side stream forked
    // here and joined after the last pass; inside a CUDA-graph capture the fork / join become parallel graph branches.
#ifndef TDT_SS_FORK
#define TDT_SS_FORK 1
#endif
    cudaStream_t side = st;
#if TDT_SS_FORK
    static thread_local cudaStream_t side_streams[16 * SS_BRANCHES] = {};
    static thread_local cudaEvent_t fork_ev[16 * SS_BRANCHES] = {}, join_ev[16 * SS_BRANCHES] = {};
    int dev_id = 0;
    TDT_CUDA(cudaGetDevice(&dev_id));
    if (dev_id >= 0 && dev_id < 16) {
        dev_id = dev_id * SS_BRANCHES + g_ss_branch;   // one side stream per (device, branch)
        if (!side_streams[dev_id]) {
            // branches >= 1 (the aggregation's sub-sorts) run next to a kernel that fills the machine (agg_direct_kernel):
            // their side streams get the caller's high priority, or the small-segment kernels would queue behind it
            int pr_least = 0, pr_greatest = 0;
            TDT_CUDA(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
            TDT_CUDA(cudaStreamCreateWithPriority(&side_streams[dev_id], cudaStreamNonBlocking,
                                                  g_ss_branch > 0 ? pr_greatest : 0));
            TDT_CUDA(cudaEventCreateWithFlags(&fork_ev[dev_id], cudaEventDisableTiming));
            TDT_CUDA(cudaEventCreateWithFlags(&join_ev[dev_id], cudaEventDisableTiming));
        }
        side = side_streams[dev_id];
        TDT_CUDA(cudaEventRecord(fork_ev[dev_id], st));
        TDT_CUDA(cudaStreamWaitEvent(side, fork_ev[dev_id], 0));
    }
#endif
    // (The tiny-segment kernel on a branch of its own: 1.277 vs 1.268 ms for the 30X step in r01, 3.906 vs 3.944 ms on
    // the tumour set in r02 -- within noise.  Started BEFORE the classify kernel, which it does not need: its 4736 CTAs
    // fill the SMs first and hold back classify and everything behind it, posB sort 0.274 -> 0.303 ms.)
    if (segid) {
        int64_t blocks = (n_max + 255) / 256;
        if (blocks > 148 * 32) blocks = 148 * 32;
        TDT_LAUNCH(segsort_tiny_kernel, (unsigned)blocks, 256, 0, side, a);
    }
    int64_t nwin = (n_max + SS_WINDOW - 1) / SS_WINDOW;
    if (nwin > 148 * 4) nwin = 148 * 4;
    TDT_LAUNCH(segsort_local_kernel, (unsigned)nwin, SS_LTHREADS, SS_LOCAL_SMEM, side, a);
    static thread_local int pass_cap = 0, hist_cap = 0, upass_cap = 0, sm_count = 0;   // resident CTAs of the two tile kernels on this device
    if (!pass_cap) {
        int dev = 0, sms = 0, per_sm = 0;
        TDT_CUDA(cudaGetDevice(&dev));
        TDT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        TDT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, segsort_pass_kernel, SS_THREADS, SS_PASS_SMEM));
        pass_cap = sms * (per_sm > 0 ? per_sm : 1);
        TDT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, segsort_hist_kernel, SS_THREADS, 0));
        hist_cap = sms * (per_sm > 0 ? per_sm : 1);
        TDT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, m3_pass_kernel<false>, SS_THREADS, M3_UPASS_SMEM));
        upass_cap = sms * (per_sm > 0 ? per_sm : 1);
        sm_count = sms;
    }
    const unsigned utiles = (unsigned)(a.L.tiles_max < upass_cap ? a.L.tiles_max : upass_cap);
    const unsigned tiles = (unsigned)(a.L.tiles_max < pass_cap ? a.L.tiles_max : pass_cap);
    const unsigned htiles = (unsigned)(a.L.tiles_max < hist_cap ? a.L.tiles_max : hist_cap);
    cudaStream_t side2 = st;   // generation 3: the later partition rounds and their finish
    if (a.msd) {
        // Generation 3 (tdt_segsort3.cuh).  Nothing is large when the whole input fits the small-segment kernel; no
        // range can exist when it fits one finish batch.
        const int n_rounds = n_max > M3_CAP ? (key_bits + 7) / 8 : 0;
        const bool byval = vals_in == nullptr;   // value = element index: grows with the position
        // The finish kernel's CTAs are short-lived (a few batches each) instead of one persistent wave: the later
        // rounds run next to it on a HIGH-PRIORITY side stream and take the SM resources a retiring CTA frees.  The
        // grids follow the input size -- a CTA that finds no batch still costs its 96 KB shared-memory carve-out, and
        // 2000 of them were 12 us of a 2.5 M-signal shard's sort.
        const int64_t wave = 2 * (int64_t)sm_count;
        int64_t fin_want = n_max / (M3_CAP / 2) + 64;   // about one CTA per expected batch
        if (fin_want > (int64_t)TDT_M3_FIN_WAVES * wave) fin_want = (int64_t)TDT_M3_FIN_WAVES * wave;
        if (fin_want > a.m3.batch_max) fin_want = a.m3.batch_max;
        const unsigned fin_grid = (unsigned)(fin_want > 0 ? fin_want : 1);
        int64_t fin1_want = n_max / (4 * M3_CAP) + 64;   // the later rounds' list: usually a few pile-ups
        if (fin1_want > 4 * wave) fin1_want = 4 * wave;
        const unsigned fin1_grid = (unsigned)fin1_want;
        const char *ser = getenv("TDT_M3_SERIAL");   // measurement aid: everything on the caller's stream
        const bool serial = ser && ser[0] == '1';
        if (n_rounds > 0) {
            // a single STABLE round leaves the data sorted: straight into out (the unordered passes of the by-value
            // mode are always followed by a finish batch)
            const int dst0 = (n_rounds == 1 && !byval) ? 2 : 1;
            TDT_LAUNCH(m3_hist_kernel, htiles, SS_THREADS, 0, st, a, 0);
            TDT_LAUNCH(m3_plan_kernel, (unsigned)((a.m3.rng_max * 32 + 255) / 256), 256, 0, st, a, 0, dst0);
            if (byval) TDT_LAUNCH(m3_pass_kernel<false>, utiles, SS_THREADS, M3_UPASS_SMEM, st, a, 0, dst0);
            else TDT_LAUNCH(m3_pass_kernel<true>, tiles, SS_THREADS, SS_PASS_SMEM, st, a, 0, dst0);
        }
        if (n_rounds > 1) {
#if TDT_SS_FORK
            static thread_local cudaStream_t side2_streams[16 * SS_BRANCHES] = {};
            static thread_local cudaEvent_t fork2_ev[16 * SS_BRANCHES] = {};
            int d2 = 0;
            TDT_CUDA(cudaGetDevice(&d2));
            if (!serial && d2 >= 0 && d2 < 16) {
                d2 = d2 * SS_BRANCHES + g_ss_branch;
                if (!side2_streams[d2]) {
                    int pr_least = 0, pr_greatest = 0;
                    TDT_CUDA(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
                    TDT_CUDA(cudaStreamCreateWithPriority(&side2_streams[d2], cudaStreamNonBlocking, pr_greatest));
                    TDT_CUDA(cudaEventCreateWithFlags(&fork2_ev[d2], cudaEventDisableTiming));
                    TDT_CUDA(cudaEventCreateWithFlags(&g_ss_join2_ev[d2], cudaEventDisableTiming));
                }
                side2 = side2_streams[d2];
                TDT_CUDA(cudaEventRecord(fork2_ev[d2], st));
                TDT_CUDA(cudaStreamWaitEvent(side2, fork2_ev[d2], 0));
                g_ss_join2_idx = d2;
            }
#endif
            for (int round = 1; round < n_rounds; round++) {
                TDT_LAUNCH(m3_hist_kernel, htiles, SS_THREADS, 0, side2, a, round);
                TDT_LAUNCH(m3_plan_kernel, (unsigned)((a.m3.rng_max * 32 + 255) / 256), 256, 0, side2, a, round, round + 1);
                if (byval) TDT_LAUNCH(m3_pass_kernel<false>, utiles, SS_THREADS, M3_UPASS_SMEM, side2, a, round, round + 1);
                else TDT_LAUNCH(m3_pass_kernel<true>, tiles, SS_THREADS, SS_PASS_SMEM, side2, a, round, round + 1);
            }
            if (byval) TDT_LAUNCH(m3_finish_kernel<true>, fin1_grid, M3_THREADS, M3_FIN_SMEM, side2, a, 1);
            else TDT_LAUNCH(m3_finish_kernel<false>, fin1_grid, M3_THREADS, M3_FIN_SMEM, side2, a, 1);
            TDT_M3_DEBUG_DUMP(side2, "list 1");
        }
        if (n_max > SS_LOCAL_MAX) {
            if (byval) TDT_LAUNCH(m3_finish_kernel<true>, fin_grid, M3_THREADS, M3_FIN_SMEM, st, a, 0);
            else TDT_LAUNCH(m3_finish_kernel<false>, fin_grid, M3_THREADS, M3_FIN_SMEM, st, a, 0);
            TDT_M3_DEBUG_DUMP(st, "list 0");
        }
    } else {
    TDT_LAUNCH(segsort_hist_kernel, htiles, SS_THREADS, 0, st, a);
    const int64_t scan_warps = a.L.nlarge_max * SS_MAX_PASSES;
    TDT_LAUNCH(segsort_scan_hist_kernel, (unsigned)((scan_warps * 32 + 255) / 256), 256, 0, st, a);
    // ping-pong so that the LAST pass lands in keys_out / vals_out
    for (int pass = 0; pass < a.n_passes; pass++) {
        const bool to_out = ((a.n_passes - 1 - pass) % 2) == 0;
        const uint32_t *sk = pass == 0 ? keys_in : (to_out ? keys_tmp : keys_out);
        const int32_t *sv = pass == 0 ? vals_in : (to_out ? vals_tmp : vals_out);
        uint32_t *dk = to_out ? keys_out : keys_tmp;
        int32_t *dv = to_out ? vals_out : vals_tmp;
        TDT_LAUNCH(segsort_pass_kernel, tiles, SS_THREADS, SS_PASS_SMEM, st, a, pass, sk, sv, dk, dv);
    }
    }
    if (side2 != st) {
        TDT_CUDA(cudaEventRecord(g_ss_join2_ev[g_ss_join2_idx], side2));
        TDT_CUDA(cudaStreamWaitEvent(st, g_ss_join2_ev[g_ss_join2_idx], 0));
    }
#if TDT_SS_FORK
    if (side != st) {
        TDT_CUDA(cudaEventRecord(join_ev[dev_id], side));
        TDT_CUDA(cudaStreamWaitEvent(st, join_ev[dev_id], 0));
    }
#endif
    return TDT_OK;
}
bool segsort_want_heads(uint32_t *heads) {
    g_ss_heads_out = heads;
    return true;
}

// The entry every caller uses.  (r02 also carried a sample-sort generation -- splitter plan, bucket histogram, partition
// pass, shared-memory interpolation finish; tdt_segsort2.cuh, commit 8a338c8 -- as an opt-in A/B: parity-green, partition
// round 0.28 ms, finish 1.9 / 4.1 ms for posA / posB.  Generation 3 superseded it and it was removed)
int segsort_pairs(const uint32_t *keys_in, const int32_t *vals_in, uint32_t *keys_out, int32_t *vals_out,
                  uint32_t *keys_tmp, int32_t *vals_tmp, const int64_t *off, const int64_t *dims, const int32_t *segid,
                  int64_t n_max, int64_t nseg_max, int key_bits, void *temp, size_t temp_bytes, int *err,
                  cudaStream_t st) {
    return segsort1_pairs(keys_in, vals_in, keys_out, vals_out, keys_tmp, vals_tmp, off, dims, segid, n_max, nseg_max,
                          key_bits, temp, temp_bytes, err, st);
}

#endif  // TDT_SEGSORT_IMPL

}  // namespace tdt

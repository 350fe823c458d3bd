// tdt_segsort2.cuh -- the segmented stable sort, second generation (sm_100a): at most two global passes per element
// instead of a histogram read + four LSD passes.
//
//   1. classify      segment starts become a bitmask over the elements (+ a sentinel at n); segments larger than
//                    S2_LCAP are queued for a partition round.
//   2. partition     (up to two rounds, sample sort): per queued segment a CTA sorts S2_OS * nb strided sample keys
//                    (bitonic, shuffles + shared memory) and keeps nb - 1 of them as SPLITTERS (nb ~ size / 2048,
//                    <= 256); a histogram pass assigns every element its bucket (8-step branch-free search over the
//                    splitters, one byte per element) and counts; the partition pass is the onesweep pass of
//                    tdt_segsort.cuh with the bucket as its digit (stable in-tile ranking, per-digit chained scan
//                    over the earlier tiles of the segment, digit-ordered coalesced write).  Splitters follow the
//                    DATA (equi-depth), so a 50 k-signal hotspot inside a chromosome-wide pair is cut as finely as
//                    the uniform background -- one round resolves both.  Every bucket start becomes a head bit:
//                    buckets ARE segments from here on.  Buckets still larger than S2_LCAP (sampling noise, pairs
//                    beyond 256 * 2048 signals) go through the second round; what is left after that (only heavy
//                    exact duplicates can be) is handed to the LSD chain of tdt_segsort.cuh, which always terminates.
//   3. finish        ONE kernel sorts every segment of <= S2_LCAP elements in shared memory, batches of whole
//                    consecutive segments (a contiguous element range, found from the head bitmask): segments of
//                    <= 32 elements rank by direct comparison; larger ones are placed by INTERPOLATION -- sub-bucket
//                    = (key - min) * count / (max - min + 1), one shared-memory atomic per element, a block scan, and
//                    a comparison fix-up inside the (mostly single-element) sub-buckets.  One global read and one
//                    global write per element, values gathered by original position.
//
// Stability: the partition is stable, ties inside the finish kernel are broken by the position in the batch.
#pragma once
#include <stdlib.h>

#include "tdt_segsort.cuh"

namespace tdt {

constexpr int S2_LCAP = 6144;              // largest segment finished in shared memory
constexpr int S2_W = 2048;                 // element window: the segments that START inside it form a batch
constexpr int S2_BATCH = S2_W + S2_LCAP;   // elements a batch can hold
constexpr int S2_LTHREADS = 1024;
constexpr int S2_EPT = S2_BATCH / S2_LTHREADS;   // elements per thread
constexpr int S2_TINY = 32;                // segments up to this size rank by comparison
#ifndef TDT_S2_TARGET
#define TDT_S2_TARGET 2048
#endif
constexpr int S2_TARGET = TDT_S2_TARGET;   // expected bucket size of a partition round
constexpr int S2_OS = 8;                   // samples per splitter
constexpr int S2_MAXB = 256;               // buckets per round
constexpr int S2_SAMPLES = S2_MAXB * S2_OS;
constexpr int S2_HW = S2_BATCH / 32 + 2;   // head words a window looks at
constexpr u32 S2_OPEN = 0xffffu;           // "no head within reach": the segment is larger than S2_LCAP

struct S2Rec {
    int64_t start;
    int32_t size, tile_base, nb, pad;
};

struct S2Counters {
    int32_t n_rec[2], n_tiles[2];
    u32 ticket;
    int32_t pad[3];
};

struct S2Layout {
    S2Counters *cnt;
    u32 *heads, *lv1, *lv2;      // one bit per element: segment start / data lives in tmp / data lives in out
    S2Rec *rec[2];
    u32 *spl, *hist;             // [2][rec_max][256]
    int32_t *tile_rec[2];
    uint8_t *bkt;                // bucket of every element of a queued segment
    int64_t rec_max, tiles_max, words;
    size_t zero_bytes;
    SSLayout v1;                 // the LSD chain (status array shared with the partition rounds, distinct epochs)
};

static inline int64_t s2_rec_max(int64_t n) { return n / (S2_LCAP + 1) + 2; }

static inline size_t segsort2_temp_bytes(int64_t n, int64_t nseg_max) {
    (void)nseg_max;
    const int64_t rm = s2_rec_max(n), tiles = n / SS_TILE + rm + 1, words = n / 32 + 4;
    return ss_align(sizeof(S2Counters)) + 3 * ss_align((size_t)words * 4) + 2 * ss_align((size_t)rm * sizeof(S2Rec)) +
           2 * ss_align((size_t)rm * 2 * 256 * 4) + 2 * ss_align((size_t)tiles * 4) + ss_align((size_t)n + 64) +
           segsort1_temp_bytes(n, rm) + 1024;
}

// what callers reserve: enough for either generation (TDT_SEGSORT_V2=1 selects the sample sort at run time)
static inline size_t segsort_temp_bytes(int64_t n, int64_t nseg_max) {
    const size_t a = segsort1_temp_bytes(n, nseg_max), b = segsort2_temp_bytes(n, nseg_max);
    return a > b ? a : b;
}

#ifdef TDT_SEGSORT_IMPL

static inline S2Layout s2_layout(void *temp, int64_t n) {
    S2Layout L;
    L.rec_max = s2_rec_max(n);
    L.tiles_max = n / SS_TILE + L.rec_max + 1;
    L.words = n / 32 + 4;
    char *p = (char *)temp;
    L.cnt = (S2Counters *)p; p += ss_align(sizeof(S2Counters));
    L.heads = (u32 *)p; p += ss_align((size_t)L.words * 4);
    L.lv1 = (u32 *)p; p += ss_align((size_t)L.words * 4);
    L.lv2 = (u32 *)p; p += ss_align((size_t)L.words * 4);
    L.zero_bytes = (size_t)(p - (char *)temp);
    for (int r = 0; r < 2; r++) { L.rec[r] = (S2Rec *)p; p += ss_align((size_t)L.rec_max * sizeof(S2Rec)); }
    L.spl = (u32 *)p; p += ss_align((size_t)L.rec_max * 2 * 256 * 4);
    L.hist = (u32 *)p; p += ss_align((size_t)L.rec_max * 2 * 256 * 4);
    for (int r = 0; r < 2; r++) { L.tile_rec[r] = (int32_t *)p; p += ss_align((size_t)L.tiles_max * 4); }
    L.bkt = (uint8_t *)p; p += ss_align((size_t)n + 64);
    L.v1 = ss_layout(p, n, L.rec_max);
    return L;
}

struct S2Args {
    const uint32_t *keys_in;
    const int32_t *vals_in;  // nullptr: value = element index
    uint32_t *keys_out;
    int32_t *vals_out;
    uint32_t *keys_tmp;
    int32_t *vals_tmp;
    const int64_t *off;
    const int64_t *dims;     // device: {n, nseg}
    int key_bits;
    S2Layout L;
    int *err;
};

// ---- 1. classify ---------------------------------------------------------------------------------------
__global__ void s2_classify_kernel(S2Args a) {
    const int64_t n = a.dims[0], nseg = a.dims[1];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s == 0) atomicOr(a.L.heads + (n >> 5), 1u << (n & 31));   // sentinel: the last segment ends here
    int64_t q = 0, size = 0;
    if (s < nseg) {
        q = a.off[s];
        size = a.off[s + 1] - q;
    }
    // consecutive segments mostly share a head word: one atomic per group of lanes
    const bool has = size > 0;
    const u32 act = __ballot_sync(0xffffffffu, has);
    if (has) {
        const int64_t w = q >> 5;
        const u32 peers = __match_any_sync(act, w);
        const u32 bits = __reduce_or_sync(peers, 1u << (q & 31));
        if ((peers & lanemask_lt()) == 0u) atomicOr(a.L.heads + w, bits);
    }
    if (size > S2_LCAP) {
        const int32_t idx = atomicAdd(&a.L.cnt->n_rec[0], 1);
        const int32_t nt = (int32_t)((size + SS_TILE - 1) / SS_TILE);
        S2Rec rec;
        rec.start = q;
        rec.size = (int32_t)size;
        rec.tile_base = atomicAdd(&a.L.cnt->n_tiles[0], nt);
        rec.nb = 0;
        rec.pad = 0;
        a.L.rec[0][idx] = rec;
    }
}

// ---- 2a. plan: splitters of a queued segment ---------------------------------------------------------------
// 2048 strided samples (padded with +inf), bitonic sort: partner distances below 32 by shuffle, 32..512 through
// shared memory, 1024 inside the thread (it holds elements t and t + 1024).
__global__ void __launch_bounds__(1024) s2_plan_kernel(S2Args a, int round) {
    __shared__ u32 xs[S2_SAMPLES];
    const int r = blockIdx.x;
    if (r >= a.L.cnt->n_rec[round]) return;
    const S2Rec rec = a.L.rec[round][r];
    const u32 *src = round == 0 ? a.keys_in : a.keys_tmp;
    int nb = (rec.size + S2_TARGET - 1) / S2_TARGET;
    nb = nb < 2 ? 2 : (nb > S2_MAXB ? S2_MAXB : nb);
    const int S = nb * S2_OS;
    const int t = threadIdx.x;
    u32 v[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
        const int i = t + h * 1024;
        v[h] = i < S ? src[rec.start + (int64_t)i * rec.size / S] : 0xffffffffu;
    }
    for (int k = 2; k <= S2_SAMPLES; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j == 1024) {
                // partner of element t is element t + 1024 (same thread); k == 2048: ascending everywhere
                const u32 lo = min(v[0], v[1]), hi = max(v[0], v[1]);
                v[0] = lo;
                v[1] = hi;
            } else if (j >= 32) {
#pragma unroll
                for (int h = 0; h < 2; h++) xs[t + h * 1024] = v[h];
                __syncthreads();
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = t + h * 1024;
                    const u32 o = xs[i ^ j];
                    const bool asc = (i & k) == 0, lower = (i & j) == 0;
                    v[h] = (lower == asc) ? min(v[h], o) : max(v[h], o);
                }
                __syncthreads();
            } else {
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int i = t + h * 1024;
                    const u32 o = __shfl_xor_sync(0xffffffffu, v[h], j);
                    const bool asc = (i & k) == 0, lower = (i & j) == 0;
                    v[h] = (lower == asc) ? min(v[h], o) : max(v[h], o);
                }
            }
        }
    }
#pragma unroll
    for (int h = 0; h < 2; h++) xs[t + h * 1024] = v[h];
    __syncthreads();
    u32 *spl = a.L.spl + ((size_t)round * a.L.rec_max + r) * 256;
    u32 *hist = a.L.hist + ((size_t)round * a.L.rec_max + r) * 256;
    if (t < 256) {
        spl[t] = t < nb - 1 ? xs[(t + 1) * S2_OS] : 0xffffffffu;   // bucket = number of splitters <= key
        hist[t] = 0u;
    }
    const int32_t nt = (rec.size + SS_TILE - 1) / SS_TILE;
    for (int32_t i = t; i < nt; i += 1024) a.L.tile_rec[round][rec.tile_base + i] = r;
    if (t == 0) a.L.rec[round][r].nb = nb;
}

// ---- 2b. histogram: the bucket of every element, counts per (segment, bucket) ---------------------------------
__global__ void __launch_bounds__(SS_THREADS) s2_hist_kernel(S2Args a, int round) {
    __shared__ u32 spl_s[256], h[256];
    const int n_tiles = a.L.cnt->n_tiles[round];
    const u32 *src = round == 0 ? a.keys_in : a.keys_tmp;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int r = a.L.tile_rec[round][tile];
        const S2Rec rec = a.L.rec[round][r];
        const int64_t t0 = rec.start + (int64_t)(tile - rec.tile_base) * SS_TILE;
        const int64_t rem = rec.start + rec.size - t0;
        const int cnt = rem < SS_TILE ? (int)rem : SS_TILE;
        if (threadIdx.x < 256) {
            spl_s[threadIdx.x] = a.L.spl[((size_t)round * a.L.rec_max + r) * 256 + threadIdx.x];
            h[threadIdx.x] = 0u;
        }
        __syncthreads();
        for (int e = threadIdx.x; e < cnt; e += SS_THREADS) {
            const u32 key = src[t0 + e];
            if (round == 0 && a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
            u32 b = 0;
#pragma unroll
            for (int step = 128; step > 0; step >>= 1)
                if (spl_s[b + step - 1] <= key) b += step;   // spl_s[255] = +inf: b stays <= 255
            a.L.bkt[t0 + e] = (uint8_t)b;
            atomicAdd(&h[b], 1u);
        }
        __syncthreads();
        if (threadIdx.x < 256 && h[threadIdx.x])
            atomicAdd(a.L.hist + ((size_t)round * a.L.rec_max + r) * 256 + threadIdx.x, h[threadIdx.x]);
        __syncthreads();
    }
}

// ---- 2c. partition pass: segsort_pass_kernel with the bucket as digit ------------------------------------------
// The first tile of a segment also does the segment's bookkeeping: head bits at the bucket starts, the level mask
// of the segment's element range, and the queue of the buckets that are still too large.
__device__ __forceinline__ void s2_fill_mask(u32 *mask, int64_t lo, int64_t hi) {   // bits [lo, hi), by one CTA
    const int64_t w0 = lo >> 5, w1 = (hi - 1) >> 5;
    for (int64_t w = w0 + threadIdx.x; w <= w1; w += blockDim.x) {
        u32 m = 0xffffffffu;
        if (w == w0) m &= 0xffffffffu << (lo & 31);
        if (w == w1) m &= 0xffffffffu >> (31 - ((hi - 1) & 31));
        if (m == 0xffffffffu) mask[w] = m; else atomicOr(mask + w, m);
    }
}

__global__ void __launch_bounds__(SS_THREADS, TDT_SS_PASS_MINBLOCKS) s2_pass_kernel(S2Args a, int round) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned char *p = ss_smem;
    uint2 *KV = (uint2 *)p; p += SS_TILE * 8;
    uint32_t(*wh)[256] = (uint32_t(*)[256])p; p += SS_WARPS * 256 * 4;
    uint32_t(*mm)[256] = (uint32_t(*)[256])p; p += SS_MM * SS_WARPS * 256 * 4;
    uint32_t *bin = (uint32_t *)p; p += 256 * 4;
    int64_t *gbase = (int64_t *)p; p += 256 * 8;
    uint8_t *DG = (uint8_t *)p;   // the bucket of every staged element (it cannot be recomputed from the key)

    const int n_tiles = a.L.cnt->n_tiles[round];
    const u32 *src_k = round == 0 ? a.keys_in : a.keys_tmp;
    const int32_t *src_v = round == 0 ? a.vals_in : a.vals_tmp;
    u32 *dst_k = round == 0 ? a.keys_tmp : a.keys_out;
    int32_t *dst_v = round == 0 ? a.vals_tmp : a.vals_out;
    const uint32_t epoch = 5u + (uint32_t)round;   // the LSD chain uses epochs 1..4 on the same status words
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        for (int i = threadIdx.x; i < SS_WARPS * 256; i += SS_THREADS) {
            (&wh[0][0])[i] = 0u;
            if (SS_MM) (&mm[0][0])[i] = 0u;
        }
        const int r = a.L.tile_rec[round][tile];
        const S2Rec L = a.L.rec[round][r];
        const int lt = tile - L.tile_base;
        const int64_t t0 = L.start + (int64_t)lt * SS_TILE;
        const int64_t rem = L.start + L.size - t0;
        const int cnt = rem < SS_TILE ? (int)rem : SS_TILE;
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        constexpr int epw = SS_TILE / SS_WARPS;

        uint32_t key[SS_CHUNKS], dig[SS_CHUNKS];
        int32_t val[SS_CHUNKS];
#pragma unroll
        for (int c = 0; c < SS_CHUNKS; c++) {
            const int e = warp * epw + c * 32 + lane;
            key[c] = 0u;
            val[c] = 0;
            dig[c] = 0u;
            if (e < cnt) {
                key[c] = src_k[t0 + e];
                val[c] = src_v ? src_v[t0 + e] : (int32_t)(t0 + e);
                dig[c] = a.L.bkt[t0 + e];
            }
        }
        // exclusive bucket offsets of the segment (thread = bucket)
        const u32 hcount = threadIdx.x < 256 ? a.L.hist[((size_t)round * a.L.rec_max + r) * 256 + threadIdx.x] : 0u;
        __syncthreads();
        const u32 gh = ss_block_excl_scan_256<SS_WARPS>(hcount, bin);
        if (lt == 0) {   // the segment's bookkeeping
            if (threadIdx.x < 256 && hcount) {
                const int64_t q = L.start + gh;
                atomicOr(a.L.heads + (q >> 5), 1u << (q & 31));
                if (hcount > (u32)S2_LCAP) {
                    const int32_t nt = (int32_t)((hcount + SS_TILE - 1) / SS_TILE);
                    if (round == 0) {
                        const int32_t idx = atomicAdd(&a.L.cnt->n_rec[1], 1);
                        S2Rec rec;
                        rec.start = q;
                        rec.size = (int32_t)hcount;
                        rec.tile_base = atomicAdd(&a.L.cnt->n_tiles[1], nt);
                        rec.nb = 0;
                        rec.pad = 0;
                        a.L.rec[1][idx] = rec;
                    } else {   // still too large after two rounds: the LSD chain sorts it (in keys_out / vals_out)
                        const int32_t idx = atomicAdd(&a.L.v1.cnt->n_large, 1);
                        SSLarge rec;
                        rec.start = q;
                        rec.size = (int64_t)hcount;
                        rec.tile_base = atomicAdd(&a.L.v1.cnt->n_tiles, nt);
                        rec.pad = 0;
                        a.L.v1.large[idx] = rec;
                    }
                }
            }
            s2_fill_mask(round == 0 ? a.L.lv1 : a.L.lv2, L.start, L.start + L.size);
        }
        uint32_t info[SS_CHUNKS];
        ss_count<SS_CHUNKS>(cnt, epw, wh[warp], mm[warp], info, 8, [&](int, int c) -> uint32_t { return dig[c]; });
        __syncthreads();
        uint32_t total, excl;
        ss_digit_bases<SS_WARPS>(wh, bin, total, excl);
        uint32_t *row = a.L.v1.status + (size_t)tile * 256 + threadIdx.x;
        const bool live = threadIdx.x < 256;
        if (live) st_volatile_u32(row, ss_pack(lt == 0 ? 2u : 1u, epoch, total));
        __syncthreads();
        ss_scatter<SS_CHUNKS>(epw, wh[warp], info, [&](int, int c) -> uint32_t { return dig[c]; },
                              [&](int, int c, uint32_t pos) {
                                  KV[pos] = make_uint2(key[c], (uint32_t)val[c]);
                                  DG[pos] = (uint8_t)dig[c];
                              });
        {
            uint32_t before = 0;
            if (lt != 0 && live) {
                constexpr int LB = TDT_SS_LB;
                const uint32_t *prow = row - 256;
                int left = lt;
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
            const uint2 kv = KV[i];
            const int64_t g = gbase[DG[i]] + i;
            dst_k[g] = kv.x;
            dst_v[g] = (int32_t)kv.y;
        }
        __syncthreads();
    }
}

// ---- 3. finish: every segment of <= S2_LCAP elements, in shared memory ------------------------------------------
constexpr size_t S2_PASS_SMEM = SS_PASS_SMEM + SS_TILE;
constexpr size_t S2_LOCAL_SMEM = (size_t)S2_BATCH * (4 + 4 + 4 + 2) + (size_t)(S2_BATCH + 16) * 4 + (size_t)(S2_W + 8) * (2 + 4 + 4) +
                                 (size_t)(S2_HW + 2) * 4 * 4 + 256;

#ifdef TDT_S2_PROFILE
__device__ unsigned long long g_s2_prof[16];
#define S2_T(k)                                                     \
    do {                                                            \
        if (threadIdx.x == 0) {                                     \
            const long long now__ = clock64();                      \
            atomicAdd(&g_s2_prof[k], (unsigned long long)(now__ - t_prev)); \
            t_prev = now__;                                         \
        }                                                           \
    } while (0)
#else
#define S2_T(k)
#endif

__global__ void __launch_bounds__(S2_LTHREADS, 1) s2_local_kernel(S2Args a) {
#ifdef TDT_S2_PROFILE
    long long t_prev = clock64();
#endif
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned char *p = ss_smem;
    u32 *K = (u32 *)p; p += S2_BATCH * 4;
    u32 *K2 = (u32 *)p; p += S2_BATCH * 4;
    int32_t *V = (int32_t *)p; p += S2_BATCH * 4;
    u32 *cn = (u32 *)p; p += (S2_BATCH + 16) * 4;
    u32 *slo = (u32 *)p; p += (S2_W + 8) * 4;
    float *ssc = (float *)p; p += (S2_W + 8) * 4;
    u32 *hw = (u32 *)p; p += (S2_HW + 2) * 4;
    u32 *l1w = (u32 *)p; p += (S2_HW + 2) * 4;
    u32 *l2w = (u32 *)p; p += (S2_HW + 2) * 4;
    u32 *wpre = (u32 *)p; p += (S2_HW + 2) * 4;
    uint16_t *P2 = (uint16_t *)p; p += S2_BATCH * 2;
    uint16_t *hpos = (uint16_t *)p;
    __shared__ u32 s_warp[S2_LTHREADS / 32];
    __shared__ u32 s_chunk, s_first, s_total;

    const int64_t n = a.dims[0];
    const int64_t nchunks = (n + S2_W - 1) / S2_W;
    const int64_t last_word = n >> 5;   // the word that holds the sentinel
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    while (true) {
        __syncthreads();
        if (tid == 0) s_chunk = atomicAdd(&a.L.cnt->ticket, 1u);
        __syncthreads();
        const int64_t chunk = s_chunk;
        if (chunk >= nchunks) break;
        const int64_t base = chunk * S2_W;
        const int64_t w0 = base >> 5;
        // ---- head / level words of the window, prefix popcounts ----------------------------------------
        u32 myw = 0;
        if (tid < S2_HW) {
            const int64_t w = w0 + tid;
            const bool in = w <= last_word;
            myw = in ? a.L.heads[w] : 0u;
            hw[tid] = myw;
            l1w[tid] = in ? a.L.lv1[w] : 0u;
            l2w[tid] = in ? a.L.lv2[w] : 0u;
        }
        {
            const u32 c = __popc(myw);
            u32 inc = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += t;
            }
            if (lane == 31) s_warp[warp] = inc;
            __syncthreads();
            u32 wb = 0;
            for (int w = 0; w < warp; w++) wb += s_warp[w];
            if (tid < S2_HW) wpre[tid] = wb + inc - c;
            if (tid == S2_HW - 1) s_total = wb + inc;
        }
        __syncthreads();
        const int nh = (int)wpre[S2_W / 32];   // segments that start inside this chunk
        S2_T(0);
        if (nh == 0) continue;
        if (tid < S2_HW) {   // positions of the heads (rank <= nh: the chunk's heads and the first one after it)
            u32 bits = myw;
            u32 rank = wpre[tid];
            while (bits && rank <= (u32)nh) {
                const int b = __ffs(bits) - 1;
                hpos[rank] = (uint16_t)(tid * 32 + b);
                bits &= bits - 1;
                rank++;
            }
        }
        if (tid == 0 && s_total == (u32)nh) hpos[nh] = (uint16_t)S2_OPEN;   // no head within reach
        // ---- batches: maximal runs of consecutive segments that fit --------------------------------------
        int r0 = 0;
        while (r0 < nh) {
            __syncthreads();
            if (tid == 0) s_first = (u32)nh;
            __syncthreads();
            for (int r = r0 + tid; r < nh; r += S2_LTHREADS) {
                const u32 e1 = hpos[r + 1];
                if (e1 == S2_OPEN || e1 - hpos[r] > (u32)S2_LCAP) atomicMin(&s_first, (u32)r);
            }
            __syncthreads();
            const int r1 = (int)s_first;   // first segment at or after r0 that does not fit
            if (r1 == r0) {
                r0++;
                continue;
            }
            S2_T(1);
            const int b_rel = hpos[r0];
            const int count = hpos[r1] - b_rel;
            const int64_t g0 = base + b_rel;
            // ---- load -------------------------------------------------------------------------------------
            for (int e = tid; e < S2_BATCH + 16; e += S2_LTHREADS) cn[e] = 0u;
            for (int e = tid; e < count; e += S2_LTHREADS) {
                const int pos = b_rel + e;
                const bool t2 = (l2w[pos >> 5] >> (pos & 31)) & 1u, t1 = (l1w[pos >> 5] >> (pos & 31)) & 1u;
                const int64_t g = g0 + e;
                const u32 key = t2 ? a.keys_out[g] : (t1 ? a.keys_tmp[g] : a.keys_in[g]);
                if (a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
                K[e] = key;
                if (t2) V[e] = a.vals_out[g];
            }
            __syncthreads();
            S2_T(2);
            // ---- key range of the larger segments (one warp each) ---------------------------------------------
            for (int r = r0 + warp; r < r1; r += S2_LTHREADS / 32) {
                const int st = hpos[r] - b_rel, c = hpos[r + 1] - hpos[r];
                if (c <= S2_TINY) continue;
                u32 mn = 0xffffffffu, mx = 0u;
                for (int j = lane; j < c; j += 32) {
                    const u32 k = K[st + j];
                    mn = min(mn, k);
                    mx = max(mx, k);
                }
                mn = __reduce_min_sync(0xffffffffu, mn);
                mx = __reduce_max_sync(0xffffffffu, mx);
                if (lane == 0) {
                    slo[r] = mn;
                    ssc[r] = mx == mn ? -1.0f : (float)c / ((float)(mx - mn) + 1.0f);   // -1: all keys equal, keep the order
                }
            }
            __syncthreads();
            S2_T(3);
            // ---- rank: small segments by comparison, larger ones counted into interpolated sub-buckets ---------
            u32 pk[S2_EPT];
#pragma unroll
            for (int k = 0; k < S2_EPT; k++) {
                const int e = tid + k * S2_LTHREADS;
                pk[k] = 0xffffffffu;
                if (e < count) {
                    const int pos = b_rel + e;
                    const int r = (int)(wpre[pos >> 5] + __popc(hw[pos >> 5] & (0xffffffffu >> (31 - (pos & 31))))) - 1;
                    const int st = hpos[r] - b_rel, c = hpos[r + 1] - hpos[r];
                    const u32 key = K[e];
                    if (c <= S2_TINY) {
                        int rank = 0;
                        for (int j = 0; j < c; j++) {
                            const u32 o = K[st + j];
                            rank += (o < key) || (o == key && st + j < e);
                        }
                        K2[st + rank] = key;
                        P2[st + rank] = (uint16_t)e;
                    } else if (ssc[r] < 0.0f) {
                        K2[e] = key;
                        P2[e] = (uint16_t)e;
                    } else {
                        int sb = (int)((float)(key - slo[r]) * ssc[r]);
                        sb = sb > c - 1 ? c - 1 : sb;
                        const u32 slot = atomicAdd(&cn[st + sb], 1u);
                        pk[k] = ((u32)sb << 16) | slot;
                    }
                }
            }
            __syncthreads();
            S2_T(4);
            // ---- exclusive scan of the sub-bucket counts (8 consecutive entries per thread) ----------------------
            {
                uint4 x0 = reinterpret_cast<uint4 *>(cn)[2 * tid], x1 = reinterpret_cast<uint4 *>(cn)[2 * tid + 1];
                const u32 sum = x0.x + x0.y + x0.z + x0.w + x1.x + x1.y + x1.z + x1.w;
                u32 inc = sum;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
                    if (lane >= o) inc += t;
                }
                if (lane == 31) s_warp[warp] = inc;
                __syncthreads();
                u32 run = inc - sum;
                for (int w = 0; w < warp; w++) run += s_warp[w];
                uint4 y0, y1;
                y0.x = run; run += x0.x; y0.y = run; run += x0.y; y0.z = run; run += x0.z; y0.w = run; run += x0.w;
                y1.x = run; run += x1.x; y1.y = run; run += x1.y; y1.z = run; run += x1.z; y1.w = run; run += x1.w;
                reinterpret_cast<uint4 *>(cn)[2 * tid] = y0;
                reinterpret_cast<uint4 *>(cn)[2 * tid + 1] = y1;
                if (tid == S2_LTHREADS - 1) cn[S2_BATCH] = run;
            }
            __syncthreads();
            S2_T(5);
            // ---- place the counted elements (unordered inside a sub-bucket) -------------------------------------
#pragma unroll
            for (int k = 0; k < S2_EPT; k++) {
                if (pk[k] != 0xffffffffu) {
                    const int e = tid + k * S2_LTHREADS;
                    const int pos = b_rel + e;
                    const int r = (int)(wpre[pos >> 5] + __popc(hw[pos >> 5] & (0xffffffffu >> (31 - (pos & 31))))) - 1;
                    const int st = hpos[r] - b_rel;
                    const int sb = (int)(pk[k] >> 16), slot = (int)(pk[k] & 0xffffu);
                    const int d0 = st + (int)(cn[st + sb] - cn[st]) + slot;
                    K2[d0] = K[e];
                    P2[d0] = (uint16_t)e;
                }
            }
            __syncthreads();
            S2_T(6);
            // ---- write: order inside the sub-buckets by (key, original position), values by original position --
            for (int i = tid; i < count; i += S2_LTHREADS) {
                const int pos = b_rel + i;
                const int r = (int)(wpre[pos >> 5] + __popc(hw[pos >> 5] & (0xffffffffu >> (31 - (pos & 31))))) - 1;
                const int st = hpos[r] - b_rel, c = hpos[r + 1] - hpos[r];
                const u32 key = K2[i];
                const int pp = P2[i];
                int f = i;
                if (c > S2_TINY && ssc[r] >= 0.0f) {
                    int sb = (int)((float)(key - slo[r]) * ssc[r]);
                    sb = sb > c - 1 ? c - 1 : sb;
                    const int a0 = st + (int)(cn[st + sb] - cn[st]), a1 = st + (int)(cn[st + sb + 1] - cn[st]);
                    if (a1 - a0 > 1) {
                        int rank = 0;
                        for (int j = a0; j < a1; j++) {
                            const u32 o = K2[j];
                            rank += (o < key) || (o == key && (int)P2[j] < pp);
                        }
                        f = a0 + rank;
                    }
                }
                const bool t2 = (l2w[pos >> 5] >> (pos & 31)) & 1u, t1 = (l1w[pos >> 5] >> (pos & 31)) & 1u;
                int32_t val;
                if (t2) val = V[pp];
                else if (t1) val = a.vals_tmp[g0 + pp];
                else val = a.vals_in ? a.vals_in[g0 + pp] : (int32_t)(g0 + pp);
                a.keys_out[g0 + f] = key;
                a.vals_out[g0 + f] = val;
            }
            S2_T(7);
            r0 = r1;
        }
    }
}

// ---- host launcher -------------------------------------------------------------------------------------------
static int segsort2_pairs(const uint32_t *keys_in, const int32_t *vals_in, uint32_t *keys_out, int32_t *vals_out,
                          uint32_t *keys_tmp, int32_t *vals_tmp, const int64_t *off, const int64_t *dims, int64_t n_max,
                          int64_t nseg_max, int key_bits, void *temp, size_t temp_bytes, int *err, cudaStream_t st) {
    if (n_max <= 0 || nseg_max <= 0) return TDT_OK;
    if (temp_bytes < segsort2_temp_bytes(n_max, nseg_max))
        return fail(TDT_E_WORKSPACE, "segmented sort needs %zu bytes of temporary storage, %zu reserved",
                    segsort2_temp_bytes(n_max, nseg_max), temp_bytes);
    if (key_bits < 1) key_bits = 1;
    if (key_bits > 32) key_bits = 32;
    S2Args a;
    a.keys_in = keys_in;
    a.vals_in = vals_in;
    a.keys_out = keys_out;
    a.vals_out = vals_out;
    a.keys_tmp = keys_tmp;
    a.vals_tmp = vals_tmp;
    a.off = off;
    a.dims = dims;
    a.key_bits = key_bits;
    a.L = s2_layout(temp, n_max);
    a.err = err;
    TDT_CUDA(cudaMemsetAsync(temp, 0, a.L.zero_bytes, st));
    TDT_CUDA(cudaMemsetAsync(a.L.v1.cnt, 0, a.L.v1.zero_bytes, st));

    static thread_local bool configured = false;
    static thread_local int sms = 0, pass_cap = 0, hist_cap = 0;
    if (!configured) {
        int dev = 0, per_sm = 0;
        TDT_CUDA(cudaGetDevice(&dev));
        TDT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        TDT_CUDA(cudaFuncSetAttribute(s2_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_LOCAL_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(s2_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S2_PASS_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(segsort_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SS_PASS_SMEM));
        TDT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s2_pass_kernel, SS_THREADS, S2_PASS_SMEM));
        pass_cap = sms * (per_sm > 0 ? per_sm : 1);
        TDT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, s2_hist_kernel, SS_THREADS, 0));
        hist_cap = sms * (per_sm > 0 ? per_sm : 1);
        configured = true;
    }
    TDT_LAUNCH(s2_classify_kernel, (unsigned)((nseg_max + 255) / 256), 256, 0, st, a);
    const unsigned ptiles = (unsigned)(a.L.tiles_max < pass_cap ? a.L.tiles_max : pass_cap);
    const unsigned htiles = (unsigned)(a.L.tiles_max < hist_cap ? a.L.tiles_max : hist_cap);
    const int rounds = n_max > S2_LCAP ? 2 : 0;   // nothing can be queued when the whole input fits a batch
    for (int round = 0; round < rounds; round++) {
        TDT_LAUNCH(s2_plan_kernel, (unsigned)a.L.rec_max, 1024, 0, st, a, round);
        TDT_LAUNCH(s2_hist_kernel, htiles, SS_THREADS, 0, st, a, round);
        TDT_LAUNCH(s2_pass_kernel, ptiles, SS_THREADS, S2_PASS_SMEM, st, a, round);
    }
    // the LSD chain for whatever two rounds could not cut (a side branch: it touches other segments than the finish kernel)
    cudaStream_t side = st;
    static thread_local cudaStream_t side_streams[16 * SS_BRANCHES] = {};
    static thread_local cudaEvent_t fork_ev[16 * SS_BRANCHES] = {}, join_ev[16 * SS_BRANCHES] = {};
    int dev_id = 0;
    TDT_CUDA(cudaGetDevice(&dev_id));
    if (rounds && dev_id >= 0 && dev_id < 16) {
        dev_id = dev_id * SS_BRANCHES + g_ss_branch;
        if (!side_streams[dev_id]) {
            TDT_CUDA(cudaStreamCreateWithFlags(&side_streams[dev_id], cudaStreamNonBlocking));
            TDT_CUDA(cudaEventCreateWithFlags(&fork_ev[dev_id], cudaEventDisableTiming));
            TDT_CUDA(cudaEventCreateWithFlags(&join_ev[dev_id], cudaEventDisableTiming));
        }
        side = side_streams[dev_id];
        TDT_CUDA(cudaEventRecord(fork_ev[dev_id], st));
        TDT_CUDA(cudaStreamWaitEvent(side, fork_ev[dev_id], 0));
    }
    if (rounds) {
        SSArgs f;
        f.keys_in = keys_out;   // the leftovers sit in the output buffers after round two
        f.vals_in = vals_out;
        f.keys_out = keys_out;
        f.vals_out = vals_out;
        f.keys_tmp = keys_tmp;
        f.vals_tmp = vals_tmp;
        f.off = off;
        f.dims = dims;
        f.segid = nullptr;
        f.tiny_max = 0;
        f.key_bits = key_bits;
        f.n_passes = (key_bits + 7) / 8;
        f.bits_per_pass = (key_bits + f.n_passes - 1) / f.n_passes;
        f.L = a.L.v1;
        f.err = err;
        const unsigned ftiles = (unsigned)(f.L.tiles_max < pass_cap ? f.L.tiles_max : pass_cap);
        TDT_LAUNCH(segsort_fill_kernel, (unsigned)f.L.nlarge_max, 256, 0, side, f);
        TDT_LAUNCH(segsort_hist_kernel, htiles, SS_THREADS, 0, side, f);
        const int64_t scan_warps = f.L.nlarge_max * SS_MAX_PASSES;
        TDT_LAUNCH(segsort_scan_hist_kernel, (unsigned)((scan_warps * 32 + 255) / 256), 256, 0, side, f);
        for (int pass = 0; pass < f.n_passes; pass++) {
            const bool to_out = ((f.n_passes - 1 - pass) % 2) == 0;
            const uint32_t *sk = pass == 0 ? f.keys_in : (to_out ? keys_tmp : keys_out);
            const int32_t *sv = pass == 0 ? f.vals_in : (to_out ? vals_tmp : vals_out);
            uint32_t *dk = to_out ? keys_out : keys_tmp;
            int32_t *dv = to_out ? vals_out : vals_tmp;
            TDT_LAUNCH(segsort_pass_kernel, ftiles, SS_THREADS, SS_PASS_SMEM, side, f, pass, sk, sv, dk, dv);
        }
    }
    TDT_LAUNCH(s2_local_kernel, (unsigned)sms, S2_LTHREADS, S2_LOCAL_SMEM, st, a);
#ifdef TDT_S2_PROFILE
    {
        unsigned long long h[16];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_s2_prof, sizeof(h));
        fprintf(stderr, "s2_local phases (Mcycles of thread 0, summed over CTAs):");
        for (int i = 0; i < 8; i++) fprintf(stderr, " P%d=%.2f", i, h[i] / 1e6);
        fprintf(stderr, "\n");
        memset(h, 0, sizeof(h));
        cudaMemcpyToSymbol(g_s2_prof, h, sizeof(h));
    }
#endif
    if (side != st) {
        TDT_CUDA(cudaEventRecord(join_ev[dev_id], side));
        TDT_CUDA(cudaStreamWaitEvent(st, join_ev[dev_id], 0));
    }
    return TDT_OK;
}

bool segsort_want_heads(uint32_t *heads) {
    const char *e = getenv("TDT_SEGSORT_V2");
    if (e && e[0] == '1') return false;
    g_ss_heads_out = heads;
    return true;
}

// The entry every caller uses.  The LSD sort of tdt_segsort.cuh is the production path; TDT_SEGSORT_V2=1 in the
// environment (read at every call) selects the sample-sort generation above for A/B runs and its parity tests.
// Measured on B200 (r02, 30X set): partition round 0.28 ms (plan 0.03 + histogram 0.10 + pass 0.16) against 0.60 ms for
// the whole LSD sort of posA -- but the shared-memory finish kernel takes 1.9 ms (posA) / 4.1 ms (posB): with one
// 1024-thread CTA per SM (170 KB of shared memory) and ~2048 elements per batch every one of its ten barrier-separated
// phases is pure latency (in-kernel phase timers: 41 % in the output phase, 25 % in finding the batch, 12 % loading).
int segsort_pairs(const uint32_t *keys_in, const int32_t *vals_in, uint32_t *keys_out, int32_t *vals_out,
                  uint32_t *keys_tmp, int32_t *vals_tmp, const int64_t *off, const int64_t *dims, const int32_t *segid,
                  int64_t n_max, int64_t nseg_max, int key_bits, void *temp, size_t temp_bytes, int *err,
                  cudaStream_t st) {
    const char *e = getenv("TDT_SEGSORT_V2");
    const bool use_v2 = e && e[0] == '1';
    if (!use_v2)
        return segsort1_pairs(keys_in, vals_in, keys_out, vals_out, keys_tmp, vals_tmp, off, dims, segid, n_max, nseg_max,
                              key_bits, temp, temp_bytes, err, st);
    return segsort2_pairs(keys_in, vals_in, keys_out, vals_out, keys_tmp, vals_tmp, off, dims, n_max, nseg_max, key_bits,
                          temp, temp_bytes, err, st);
}

#endif  // TDT_SEGSORT_IMPL

}  // namespace tdt

// tdt_cluster.cu -- TIDDIT signal clustering on B200 (sm_100a).
//
// Replaces, for all (chrA,chrB) pairs at once, the reference's per-pair
//     sorted(key=posA) -> DBSCAN.main -> labels back in insertion order
// (tiddit/tiddit_cluster.pyx:140-160, tiddit/DBSCAN.py:33-129).  The reference "DBSCAN" is a
// sliding-window run detector on the posA-sorted signals followed by the same detector on posB inside
// every x-cluster; the kernels below evaluate its closed form:
//
//   ok[i]      = the window of the next `reach` signals of i's segment stays within eps
//   run        = maximal interval of consecutive ok
//   label[j]   = index of the latest run whose start is <= j, if an ok lies in [j-m+1, j]; else noise
//
// Pipeline (no host synchronisation until the final status read; all sizes that depend on the data --
// #x-runs, #labelled signals -- stay on the device, grids are sized by upper bounds):
//   segsort (posA per pair)      hand-written segmented radix sort, 32-bit keys          [tdt_segsort.cuh]
//   heads_from_offsets           segment starts as a bitmask
//   window_runs<X>               THE eps-range-query kernel: 4 B key in, run-start / labelled bitmasks out
//   pair_first_run               #runs before every pair (ids are per pair)
//   pack_y                       labelled signals compacted: posB gathered, x-run id, x-run offsets
//   segsort (posB per x-run)     same sort, segments = x-runs
//   window_runs<Y>               eps-range query on posB inside every x-run
//   group_extra_scan/bases       DBSCAN.py:113-122 id arithmetic per x-run
//   final_labels                 reference ids scattered back to insertion order
#include "tdt_common.cuh"
#define TDT_SEGSORT_IMPL
#include "tdt_segsort.cuh"

namespace tdt {

constexpr int WR_THREADS = 256;
constexpr int WR_WARPS = WR_THREADS / 32;
constexpr int WR_TILE = 4096;            // signals per CTA
constexpr int WR_WORDS = WR_TILE / 32;   // ballot words per tile
constexpr int WR_MAX_M = 8192;           // shared-memory halo limit: (TILE + 2*m) keys per CTA
constexpr uint32_t WR_TMA_CHUNK = 32768; // bytes per bulk copy

enum { MODE_X = 0, MODE_Y = 1 };
enum { ERR_NONE = 0, ERR_RANGE_A = 1, ERR_RANGE_B = 2, ERR_PAIR = 3 };

struct Dims {          // device-resident sizes
    int64_t n, nseg;
};

struct WRParams {
    const u32 *keys;      // coordinates, sorted inside every segment; readable up to the next multiple of 4 past n
    const u32 *heads;     // bit j set <=> j is the first element of a segment; zero beyond n
    const Dims *dims;     // dims->n = number of elements
    int m;                // min_pts
    u32 eps;              // 0: nothing is ever within eps
    u64 *status;          // look-back tile states, zeroed
    u32 *stw;             // out: run-start bits      (one word per 32 elements)
    u32 *cvw;             // out: "labelled" bits
    u32 *tile_pref;       // out: [tile][2] = run starts before the tile, second sum before the tile
    int two_phase;        // 1: tile_pref receives the tile's OWN two sums; wr_tile_scan_kernel scans them in place
    u32 *totals;          // out: [0] = #runs, [1] = second sum (X: #labelled elements, Y: #segments)
    int32_t *gcnt_start;  // Y: [segment rank] = run starts before the segment; [#segments] = #runs
    int32_t *rank_of;     // Y, optional: rank_of[seg_value[j]] = segment rank, written at segment heads
    const int32_t *seg_value;
};

static size_t wr_smem_bytes(int m) {
    const size_t HL = ((size_t)(m - 1) + 31) & ~(size_t)31;
    const size_t HR = ((size_t)m + 3) & ~(size_t)3;
    const size_t head_words = (HL + WR_TILE + HR) / 32 + 2;
    return (HL + WR_TILE + HR) * 4 + ((HL + WR_TILE) / 32 + head_words + 4 * WR_WORDS) * sizeof(u32);
}

// any segment head among shared-memory bit positions [a, b] (a <= b)?
__device__ __forceinline__ bool any_head(const u32 *hw, int a, int b) {
    const int wa = a >> 5, wb = b >> 5;
    const u32 first = 0xffffffffu << (a & 31);
    const u32 last = 0xffffffffu >> (31 - (b & 31));
    if (wa == wb) return (hw[wa] & first & last) != 0u;
    if (hw[wa] & first) return true;
    for (int w = wa + 1; w < wb; w++)
        if (hw[w]) return true;
    return (hw[wb] & last) != 0u;
}

// ----------------------------------------------------------------------------------------------
// The eps-range-query + run-labelling kernel (DBSCAN.py:40-62 for posA, :90-110 for posB).
//
// One CTA = one tile of WR_TILE sorted coordinates.  Thread 0 takes the tile ticket and issues 1-D TMA
// bulk copies (UBLKCP) of the tile plus a left halo of m-1 and a right halo of m keys into shared
// memory; every warp then evaluates 32 windows at a time and packs the outcome with a ballot:
//   X (reach m, DBSCAN.py:44 slices data[i+1:i+m+1], cut at the segment end):
//        ok[i] = i+m in i's segment ? key[i+m] - key[i] < eps
//                                   : i+m-1 in i's segment && key[i+m-1] - key[i] < eps
//   Y (reach m-1, DBSCAN.py:93 slices y[i+1:i+m], never cut inside range(0, k-m+1)):
//        ok[i] = i+m-1 in i's segment && key[i+m-1] - key[i] < eps
// (GENERAL evaluates the reference's max(|x[j]-x[i]|) literally for input that is not sorted.)
// "t in i's segment" <=> no segment-head bit in (i, t].  The last m-1 elements of a segment are never ok, so
// neither run starts nor the m-1 look-back of the labelling rule leak across segments, and a left halo of
// m-1 makes both local to the tile: only the NUMBER of run starts (and a second sum) before the tile is
// carried, by the single-pass look-back scan of tdt_common.cuh.  Output is two bits per element.
// ----------------------------------------------------------------------------------------------
template <int MODE, bool GENERAL>
__global__ void __launch_bounds__(WR_THREADS) window_runs_kernel(const WRParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ u32 s_pref[4];  // exS, exC, aggS, aggC

    const int m = p.m;
    const int HL = ((m - 1) + 31) & ~31;
    const int HR = (m + 3) & ~3;
    const int EXT = HL + WR_TILE;  // positions whose ok bit is evaluated here
    const int HEADW = (EXT + HR) / 32 + 2;
    u32 *keys_s = (u32 *)smem_raw;
    u32 *okw = keys_s + (HL + WR_TILE + HR);
    u32 *hw = okw + EXT / 32;
    u32 *stw = hw + HEADW;
    u32 *cvw = stw + WR_WORDS;
    u32 *pfxS = cvw + WR_WORDS;
    u32 *pfxC = pfxS + WR_WORDS;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n = p.dims->n;

    // Tiles are taken in blockIdx order: the look-back below waits only on tiles with a smaller index, which
    // the hardware has already scheduled (the forward-progress assumption CUB's DeviceScan makes as well).
    const int tile = blockIdx.x;
    const int64_t tile_base = (int64_t)tile * WR_TILE;
    if (tile_base >= n) return;               // grids are sized by an upper bound of n
    const int64_t ext_base = tile_base - HL;  // global index of shared position 0

    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        const int64_t n_pad = (n + 3) & ~(int64_t)3;
        const int64_t jlo = ext_base < 0 ? 0 : ext_base;
        int64_t jhi = tile_base + WR_TILE + HR;
        if (jhi > n_pad) jhi = n_pad;
        const uint32_t bytes = (uint32_t)((jhi - jlo) * 4);
        mbar_expect_tx(&mbar, bytes);
        const unsigned char *src = (const unsigned char *)(p.keys + jlo);
        unsigned char *dst = (unsigned char *)(keys_s + (jlo - ext_base));
        for (uint32_t off = 0; off < bytes; off += WR_TMA_CHUNK) {
            const uint32_t len = bytes - off < WR_TMA_CHUNK ? bytes - off : WR_TMA_CHUNK;
            tma_load_1d(dst + off, src + off, len, &mbar);
        }
    }
    {   // segment-head words of [ext_base, ext_base + EXT + HR + 32): plain loads, they overlap the bulk copy
        const int64_t w0 = ext_base >> 5;  // ext_base is a multiple of 32 (may be negative)
        const int64_t nwords = (n + 31) >> 5;
        for (int i = threadIdx.x; i < HEADW; i += WR_THREADS) {
            const int64_t w = w0 + i;
            hw[i] = (w >= 0 && w < nwords) ? p.heads[w] : 0u;
        }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);

    // ---- phase 1: one ballot word per 32 windows -------------------------------------------
    const u32 mask_m = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);             // heads in (e, e+m]
    const u32 mask_m1 = m - 1 >= 32 ? 0xffffffffu : ((1u << (m - 1)) - 1u);  // heads in (e, e+m-1]
    for (int w = warp; w < EXT / 32; w += WR_WARPS) {
        const int e = w * 32 + lane;
        const int64_t j = ext_base + e;
        bool ok = false;
        if (j >= 0 && j + m - 1 < n) {
            const u32 kj = keys_s[e];
            if (!GENERAL && m <= 32) {
                // head bits of positions e+1 .. e+32 in one funnel shift of two head words
                const u32 hb = lane == 31 ? hw[w + 1] : __funnelshift_r(hw[w], hw[w + 1], lane + 1);
                if (MODE == MODE_Y) {
                    ok = !(hb & mask_m1) && (keys_s[e + m - 1] - kj < p.eps);
                } else if (j + m < n && !(hb & mask_m)) {
                    ok = keys_s[e + m] - kj < p.eps;
                } else {
                    ok = !(hb & mask_m1) && (keys_s[e + m - 1] - kj < p.eps);
                }
            } else if (GENERAL) {
                // DBSCAN.py:44-51 literally: max |x[i+d] - x[i]| over the next m elements (cut at n); one segment
                const int64_t left = n - 1 - j;
                const int cnt = left < (int64_t)m ? (int)left : m;
                u32 dmax = 0;
                for (int d = 1; d <= cnt; d++) {
                    const u32 kt = keys_s[e + d];
                    const u32 dd = kt > kj ? kt - kj : kj - kt;
                    dmax = dd > dmax ? dd : dmax;
                }
                ok = dmax < p.eps;
            } else if (MODE == MODE_Y) {
                ok = !any_head(hw, e + 1, e + m - 1) && (keys_s[e + m - 1] - kj < p.eps);
            } else {
                if (j + m < n && !any_head(hw, e + 1, e + m)) {
                    ok = keys_s[e + m] - kj < p.eps;
                } else {
                    ok = !any_head(hw, e + 1, e + m - 1) && (keys_s[e + m - 1] - kj < p.eps);
                }
            }
        }
        const u32 word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) okw[w] = word;
    }
    __syncthreads();

    // ---- phase 2: run-start bits and "labelled" bits of the tile's own words ----------------
    for (int tw = warp; tw < WR_WORDS; tw += WR_WARPS) {
        const int w = HL / 32 + tw;
        const u32 cur = okw[w], prev = okw[w - 1];
        const u32 st = cur & ~((cur << 1) | (prev >> 31));
        bool c;
        if (m <= 32) {
            // an ok among the m positions ending at this lane: bits [33+lane-m, 32+lane] of prev:cur
            const u64 v = ((u64)cur << 32) | (u64)prev;
            const u64 mask = (m == 32) ? 0xffffffffull : ((1ull << m) - 1ull);
            c = ((v >> (33 + lane - m)) & mask) != 0ull;
        } else {
            c = (cur & lanemask_le()) != 0u;
            if (!c) {
                for (int d = 1; d <= HL / 32; d++) {
                    const u32 ww = okw[w - d];
                    if (ww) {
                        const int last = 32 * (w - d) + 31 - __clz((int)ww);
                        c = (32 * w + lane - last) <= m - 1;
                        break;
                    }
                }
            }
        }
        const u32 cv = __ballot_sync(0xffffffffu, c);
        if (lane == 0) {
            stw[tw] = st;
            cvw[tw] = cv;
        }
    }
    __syncthreads();

    // ---- phase 3: word prefixes inside the tile, then the carry from earlier tiles ----------
    if (warp == 0) {
        constexpr int PER = WR_WORDS / 32;
        u32 s[PER], c[PER], ts = 0, tc = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            s[k] = __popc(stw[lane * PER + k]);
            // second sum: labelled elements (X, feeds the compaction) or segment heads (Y, ranks the segments)
            c[k] = MODE == MODE_X ? __popc(cvw[lane * PER + k]) : __popc(hw[HL / 32 + lane * PER + k]);
            ts += s[k];
            tc += c[k];
        }
        u32 is = ts, ic = tc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, is, o);
            const u32 b = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) {
                is += a;
                ic += b;
            }
        }
        u32 es = is - ts, ec = ic - tc;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            pfxS[lane * PER + k] = es;
            pfxC[lane * PER + k] = ec;
            es += s[k];
            ec += c[k];
        }
        const u32 aggS = __shfl_sync(0xffffffffu, is, 31);
        const u32 aggC = __shfl_sync(0xffffffffu, ic, 31);
        u32 exS, exC;
        lookback(p.status, tile, aggS, aggC, exS, exC);
        if (lane == 0) {
            s_pref[0] = exS;
            s_pref[1] = exC;
            s_pref[2] = aggS;
            s_pref[3] = aggC;
            p.tile_pref[2 * (int64_t)tile] = exS;
            p.tile_pref[2 * (int64_t)tile + 1] = exC;
        }
    } else {
        // the two result bits per element (words beyond n are zero: their ok bits are zero)
        for (int tw = threadIdx.x - 32; tw < WR_WORDS; tw += WR_THREADS - 32) {
            const int64_t gw = (tile_base >> 5) + tw;
            if (gw * 32 < n) {
                p.stw[gw] = stw[tw];
                p.cvw[gw] = cvw[tw];
            }
        }
    }
    __syncthreads();
    const u32 exS = s_pref[0], exC = s_pref[1];

    if (MODE == MODE_Y) {  // segment heads record how many runs precede their segment (DBSCAN.py:88 restarts at 0)
        for (int tw = warp; tw < WR_WORDS; tw += WR_WARPS) {
            const int64_t j = tile_base + tw * 32 + lane;
            const u32 h = hw[HL / 32 + tw];
            if (j < n && ((h >> lane) & 1u)) {
                const u32 rank = exC + pfxC[tw] + __popc(h & lanemask_lt());
                p.gcnt_start[rank] = (int32_t)(exS + pfxS[tw] + __popc(stw[tw] & lanemask_lt()));
                if (p.rank_of) p.rank_of[p.seg_value[j]] = (int32_t)rank;
            }
        }
    }
    if (threadIdx.x == 0 && tile_base + WR_TILE >= n) {  // the last tile publishes the totals
        const u32 totS = exS + s_pref[2], totC = exC + s_pref[3];
        p.totals[0] = totS;
        p.totals[1] = totC;
        if (MODE == MODE_Y) p.gcnt_start[totC] = (int32_t)totS;
    }
}

// ----------------------------------------------------------------------------------------------
// The same kernel specialised for min_pts <= 32 and sorted input (every real TIDDIT run: -l defaults to 3).
// A warp owns 16 consecutive ballot words of the tile and recomputes the one word before them, so run
// starts and "labelled" bits need no other warp: lane t keeps word t in a register, the m-wide OR that marks
// labelled elements is a log-step shift/OR on the 64-bit (previous:current) word pair, and the prefix over
// the 16 words is a warp shuffle scan.  Two block barriers in all (tile staged / carry known).
// ----------------------------------------------------------------------------------------------
//
// TWO_PHASE (the production form): the kernel stops after the bitmasks -- the warps add their two sums in
// shared memory, the last one to arrive stores the tile's sums and everybody exits; wr_tile_scan_kernel turns
// the sums into tile prefixes (in place) and wr_ranks_kernel (Y) writes the per-segment counts.  The chained form keeps a CTA resident until its look-back through the earlier tiles is
// over (a third of all stall samples in the r01_v3 capture, 51 us for 4883 tiles); without the chain a CTA
// lives for one TMA round trip plus 17 ballots.
#ifndef TDT_WR_TWO_PHASE
#define TDT_WR_TWO_PHASE 1
#endif
template <int MODE, bool TWO_PHASE>
__global__ void __launch_bounds__(WR_THREADS) window_runs_small_kernel(const WRParams p) {
    constexpr int HL = 32, WPW = WR_WORDS / WR_WARPS;  // left halo, words per warp (16)
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ u32 s_wtot[WR_WARPS][2], s_wbase[WR_WARPS][2], s_pref[4], s_sum[3];

    const int m = p.m;
    const int HR = (m + 3) & ~3;
    const int HEADW = (HL + WR_TILE + HR) / 32 + 2;
    u32 *keys_s = (u32 *)smem_raw;
    u32 *hw = keys_s + (HL + WR_TILE + HR);

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t n = p.dims->n;
    const int tile = blockIdx.x;  // blockIdx order: see window_runs_kernel
    const int64_t tile_base = (int64_t)tile * WR_TILE;
    if (tile_base >= n) return;
    const int64_t ext_base = tile_base - HL;

    if (threadIdx.x == 0) {
        if (TWO_PHASE) s_sum[0] = s_sum[1] = s_sum[2] = 0u;
        mbar_init(&mbar, 1);
        fence_mbar_init();
        const int64_t n_pad = (n + 3) & ~(int64_t)3;
        const int64_t jlo = ext_base < 0 ? 0 : ext_base;
        int64_t jhi = tile_base + WR_TILE + HR;
        if (jhi > n_pad) jhi = n_pad;
        const uint32_t bytes = (uint32_t)((jhi - jlo) * 4);
        mbar_expect_tx(&mbar, bytes);
        tma_load_1d(keys_s + (jlo - ext_base), p.keys + jlo, bytes, &mbar);
    }
    {
        const int64_t w0 = ext_base >> 5, nwords = (n + 31) >> 5;
        for (int i = threadIdx.x; i < HEADW; i += WR_THREADS) {
            const int64_t w = w0 + i;
            hw[i] = (w >= 0 && w < nwords) ? p.heads[w] : 0u;
        }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);

    // ---- windows: 17 ballot words per warp (its 16 + the one before), word t parked in lane t ----
    const int e_min = tile == 0 ? HL : 0;                                   // positions before element 0
    const int64_t left = n - ext_base;
    const int n_rel = left > (1 << 30) ? (1 << 30) : (int)left;             // shared position of element n
    const u32 mask_m = m >= 32 ? 0xffffffffu : ((1u << m) - 1u);            // heads in (e, e+m]
    const u32 mask_m1 = (1u << (m - 1)) - 1u;                               // heads in (e, e+m-1]
    const u32 eps = p.eps;
    u32 okreg = 0;
#pragma unroll
    for (int t = 0; t <= WPW; t++) {
        const int xw = warp * WPW + t;  // shared word; tile word = xw - 1
        const int e = xw * 32 + lane;
        bool ok = false;
        if (e >= e_min && e + m - 1 < n_rel) {
            const u32 kj = keys_s[e];
            const u32 hb = lane == 31 ? hw[xw + 1] : __funnelshift_r(hw[xw], hw[xw + 1], lane + 1);
            if (MODE == MODE_Y) {
                ok = !(hb & mask_m1) && (keys_s[e + m - 1] - kj < eps);
            } else if (e + m < n_rel && !(hb & mask_m)) {
                ok = keys_s[e + m] - kj < eps;
            } else {
                ok = !(hb & mask_m1) && (keys_s[e + m - 1] - kj < eps);
            }
        }
        const u32 word = __ballot_sync(0xffffffffu, ok);
        if (lane == t) okreg = word;
    }
    // ---- lane t in 1..16: run starts and labelled bits of tile word 16*warp + t - 1 ----------------
    const u32 prev = __shfl_up_sync(0xffffffffu, okreg, 1);
    const bool mine = lane >= 1 && lane <= WPW;
    u32 st = 0, cv = 0, hword = 0;
    if (mine) {
        st = okreg & ~((okreg << 1) | (prev >> 31));
        u64 v = ((u64)okreg << 32) | (u64)prev;   // OR of the m positions ending at every bit
        int span = 1;
        while (span * 2 <= m) {
            v |= v << span;
            span *= 2;
        }
        if (span < m) v |= v << (m - span);
        cv = (u32)(v >> 32);
        hword = hw[warp * WPW + lane];
    }
    const u32 cs = __popc(st), cc = MODE == MODE_X ? __popc(cv) : __popc(hword);
    u32 is = cs, ic = cc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, is, o);
        const u32 b = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) {
            is += a;
            ic += b;
        }
    }
    const int64_t gw = (tile_base >> 5) + warp * WPW + lane - 1;
    if (mine && gw * 32 < n) {
        p.stw[gw] = st;
        p.cvw[gw] = cv;
    }
    if (TWO_PHASE) {
        if (lane == WPW) {
            atomicAdd(&s_sum[0], is);
            atomicAdd(&s_sum[1], ic);
            __threadfence_block();
            if (atomicAdd(&s_sum[2], 1u) == WR_WARPS - 1) {  // every warp's sums are in
                __threadfence_block();
                *(uint2 *)(p.tile_pref + 2 * (int64_t)tile) =
                    make_uint2(*(volatile u32 *)&s_sum[0], *(volatile u32 *)&s_sum[1]);
            }
        }
        return;
    }
    if (lane == WPW) {
        s_wtot[warp][0] = is;
        s_wtot[warp][1] = ic;
    }
    __syncthreads();
    if (warp == 0) {
        u32 ts = lane < WR_WARPS ? s_wtot[lane][0] : 0u, tc = lane < WR_WARPS ? s_wtot[lane][1] : 0u;
        u32 ws = ts, wc = tc;
#pragma unroll
        for (int o = 1; o < WR_WARPS; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, ws, o);
            const u32 b = __shfl_up_sync(0xffffffffu, wc, o);
            if (lane >= o) {
                ws += a;
                wc += b;
            }
        }
        if (lane < WR_WARPS) {
            s_wbase[lane][0] = ws - ts;
            s_wbase[lane][1] = wc - tc;
        }
        const u32 aggS = __shfl_sync(0xffffffffu, ws, WR_WARPS - 1);
        const u32 aggC = __shfl_sync(0xffffffffu, wc, WR_WARPS - 1);
        u32 exS, exC;
        lookback(p.status, tile, aggS, aggC, exS, exC);
        if (lane == 0) {
            s_pref[0] = exS;
            s_pref[1] = exC;
            s_pref[2] = aggS;
            s_pref[3] = aggC;
            p.tile_pref[2 * (int64_t)tile] = exS;
            p.tile_pref[2 * (int64_t)tile + 1] = exC;
        }
    }
    if (MODE == MODE_X && !(tile_base + WR_TILE >= n)) return;  // X tiles are done; the last one adds the totals
    __syncthreads();
    const u32 exS = s_pref[0], exC = s_pref[1];
    if (MODE == MODE_Y && mine) {
        // segment heads record how many runs precede their segment (DBSCAN.py:88 restarts the count at 0)
        const u32 baseS = exS + s_wbase[warp][0] + (is - cs), baseC = exC + s_wbase[warp][1] + (ic - cc);
        u32 h = hword;
        while (h) {
            const int b = __ffs(h) - 1;
            const u32 below = (1u << b) - 1u;
            const u32 rank = baseC + __popc(hword & below);
            p.gcnt_start[rank] = (int32_t)(baseS + __popc(st & below));
            if (p.rank_of) p.rank_of[p.seg_value[gw * 32 + b]] = (int32_t)rank;
            h &= h - 1;
        }
    }
    if (threadIdx.x == 0 && tile_base + WR_TILE >= n) {  // the last tile publishes the totals
        const u32 totS = exS + s_pref[2], totC = exC + s_pref[3];
        p.totals[0] = totS;
        p.totals[1] = totC;
        if (MODE == MODE_Y) p.gcnt_start[totC] = (int32_t)totS;
    }
}

// ----------------------------------------------------------------------------------------------
// Second phase of the two-phase range query: tile prefixes from the per-warp sums (one CTA; the 30X set has
// 4883 tiles, 64 bytes each), the totals, and -- y axis -- the per-segment run counts.
// ----------------------------------------------------------------------------------------------
constexpr int TS_THREADS = 1024, TS_PER = 8;   // one round for up to 8192 tiles (33.5 M elements)

// in place: tile_pref[tile] = (the tile's own sums) -> (sums over the earlier tiles); totals = sums over all tiles
__global__ void __launch_bounds__(TS_THREADS) wr_tile_scan_kernel(const Dims *dims, u32 *tile_pref, u32 *totals,
                                                                  int32_t *gcnt_start) {
    __shared__ u32 s_part[TS_THREADS / 32][2];
    const int64_t n = dims->n;
    const int n_tiles = (int)((n + WR_TILE - 1) / WR_TILE);
    if (n_tiles == 0) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint2 *tp = (uint2 *)tile_pref;
    u32 carryA = 0, carryB = 0;
    for (int base = 0; base < n_tiles; base += TS_THREADS * TS_PER) {
        const int t0 = base + threadIdx.x * TS_PER;
        uint2 v[TS_PER];
#pragma unroll
        for (int k = 0; k < TS_PER; k++) v[k] = t0 + k < n_tiles ? tp[t0 + k] : make_uint2(0u, 0u);
        u32 sumA = 0, sumB = 0;
#pragma unroll
        for (int k = 0; k < TS_PER; k++) {
            sumA += v[k].x;
            sumB += v[k].y;
        }
        u32 incA = sumA, incB = sumB;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, incA, o);
            const u32 b = __shfl_up_sync(0xffffffffu, incB, o);
            if (lane >= o) {
                incA += a;
                incB += b;
            }
        }
        if (lane == 31) {
            s_part[warp][0] = incA;
            s_part[warp][1] = incB;
        }
        __syncthreads();
        u32 baseA = carryA + incA - sumA, baseB = carryB + incB - sumB;
#pragma unroll 8
        for (int w = 0; w < TS_THREADS / 32; w++) {
            const u32 a = s_part[w][0], b = s_part[w][1];
            if (w < warp) {
                baseA += a;
                baseB += b;
            }
            carryA += a;
            carryB += b;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TS_PER; k++) {
            if (t0 + k < n_tiles) tp[t0 + k] = make_uint2(baseA, baseB);
            baseA += v[k].x;
            baseB += v[k].y;
        }
    }
    if (threadIdx.x == 0) {
        totals[0] = carryA;
        totals[1] = carryB;
        if (gcnt_start) gcnt_start[carryB] = (int32_t)carryA;
    }
}

// y axis: every segment head records how many runs precede its segment (DBSCAN.py:88 restarts the count at 0)
// and, optionally, the rank of its segment.  One CTA per tile: 128 threads, one word each.
__global__ void __launch_bounds__(WR_WORDS) wr_ranks_kernel(const u32 *__restrict__ heads, const u32 *__restrict__ stw,
                                                            const u32 *__restrict__ tile_pref, const Dims *dims,
                                                            int32_t *__restrict__ gcnt_start, int32_t *rank_of,
                                                            const int32_t *__restrict__ seg_value) {
    __shared__ u32 s_part[WR_WORDS / 32][2];
    const int64_t n = dims->n;
    const int64_t gw = (int64_t)blockIdx.x * WR_WORDS + threadIdx.x;
    if ((int64_t)blockIdx.x * WR_TILE >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const bool live = gw * 32 < n;
    const u32 hword = live ? heads[gw] : 0u, st = live ? stw[gw] : 0u;
    const u32 cs = __popc(st), cc = __popc(hword);
    u32 is = cs, ic = cc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, is, o);
        const u32 b = __shfl_up_sync(0xffffffffu, ic, o);
        if (lane >= o) {
            is += a;
            ic += b;
        }
    }
    if (lane == 31) {
        s_part[warp][0] = is;
        s_part[warp][1] = ic;
    }
    __syncthreads();
    u32 baseS = tile_pref[2 * (int64_t)blockIdx.x] + (is - cs), baseC = tile_pref[2 * (int64_t)blockIdx.x + 1] + (ic - cc);
    for (int w = 0; w < warp; w++) {
        baseS += s_part[w][0];
        baseC += s_part[w][1];
    }
    u32 h = hword;
    while (h) {
        const int b = __ffs(h) - 1;
        const u32 below = (1u << b) - 1u;
        const u32 rank = baseC + __popc(hword & below);
        gcnt_start[rank] = (int32_t)(baseS + __popc(st & below));
        if (rank_of) rank_of[seg_value[gw * 32 + b]] = (int32_t)rank;
        h &= h - 1;
    }
}

// ----------------------------------------------------------------------------------------------
// small helpers around the range-query kernel
// ----------------------------------------------------------------------------------------------
__global__ void heads_from_offsets_kernel(const int64_t *__restrict__ off, const Dims *dims, u32 *heads) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= dims->nseg) return;
    const int64_t q = off[s];
    if (off[s + 1] > q) atomicOr(heads + (q >> 5), 1u << (q & 31));
}

__global__ void set_dims_kernel(Dims *dims, int64_t n, int64_t nseg, int64_t *off2) {
    dims->n = n;
    dims->nseg = nseg;
    if (off2) {  // the single-segment offsets {0, n}
        off2[0] = 0;
        off2[1] = n;
    }
}

// run starts before element q, from the start-bit words and the per-tile carries (one warp per query)
__device__ __forceinline__ u32 runs_before(const u32 *stw, const u32 *tile_pref, int64_t q, int lane) {
    const int64_t tile = q / WR_TILE;
    const int64_t w0 = tile * WR_WORDS, wq = q >> 5;
    u32 c = 0;
    for (int64_t w = w0 + lane; w <= wq; w += 32) {
        u32 bits = stw[w];
        if (w == wq) bits &= (q & 31) ? (0xffffffffu >> (32 - (q & 31))) : 0u;
        c += __popc(bits);
    }
    return tile_pref[2 * tile] + warp_sum(c);
}

// gfirst[p] = number of x-runs that start before pair p (x ids of pair p are gfirst[p] .. gfirst[p+1]-1)
__global__ void pair_first_run_kernel(const int64_t *__restrict__ seg_off, int P, const Dims *dims,
                                      const u32 *__restrict__ stw, const u32 *__restrict__ tile_pref,
                                      const u32 *__restrict__ totals, int32_t *__restrict__ gfirst) {
    const int64_t wid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (wid > P) return;
    const int64_t q = seg_off[wid];
    const u32 v = q >= dims->n ? totals[0] : runs_before(stw, tile_pref, q, lane);
    if (lane == 0) gfirst[wid] = (int32_t)v;
}

// ----------------------------------------------------------------------------------------------
// pack_y: the labelled signals, in posA order, become the input of the posB sort: ykey = posB, yval =
// insertion index, gx = x-run; goff[g] = where x-run g starts (x-runs are contiguous in posA order).
// ----------------------------------------------------------------------------------------------
struct PackParams {
    const u32 *stw, *cvw, *tile_pref, *totals;
    const Dims *dims;
    const int32_t *xv;     // insertion index per sorted position (nullptr: identity)
    const int32_t *posB;
    int32_t max_pos;
    u32 *ykey;
    int32_t *yval, *gx;
    int64_t *goff;
    Dims *dims_y;
    int *err;
};

__global__ void __launch_bounds__(WR_THREADS) pack_y_kernel(const PackParams p) {
    __shared__ u32 stw[WR_WORDS], cvw[WR_WORDS], pfxS[WR_WORDS], pfxC[WR_WORDS];
    const int64_t n = p.dims->n;
    const int64_t tile = blockIdx.x;
    const int64_t tile_base = tile * WR_TILE;
    if (tile_base >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x < WR_WORDS) {
        const int64_t gw = tile * WR_WORDS + threadIdx.x;
        const bool in = gw * 32 < n;
        stw[threadIdx.x] = in ? p.stw[gw] : 0u;
        cvw[threadIdx.x] = in ? p.cvw[gw] : 0u;
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int PER = WR_WORDS / 32;
        u32 s[PER], c[PER], ts = 0, tc = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            s[k] = __popc(stw[lane * PER + k]);
            c[k] = __popc(cvw[lane * PER + k]);
            ts += s[k];
            tc += c[k];
        }
        u32 is = ts, ic = tc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, is, o);
            const u32 b = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) {
                is += a;
                ic += b;
            }
        }
        u32 es = is - ts, ec = ic - tc;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            pfxS[lane * PER + k] = es;
            pfxC[lane * PER + k] = ec;
            es += s[k];
            ec += c[k];
        }
    }
    __syncthreads();
    const u32 exS = p.tile_pref[2 * tile], exC = p.tile_pref[2 * tile + 1];
#pragma unroll 4
    for (int tw = warp; tw < WR_WORDS; tw += WR_WARPS) {
        const int64_t j = tile_base + tw * 32 + lane;
        const u32 st = stw[tw], cv = cvw[tw];
        if (j < n && ((cv >> lane) & 1u)) {
            const u32 pos = exC + pfxC[tw] + __popc(cv & lanemask_lt());
            const u32 g = exS + pfxS[tw] + __popc(st & lanemask_le()) - 1u;
            const int32_t idx = p.xv ? p.xv[j] : (int32_t)j;
            const int32_t yb = p.posB[idx];
            if (yb < 0 || yb > p.max_pos) atomicMax(p.err, ERR_RANGE_B);
            p.ykey[pos] = (u32)yb;
            p.yval[pos] = idx;
            p.gx[pos] = (int32_t)g;
            if ((st >> lane) & 1u) p.goff[g] = (int64_t)pos;
        }
    }
    if (threadIdx.x == 0 && tile_base + WR_TILE >= n) {
        p.goff[p.totals[0]] = (int64_t)p.totals[1];
        p.dims_y->n = (int64_t)p.totals[1];
        p.dims_y->nseg = (int64_t)p.totals[0];
    }
}

// ----------------------------------------------------------------------------------------------
// DBSCAN.py:121-122: after x-run g the running id grows by max(sub-runs(g) - 1, 0).
// X[g] = sum over g' < g, single-pass look-back scan, 2048 x-runs per CTA.
// ----------------------------------------------------------------------------------------------
constexpr int GS_THREADS = 256;
constexpr int GS_ITEMS = 8;
constexpr int GS_TILE = GS_THREADS * GS_ITEMS;

__global__ void __launch_bounds__(GS_THREADS) group_extra_scan_kernel(const int32_t *__restrict__ gcnt_start,
                                                                      const Dims *dims, int32_t *__restrict__ X,
                                                                      u64 *status, u32 *ticket) {
    __shared__ int s_tile;
    __shared__ u32 s_warp[GS_THREADS / 32];
    __shared__ u32 s_ex;
    const int64_t G = dims->nseg;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    if ((int64_t)tile * GS_TILE >= G) {
        if (G == 0 && tile == 0 && threadIdx.x == 0) X[0] = 0;
        return;
    }
    const int64_t g0 = (int64_t)tile * GS_TILE + (int64_t)threadIdx.x * GS_ITEMS;
    u32 v[GS_ITEMS], tsum = 0;
    int32_t prev = g0 < G ? gcnt_start[g0] : 0;
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t g = g0 + k;
        u32 extra = 0;
        if (g < G) {
            const int32_t next = gcnt_start[g + 1];
            const int32_t subs = next - prev;
            extra = subs > 1 ? (u32)(subs - 1) : 0u;
            prev = next;
        }
        v[k] = extra;
        tsum += extra;
    }
    u32 inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < GS_THREADS / 32 ? s_warp[lane] : 0u;
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += a;
        }
        const u32 agg = __shfl_sync(0xffffffffu, wi, 31);
        if (lane < GS_THREADS / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
        u32 exA, exB;
        lookback(status, tile, agg, 0u, exA, exB);
        if (lane == 0) s_ex = exA;
    }
    __syncthreads();
    u32 run = s_ex + s_warp[warp] + (inc - tsum);
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t g = g0 + k;
        if (g < G) X[g] = (int32_t)run;
        run += v[k];
        if (g == G - 1) X[G] = (int32_t)run;
    }
}

// DBSCAN.py:113-117 per x-run g of pair s (x ids of the pair are gfirst[s] .. gfirst[s+1]-1):
//   sub-run 1 keeps the pair-local x id            base1 = g - gfirst[s]
//   sub-run k >= 2 gets  k + cluster_id - 1        base2 + k, base2 = nx + (X[g] - X[gfirst[s]]) - 2
__global__ void group_bases_kernel(const int32_t *__restrict__ gfirst, int P, const int32_t *__restrict__ X,
                                   const Dims *dims, int32_t *__restrict__ base1, int32_t *__restrict__ base2) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= dims->nseg) return;
    int lo = 0, hi = P;  // the pair of x-run g: largest s with gfirst[s] <= g (pairs without runs repeat a value)
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if ((int64_t)gfirst[mid] <= g) lo = mid; else hi = mid;
    }
    const int32_t gf = gfirst[lo];
    const int32_t nx = gfirst[lo + 1] - gf;
    base1[g] = (int32_t)g - gf;
    base2[g] = nx + (X[g] - X[gf]) - 2;
}

struct FinalParams {
    const u32 *stw, *cvw, *tile_pref;
    const Dims *dims;                 // y dims
    const int32_t *yval, *gx, *gcnt_start, *base1, *base2;
    const int32_t *rank_of;           // plain mode: caller id -> segment rank (nullptr: gx is the rank)
    int32_t plain_cluster_id;         // plain mode: ids continue after this one
    const int32_t *X;
    int32_t *labels_out;
    int32_t *cluster_id_out;          // plain mode
};

// final ids (DBSCAN.py:112-119), scattered to insertion order; noise stays at the -1 the output was filled with
// 4 CTAs per SM (64 registers): measured 0.117 ms for the 30X set, 0.134 ms uncapped (80 registers, 3 CTAs), 0.160 ms
// at 6; the one-access-at-a-time form this replaces took 0.132 ms.  The same restructuring made pack_y slower (0.108 ->
// 0.13 ms: it loads the insertion index of every element instead of the labelled ones only), so pack_y keeps its loop.
template <bool PLAIN>
__global__ void __launch_bounds__(WR_THREADS, 4) final_labels_kernel(const FinalParams p) {
    __shared__ u32 stw[WR_WORDS], cvw[WR_WORDS], pfxS[WR_WORDS];
    const int64_t n = p.dims->n;
    const int64_t tile = blockIdx.x;
    const int64_t tile_base = tile * WR_TILE;
    if (PLAIN && blockIdx.x == 0 && threadIdx.x == 0) *p.cluster_id_out = p.plain_cluster_id + p.X[p.dims->nseg];
    if (tile_base >= n) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int ROUNDS = WR_WORDS / WR_WARPS;
    // coalesced loads first (in flight during the bitmask bookkeeping), then the per-run gathers, then the scatter
    int32_t gv[ROUNDS], dst[ROUNDS];
#pragma unroll
    for (int k = 0; k < ROUNDS; k++) {
        const int64_t j = tile_base + (warp + k * WR_WARPS) * 32 + lane;
        gv[k] = j < n ? p.gx[j] : 0;
        dst[k] = j < n ? p.yval[j] : 0;
    }
    if (threadIdx.x < WR_WORDS) {
        const int64_t gw = tile * WR_WORDS + threadIdx.x;
        const bool in = gw * 32 < n;
        stw[threadIdx.x] = in ? p.stw[gw] : 0u;
        cvw[threadIdx.x] = in ? p.cvw[gw] : 0u;
    }
    __syncthreads();
    if (warp == 0) {
        constexpr int PER = WR_WORDS / 32;
        u32 s[PER], ts = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            s[k] = __popc(stw[lane * PER + k]);
            ts += s[k];
        }
        u32 is = ts;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, is, o);
            if (lane >= o) is += a;
        }
        u32 es = is - ts;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            pfxS[lane * PER + k] = es;
            es += s[k];
        }
    }
    __syncthreads();
    const u32 exS = p.tile_pref[2 * tile];
    int32_t label[ROUNDS];
#pragma unroll
    for (int k = 0; k < ROUNDS; k++) {
        const int tw = warp + k * WR_WARPS;
        const int64_t j = tile_base + tw * 32 + lane;
        const u32 st = stw[tw], cv = cvw[tw];
        label[k] = -1;
        if (j < n && ((cv >> lane) & 1u)) {
            const int32_t starts_incl = (int32_t)(exS + pfxS[tw] + __popc(st & lanemask_le()));
            const int32_t g = PLAIN ? p.rank_of[gv[k]] : gv[k];
            const int32_t sub = starts_incl - p.gcnt_start[g];
            if (PLAIN) label[k] = sub == 1 ? gv[k] : p.plain_cluster_id + p.X[g] + sub - 1;
            else label[k] = sub == 1 ? p.base1[g] : p.base2[g] + sub;
        }
    }
#pragma unroll
    for (int k = 0; k < ROUNDS; k++) {
        const int tw = warp + k * WR_WARPS;
        const int64_t j = tile_base + tw * 32 + lane;
        if (j < n && ((cvw[tw] >> lane) & 1u)) p.labels_out[dst[k]] = label[k];
    }
}

// x as given (no sort): the stand-alone DBSCAN.py entry points cluster the caller's order
__global__ void pack_plain_kernel(const int32_t *__restrict__ x, int64_t n, int32_t max_pos, u32 *__restrict__ keys,
                                  int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int32_t a = x[j];
        if (a < 0 || a > max_pos) atomicMax(err, ERR_RANGE_A);
        keys[j] = (u32)a;
    }
}

// x-pass labels in the caller's order (DBSCAN.py:33-64 return value): one warp per 32 elements
__global__ void __launch_bounds__(WR_THREADS) expand_xlabels_kernel(const u32 *__restrict__ stw_g,
                                                                    const u32 *__restrict__ cvw_g,
                                                                    const u32 *__restrict__ tile_pref, int64_t n,
                                                                    const u32 *__restrict__ totals,
                                                                    int32_t *__restrict__ labels_out,
                                                                    int32_t *last_id_out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * WR_THREADS + threadIdx.x) >> 5;
    const int64_t warps = ((int64_t)gridDim.x * WR_THREADS) >> 5;
    if (blockIdx.x == 0 && threadIdx.x == 0 && last_id_out) *last_id_out = (int32_t)totals[0] - 1;
    for (int64_t gw = warp_global; gw * 32 < n; gw += warps) {
        const int64_t j = gw * 32 + lane;
        const u32 before = runs_before(stw_g, tile_pref, gw * 32, lane);
        const u32 st = stw_g[gw], cv = cvw_g[gw];
        if (j < n) labels_out[j] = ((cv >> lane) & 1u) ? (int32_t)(before + __popc(st & lanemask_le())) - 1 : -1;
    }
}

// stand-alone y-pass (DBSCAN.py:66-123 on caller-supplied ids): sort keys = the id, noise behind every real id
__global__ void ypass_label_keys_kernel(const int32_t *__restrict__ labels, int64_t n, int32_t cluster_id,
                                        u32 *__restrict__ keys, int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        int32_t l = labels[j];
        if (l < -1 || l > cluster_id) {
            atomicMax(err, ERR_PAIR);
            l = -1;
        }
        keys[j] = l < 0 ? (u32)(cluster_id + 1) : (u32)l;
    }
}

// sorted ids -> goff[id] = first position of the id (ids without members are filled in afterwards),
// ykey = y gathered, gx = id
__global__ void ypass_offsets_kernel(const u32 *__restrict__ ids, const int32_t *__restrict__ idx, int64_t n,
                                     const int32_t *__restrict__ y, int32_t max_pos, int64_t *__restrict__ goff,
                                     u32 *__restrict__ ykey, int32_t *__restrict__ gx, int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const u32 v = ids[j];
        if (j == 0 || ids[j - 1] != v) goff[v] = j;
        const int32_t b = y[idx[j]];
        if (b < 0 || b > max_pos) atomicMax(err, ERR_RANGE_B);
        ykey[j] = (u32)b;
        gx[j] = (int32_t)v;
    }
}

// goff has n_ids + 2 entries: ids 0..n_ids-1, the noise id, the end; unset (-1) ones take the next set one
__global__ void ypass_fill_offsets_kernel(int64_t *goff, int64_t n_ids, int64_t n) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g > n_ids + 1) return;
    if (g == n_ids + 1) {
        goff[g] = n;
        return;
    }
    if (goff[g] >= 0) return;
    int64_t q = g + 1;
    while (q <= n_ids && goff[q] < 0) q++;
    goff[g] = q <= n_ids ? goff[q] : n;  // set entries are only read, unset ones only written
}

// stand-alone y-pass: ids without members were skipped by the segment ranking; continue with the ranked count
__global__ void set_nseg_from_totals_kernel(Dims *dims_y, const u32 *totals_y) { dims_y->nseg = (int64_t)totals_y[1]; }

__global__ void ypass_dims_kernel(const int64_t *goff, int64_t n_ids, Dims *dims_y) {
    dims_y->n = goff[n_ids];  // where the noise id starts = number of clustered elements
    dims_y->nseg = n_ids;
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
struct Small {  // device scalars, one 256-byte block, zeroed per call
    u32 ticket[4];
    u32 totals_x[2];
    u32 totals_y[2];
    int err;
    int32_t pad;
    Dims dims_x, dims_y;
    int64_t off2[2];
};

struct ClusterPlan {
    int64_t n, n_pad, tiles, words, gmax;
    int P;
    size_t arr, bits, status, tile_pref, gfirst, g32, g64, sort;
    size_t total;
};

static int64_t wr_tiles(int64_t n) { return (n + WR_TILE - 1) / WR_TILE; }
static size_t al256(size_t b) { return (b + 255) & ~(size_t)255; }

static ClusterPlan make_plan(int64_t n, int32_t P) {
    ClusterPlan pl;
    pl.n = n;
    pl.P = P;
    pl.n_pad = (n + 63) & ~(int64_t)63;
    pl.tiles = wr_tiles(n) + 1;
    pl.words = pl.tiles * WR_WORDS + 64;
    pl.gmax = n / 2 + 4;
    pl.arr = al256((size_t)pl.n_pad * 4 + 256);
    pl.bits = al256((size_t)pl.words * 4);
    pl.status = al256((size_t)(pl.tiles + 64) * 8);
    pl.tile_pref = al256((size_t)pl.tiles * 8);
    pl.gfirst = al256(((size_t)P + 2) * 4);
    pl.g32 = al256((size_t)pl.gmax * 4);
    pl.g64 = al256((size_t)pl.gmax * 8);
    const size_t s1 = segsort_temp_bytes(n, P > 0 ? P : 1), s2 = segsort_temp_bytes(n, pl.gmax);
    pl.sort = al256(s1 > s2 ? s1 : s2);
    pl.total = 7 * pl.arr + 6 * pl.bits + 3 * pl.status + 2 * pl.tile_pref + 256 + pl.gfirst + 4 * pl.g32 + pl.g64 +
               pl.sort + 4096;
    return pl;
}

struct Buffers {
    u32 *xs, *tmpK, *ykey;
    int32_t *xv, *tmpV, *yval, *gx;
    u32 *headsX, *stwX, *cvwX, *headsY, *stwY, *cvwY;
    u64 *statusX, *statusY, *statusG;
    u32 *tprefX, *tprefY;
    Small *small;
    int32_t *gfirst, *gcnt_start, *X, *base1, *base2;
    int64_t *goff;
    void *sort_temp;
    bool ok;
};

static Buffers carve(const ClusterPlan &pl, void *ws, size_t ws_bytes) {
    Arena ar(ws, ws_bytes);
    Buffers b;
    b.xs = (u32 *)ar.take<char>(pl.arr);
    b.xv = (int32_t *)ar.take<char>(pl.arr);
    b.tmpK = (u32 *)ar.take<char>(pl.arr);
    b.tmpV = (int32_t *)ar.take<char>(pl.arr);
    b.ykey = (u32 *)ar.take<char>(pl.arr);
    b.yval = (int32_t *)ar.take<char>(pl.arr);
    b.gx = (int32_t *)ar.take<char>(pl.arr);
    // everything from headsX up to and including `small` is zeroed by one memset (zero_span)
    b.headsX = (u32 *)ar.take<char>(pl.bits);
    b.headsY = (u32 *)ar.take<char>(pl.bits);
    b.statusX = (u64 *)ar.take<char>(pl.status);
    b.statusY = (u64 *)ar.take<char>(pl.status);
    b.statusG = (u64 *)ar.take<char>(pl.status);
    b.small = (Small *)ar.take<char>(256);
    b.stwX = (u32 *)ar.take<char>(pl.bits);
    b.cvwX = (u32 *)ar.take<char>(pl.bits);
    b.stwY = (u32 *)ar.take<char>(pl.bits);
    b.cvwY = (u32 *)ar.take<char>(pl.bits);
    b.tprefX = (u32 *)ar.take<char>(pl.tile_pref);
    b.tprefY = (u32 *)ar.take<char>(pl.tile_pref);
    b.gfirst = (int32_t *)ar.take<char>(pl.gfirst);
    b.gcnt_start = (int32_t *)ar.take<char>(pl.g32);
    b.X = (int32_t *)ar.take<char>(pl.g32);
    b.base1 = (int32_t *)ar.take<char>(pl.g32);
    b.base2 = (int32_t *)ar.take<char>(pl.g32);
    b.goff = (int64_t *)ar.take<char>(pl.g64);
    b.sort_temp = ar.take<char>(pl.sort);
    b.ok = b.sort_temp != nullptr;
    return b;
}

static size_t zero_span(const ClusterPlan &pl) { return 2 * pl.bits + 3 * pl.status + 256; }

static int grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

// second phase (p.two_phase and TDT_WR_TWO_PHASE): tile prefixes, totals, y-axis segment counts
template <int MODE>
static int launch_window_runs_finish(const WRParams &p, int64_t n_max, cudaStream_t st) {
#if TDT_WR_TWO_PHASE
    if (!p.two_phase || p.m > 32) return TDT_OK;
    TDT_LAUNCH(wr_tile_scan_kernel, 1, TS_THREADS, 0, st, p.dims, p.tile_pref, p.totals,
               MODE == MODE_Y ? p.gcnt_start : nullptr);
    if (MODE == MODE_Y)
        TDT_LAUNCH(wr_ranks_kernel, (unsigned)wr_tiles(n_max), WR_WORDS, 0, st, p.heads, p.stw, p.tile_pref, p.dims,
                   p.gcnt_start, p.rank_of, p.seg_value);
#endif
    return TDT_OK;
}

template <int MODE, bool GENERAL>
static int launch_window_runs(const WRParams &p, int64_t n_max, cudaStream_t st) {
    if (!GENERAL && p.m <= 32) {
        const size_t HR = ((size_t)p.m + 3) & ~(size_t)3;
        const size_t smem = (32 + WR_TILE + HR) * 4 + ((32 + WR_TILE + HR) / 32 + 2) * 4;
#if TDT_WR_TWO_PHASE
        if (p.two_phase) {
            TDT_LAUNCH((window_runs_small_kernel<MODE, true>), (unsigned)wr_tiles(n_max), WR_THREADS, smem, st, p);
            return TDT_OK;
        }
#endif
        TDT_LAUNCH((window_runs_small_kernel<MODE, false>), (unsigned)wr_tiles(n_max), WR_THREADS, smem, st, p);
        return TDT_OK;
    }
    const size_t smem = wr_smem_bytes(p.m);
    if (smem > 48 * 1024)
        TDT_CUDA(cudaFuncSetAttribute(window_runs_kernel<MODE, GENERAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
    TDT_LAUNCH((window_runs_kernel<MODE, GENERAL>), (unsigned)wr_tiles(n_max), WR_THREADS, smem, st, p);
    return TDT_OK;
}

static int check_common(int64_t n, int32_t min_pts, void *ws, size_t ws_bytes, size_t need) {
    if (n < 0) return fail(TDT_E_ARG, "n = %lld is negative", (long long)n);
    if (n > 2000000000LL) return fail(TDT_E_ARG, "n = %lld exceeds the 2e9 signals one call supports", (long long)n);
    if (min_pts < 2)
        return fail(TDT_E_ARG, "min_pts = %d: the reference raises ValueError (max() of an empty window) for m < 2",
                    min_pts);
    if (min_pts > WR_MAX_M) return fail(TDT_E_ARG, "min_pts = %d exceeds the supported maximum %d", min_pts, WR_MAX_M);
    if (n > 0 && (ws == nullptr || ws_bytes < need))
        return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, need);
    return TDT_OK;
}

static int err_to_code(int e) {
    switch (e) {
        case ERR_NONE: return TDT_OK;
        case ERR_RANGE_A: return fail(TDT_E_RANGE, "a posA / x coordinate is negative or above max_pos");
        case ERR_RANGE_B: return fail(TDT_E_RANGE, "a posB / y coordinate is negative or above max_pos");
        case SS_ERR_INTERNAL: return fail(TDT_E_CUDA, "segmented sort: a work list overflowed (internal error)");
        default: return fail(TDT_E_RANGE, "a pair / cluster id is outside its range");
    }
}

static int finish(Buffers &b, cudaStream_t st) {
    Small h;
    TDT_CUDA(cudaMemcpyAsync(&h, b.small, sizeof(Small), cudaMemcpyDeviceToHost, st));
    TDT_CUDA(cudaStreamSynchronize(st));
    return err_to_code(h.err);
}

// asynchronous callers: fold this call's error flag into the caller's device-side status word
__global__ void accumulate_status_kernel(const Small *small, int32_t *status_accum) {
    if (small->err) atomicMax(status_accum, (int32_t)small->err);
}

static int pos_bits(int32_t max_pos) { return bit_width_u32(max_pos > 0 ? (uint32_t)max_pos : 0x7fffffffu); }

// y-pass on the compacted (posB, insertion index, x-run) triples in ykey/yval/gx with offsets goff and
// sizes dims_y; writes the final ids.  PLAIN: caller ids (stand-alone y-pass).
template <bool PLAIN>
static int run_ypass(const ClusterPlan &pl, Buffers &b, int32_t eps, int32_t m, int key_bits, int P,
                     int32_t plain_cluster_id, int32_t *rank_of, int32_t *labels_out, int32_t *cluster_id_out,
                     cudaStream_t st) {
    const int64_t n = pl.n;
    const int64_t gmax = PLAIN ? (int64_t)plain_cluster_id + 1 : pl.gmax;
    {
        ProfScope ps("sort_y", st);
        const bool fused_heads = segsort_want_heads(b.headsY);   // the sort's classify kernel walks the segments anyway
        int rc = segsort_pairs(b.ykey, b.yval, b.xs, b.xv, b.tmpK, b.tmpV, b.goff, (const int64_t *)&b.small->dims_y,
                               b.gx, n, gmax, key_bits, b.sort_temp, pl.sort, &b.small->err, st);
        if (rc) return rc;
        if (!fused_heads)
            TDT_LAUNCH(heads_from_offsets_kernel, (unsigned)((gmax + 255) / 256), 256, 0, st, b.goff, &b.small->dims_y,
                       b.headsY);
    }
    WRParams p = {};
    p.keys = b.xs;
    p.heads = b.headsY;
    p.dims = &b.small->dims_y;
    p.m = m;
    p.eps = eps > 0 ? (u32)eps : 0u;
    p.status = b.statusY;
    p.stw = b.stwY;
    p.cvw = b.cvwY;
    p.tile_pref = b.tprefY;
    p.totals = b.small->totals_y;
    p.gcnt_start = b.gcnt_start;
    p.rank_of = PLAIN ? rank_of : nullptr;
    p.seg_value = b.gx;
    p.two_phase = 1;
    {
        ProfScope ps("window_runs_y", st);
        int rc = launch_window_runs<MODE_Y, false>(p, n, st);
        if (rc) return rc;
    }
    {
        ProfScope ps("tile_scan_y", st);
        int rc = launch_window_runs_finish<MODE_Y>(p, n, st);
        if (rc) return rc;
        if (PLAIN) TDT_LAUNCH(set_nseg_from_totals_kernel, 1, 1, 0, st, &b.small->dims_y, b.small->totals_y);
    }
    {
        ProfScope ps("group_ids", st);
        const unsigned gs_tiles = (unsigned)((gmax + GS_TILE - 1) / GS_TILE);
        TDT_LAUNCH(group_extra_scan_kernel, gs_tiles, GS_THREADS, 0, st, b.gcnt_start, &b.small->dims_y, b.X,
                   b.statusG, &b.small->ticket[2]);
        if (!PLAIN)
            TDT_LAUNCH(group_bases_kernel, (unsigned)((gmax + 255) / 256), 256, 0, st, b.gfirst, P, b.X,
                       &b.small->dims_y, b.base1, b.base2);
    }
    FinalParams f = {};
    f.stw = b.stwY;
    f.cvw = b.cvwY;
    f.tile_pref = b.tprefY;
    f.dims = &b.small->dims_y;
    f.yval = b.xv;
    f.gx = b.gx;
    f.gcnt_start = b.gcnt_start;
    f.base1 = b.base1;
    f.base2 = b.base2;
    f.rank_of = rank_of;
    f.plain_cluster_id = plain_cluster_id;
    f.X = b.X;
    f.labels_out = labels_out;
    f.cluster_id_out = cluster_id_out;
    ProfScope ps("final_labels", st);
    TDT_LAUNCH(final_labels_kernel<PLAIN>, (unsigned)wr_tiles(n), WR_THREADS, 0, st, f);
    return TDT_OK;
}

// seg_off: device offsets of the P pairs, or nullptr with presorted_plain (one segment, no sort)
static int cluster_impl(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, int64_t n, int32_t P,
                        int32_t eps, int32_t m, int32_t max_pos, int32_t *labels_out, void *ws, size_t ws_bytes,
                        cudaStream_t st, bool presorted_plain, int32_t *status_accum = nullptr) {
    const ClusterPlan pl = make_plan(n, P);
    Buffers b = carve(pl, ws, ws_bytes);
    if (!b.ok) return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);
    const int key_bits = pos_bits(max_pos);

    // labels start as -1 (noise).  (Issued right before the scatter instead -- so that final_labels stores into lines
    // that are still in L2 -- measured the same: 0.9946 vs 0.9906 ms for the 30X step.)
    TDT_CUDA(cudaMemsetAsync(labels_out, 0xff, (size_t)n * 4, st));
    TDT_CUDA(cudaMemsetAsync(b.headsX, 0, zero_span(pl), st));
    TDT_LAUNCH(set_dims_kernel, 1, 1, 0, st, &b.small->dims_x, n, (int64_t)P, b.small->off2);

    const int32_t *xv = nullptr;
    bool heads_done = false;
    if (presorted_plain) {
        // DBSCAN.main on the caller's order: no sort, identity permutation, one segment
        TDT_LAUNCH(pack_plain_kernel, grid_for(n, 256), 256, 0, st, posA, n, max_pos, b.xs, &b.small->err);
        seg_off = b.small->off2;
    } else {
        ProfScope ps("sort_x", st);
        heads_done = segsort_want_heads(b.headsX);
        // posA as unsigned keys: a negative coordinate has bit 31 set and trips the key-range check
        int rc = segsort_pairs((const u32 *)posA, nullptr, b.xs, b.xv, b.tmpK, b.tmpV, seg_off,
                               (const int64_t *)&b.small->dims_x, nullptr, n, P, key_bits, b.sort_temp, pl.sort,
                               &b.small->err, st);
        if (rc) return rc;
        xv = b.xv;
    }
    if (!heads_done)
        TDT_LAUNCH(heads_from_offsets_kernel, (unsigned)((P + 255) / 256), 256, 0, st, seg_off, &b.small->dims_x, b.headsX);

    WRParams p = {};
    p.keys = b.xs;
    p.heads = b.headsX;
    p.dims = &b.small->dims_x;
    p.m = m;
    p.eps = eps > 0 ? (u32)eps : 0u;
    p.status = b.statusX;
    p.stw = b.stwX;
    p.cvw = b.cvwX;
    p.tile_pref = b.tprefX;
    p.totals = b.small->totals_x;
    p.two_phase = presorted_plain ? 0 : 1;
    {
        ProfScope ps("window_runs_x", st);
        int rc = presorted_plain ? launch_window_runs<MODE_X, true>(p, n, st)
                                 : launch_window_runs<MODE_X, false>(p, n, st);
        if (rc) return rc;
    }
    {
        ProfScope ps("tile_scan_x", st);
        int rc = launch_window_runs_finish<MODE_X>(p, n, st);
        if (rc) return rc;
    }
    {
        ProfScope ps("pack_y", st);
        TDT_LAUNCH(pair_first_run_kernel, (unsigned)(((int64_t)(P + 1) * 32 + 255) / 256), 256, 0, st, seg_off, P,
                   &b.small->dims_x, b.stwX, b.tprefX, b.small->totals_x, b.gfirst);
        PackParams k = {};
        k.stw = b.stwX;
        k.cvw = b.cvwX;
        k.tile_pref = b.tprefX;
        k.totals = b.small->totals_x;
        k.dims = &b.small->dims_x;
        k.xv = xv;
        k.posB = posB;
        k.max_pos = max_pos;
        k.ykey = b.ykey;
        k.yval = b.yval;
        k.gx = b.gx;
        k.goff = b.goff;
        k.dims_y = &b.small->dims_y;
        k.err = &b.small->err;
        TDT_LAUNCH(pack_y_kernel, (unsigned)wr_tiles(n), WR_THREADS, 0, st, k);
    }
    int rc = run_ypass<false>(pl, b, eps, m, key_bits, P, 0, nullptr, labels_out, nullptr, st);
    if (rc) return rc;
    if (status_accum) {  // no host round trip: the caller reads its status word when it synchronises
        TDT_LAUNCH(accumulate_status_kernel, 1, 1, 0, st, b.small, status_accum);
        return TDT_OK;
    }
    return finish(b, st);
}

}  // namespace tdt

using namespace tdt;

extern "C" {

size_t tdt_cluster_workspace_bytes(int64_t n, int32_t P) {
    if (n < 0) n = 0;
    if (P < 1) P = 1;
    return make_plan(n, P).total;
}

int tdt_cluster_labels(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, int64_t n, int32_t P,
                       int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_out, void *ws, size_t ws_bytes,
                       void *stream) {
    if (P < 1 && n > 0) return fail(TDT_E_ARG, "P = %d pairs for %lld signals", P, (long long)n);
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, ws, ws_bytes, make_plan(n, P < 1 ? 1 : P).total);
    if (rc || n == 0) return rc;
    if (!posA || !posB || !seg_off || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    if (max_pos == 0) max_pos = 0x7fffffff;
    return cluster_impl(posA, posB, seg_off, n, P, eps, min_pts, max_pos, labels_out, ws, ws_bytes,
                        (cudaStream_t)stream, false);
}

int tdt_cluster_labels_async(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, int64_t n, int32_t P,
                             int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_out, void *ws,
                             size_t ws_bytes, int32_t *status_accum, void *stream) {
    if (!status_accum) return fail(TDT_E_ARG, "status_accum is null");
    if (P < 1 && n > 0) return fail(TDT_E_ARG, "P = %d pairs for %lld signals", P, (long long)n);
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, ws, ws_bytes, make_plan(n, P < 1 ? 1 : P).total);
    if (rc || n == 0) return rc;
    if (!posA || !posB || !seg_off || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    if (max_pos == 0) max_pos = 0x7fffffff;
    return cluster_impl(posA, posB, seg_off, n, P, eps, min_pts, max_pos, labels_out, ws, ws_bytes,
                        (cudaStream_t)stream, false, status_accum);
}

int tdt_dbscan_main(const int32_t *x, const int32_t *y, int64_t n, int32_t eps, int32_t min_pts, int32_t max_pos,
                    int32_t *labels_out, void *ws, size_t ws_bytes, void *stream) {
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, ws, ws_bytes, make_plan(n, 1).total);
    if (rc || n == 0) return rc;
    if (!x || !y || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    if (max_pos == 0) max_pos = 0x7fffffff;
    return cluster_impl(x, y, nullptr, n, 1, eps, min_pts, max_pos, labels_out, ws, ws_bytes, (cudaStream_t)stream,
                        true);
}

int tdt_xpass_labels(const int32_t *x, int64_t n, int32_t eps, int32_t min_pts, int32_t *labels_out,
                     int32_t *last_id_out, void *ws, size_t ws_bytes, void *stream) {
    int rc = check_common(n, min_pts, ws, ws_bytes, make_plan(n, 1).total);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (last_id_out) TDT_CUDA(cudaMemsetAsync(last_id_out, 0xff, 4, st));
        return TDT_OK;
    }
    if (!x || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    const ClusterPlan pl = make_plan(n, 1);
    Buffers b = carve(pl, ws, ws_bytes);
    if (!b.ok) return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);
    TDT_CUDA(cudaMemsetAsync(b.headsX, 0, zero_span(pl), st));
    TDT_LAUNCH(set_dims_kernel, 1, 1, 0, st, &b.small->dims_x, n, (int64_t)1, b.small->off2);
    TDT_LAUNCH(pack_plain_kernel, grid_for(n, 256), 256, 0, st, x, n, 0x7fffffff, b.xs, &b.small->err);
    TDT_LAUNCH(heads_from_offsets_kernel, 1, 256, 0, st, b.small->off2, &b.small->dims_x, b.headsX);
    WRParams p = {};
    p.keys = b.xs;
    p.heads = b.headsX;
    p.dims = &b.small->dims_x;
    p.m = min_pts;
    p.eps = eps > 0 ? (u32)eps : 0u;
    p.status = b.statusX;
    p.stw = b.stwX;
    p.cvw = b.cvwX;
    p.tile_pref = b.tprefX;
    p.totals = b.small->totals_x;
    rc = launch_window_runs<MODE_X, true>(p, n, st);
    if (rc) return rc;
    TDT_LAUNCH(expand_xlabels_kernel, grid_for(n, WR_THREADS), WR_THREADS, 0, st, b.stwX, b.cvwX, b.tprefX, n,
               b.small->totals_x, labels_out, last_id_out);
    return finish(b, st);
}

int tdt_ypass_labels(const int32_t *y, int64_t n, int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_io,
                     int32_t *cluster_id_io, void *ws, size_t ws_bytes, void *stream) {
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, ws, ws_bytes, make_plan(n, 1).total);
    if (rc || n == 0) return rc;
    if (!y || !labels_io || !cluster_id_io) return fail(TDT_E_ARG, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int key_bits = pos_bits(max_pos);
    if (max_pos == 0) max_pos = 0x7fffffff;
    int32_t cluster_id = -1;
    TDT_CUDA(cudaMemcpyAsync(&cluster_id, cluster_id_io, 4, cudaMemcpyDeviceToHost, st));
    TDT_CUDA(cudaStreamSynchronize(st));
    if (cluster_id < -1) return fail(TDT_E_ARG, "cluster_id = %d", cluster_id);
    if (cluster_id < 0) return TDT_OK;  // no clusters: nothing to do
    const ClusterPlan pl = make_plan(n, 1);
    if ((int64_t)cluster_id + 4 > pl.gmax)
        return fail(TDT_E_ARG, "cluster_id = %d: the x-pass never yields more than n/2 ids (n = %lld)", cluster_id,
                    (long long)n);
    Buffers b = carve(pl, ws, ws_bytes);
    if (!b.ok) return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);
    const int64_t n_ids = (int64_t)cluster_id + 1;  // real ids 0..cluster_id; the noise id is n_ids

    TDT_CUDA(cudaMemsetAsync(b.headsX, 0, zero_span(pl), st));
    TDT_CUDA(cudaMemsetAsync(b.goff, 0xff, (size_t)(n_ids + 2) * 8, st));
    TDT_LAUNCH(set_dims_kernel, 1, 1, 0, st, &b.small->dims_x, n, (int64_t)1, b.small->off2);
    // 1. group the elements by id (stable): one segment, keys = id
    TDT_LAUNCH(ypass_label_keys_kernel, grid_for(n, 256), 256, 0, st, labels_io, n, cluster_id, b.ykey, &b.small->err);
    rc = segsort_pairs(b.ykey, nullptr, b.xs, b.xv, b.tmpK, b.tmpV, b.small->off2, (const int64_t *)&b.small->dims_x,
                       nullptr, n, 1, bit_width_u32((uint32_t)(cluster_id + 1)), b.sort_temp, pl.sort, &b.small->err,
                       st);
    if (rc) return rc;
    // 2. offsets per id, y gathered; the noise block (last) is left out of everything that follows
    TDT_LAUNCH(ypass_offsets_kernel, grid_for(n, 256), 256, 0, st, b.xs, b.xv, n, y, max_pos, b.goff, b.ykey, b.gx,
               &b.small->err);
    TDT_LAUNCH(ypass_fill_offsets_kernel, (unsigned)((n_ids + 2 + 255) / 256), 256, 0, st, b.goff, n_ids, n);
    TDT_LAUNCH(ypass_dims_kernel, 1, 1, 0, st, b.goff, n_ids, &b.small->dims_y);
    // yval = insertion index in id order (b.xv): copied, because the y sort writes its result into xs / xv
    TDT_CUDA(cudaMemcpyAsync(b.yval, b.xv, (size_t)n * 4, cudaMemcpyDeviceToDevice, st));
    // labels of clustered elements are rewritten by final_labels; members of sub-clusters that dissolve become noise
    TDT_CUDA(cudaMemsetAsync(labels_io, 0xff, (size_t)n * 4, st));
    int32_t *rank_of = b.base1;  // id -> rank among the ids that have members (base1 is unused in plain mode)
    rc = run_ypass<true>(pl, b, eps, min_pts, key_bits, 1, cluster_id, rank_of, labels_io, cluster_id_io, st);
    if (rc) return rc;
    return finish(b, st);
}

/* test hook: the segmented sort on its own (keys/vals/off are device arrays; nseg segments, n elements) */
int tdt_debug_segsort(const uint32_t *keys, const int32_t *vals, const int64_t *off, const int32_t *segid, int64_t nseg,
                      int64_t n, int32_t key_bits, uint32_t *keys_out, int32_t *vals_out, void *ws, size_t ws_bytes,
                      void *stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (n <= 0 || nseg <= 0) return TDT_OK;
    const size_t arr = al256((size_t)n * 4 + 256), tmp = al256(segsort_temp_bytes(n, nseg));
    if (ws_bytes < 2 * arr + 512 + tmp)
        return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, 2 * arr + 512 + tmp);
    Arena ar(ws, ws_bytes);
    u32 *tk = (u32 *)ar.take<char>(arr);
    int32_t *tv = (int32_t *)ar.take<char>(arr);
    Small *small = (Small *)ar.take<char>(256);
    void *temp = ar.take<char>(tmp);
    TDT_CUDA(cudaMemsetAsync(small, 0, 256, st));
    TDT_LAUNCH(set_dims_kernel, 1, 1, 0, st, &small->dims_x, n, nseg, (int64_t *)nullptr);
    int rc = segsort_pairs(keys, vals, keys_out, vals_out, tk, tv, off, (const int64_t *)&small->dims_x, segid, n, nseg,
                           key_bits, temp, tmp, &small->err, st);
    if (rc) return rc;
    Small h;
    TDT_CUDA(cudaMemcpyAsync(&h, small, sizeof(Small), cudaMemcpyDeviceToHost, st));
    TDT_CUDA(cudaStreamSynchronize(st));
    if (h.err == SS_ERR_INTERNAL) return fail(TDT_E_CUDA, "segmented sort: a work list overflowed (internal error)");
    return h.err ? fail(TDT_E_RANGE, "a key has bits above key_bits") : TDT_OK;
}

}  // extern "C"

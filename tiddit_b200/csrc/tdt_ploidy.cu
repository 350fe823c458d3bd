// tdt_ploidy.cu -- masked coverage medians on B200 (sm_100a).
//
// Replaces the bin loops of tiddit/tiddit_coverage_analysis.pyx:14-29 (determine_ploidy): per contig the median of
// the coverage bins with coverage > 0 and GC != -1 (:17-22, numpy.median), and the same over all contigs (:26-27).
// numpy.median of k values is the middle value (k odd) or the mean of the two middle values (k even); both are
// found EXACTLY by a most-significant-digit radix select on the 64-bit patterns of the (positive) doubles:
//
//   med_pass    6 passes of 11 / 11 / 11 / 11 / 11 / 9 bits: per segment (every contig + "all") a histogram of the current
//               digit over the bins whose higher digits match the segment's prefix.  A CTA owns a CONTIGUOUS range of
//               tiles and keeps two 2048-bin histograms in shared memory across them -- one for "all", one for the
//               contig it is in (flushed when the contig changes) -- with per-thread run-length aggregation
//               (neighbouring bins mostly share the high digits); bins of other contigs inside a boundary tile go to
//               global atomics
//   med_select  one warp per segment: the digit that holds the wanted rank, prefix and rank narrowed
//               (last pass: also the smallest key above the prefix bucket, for the value after the lower median)
//   med_finish  (v1 + v2) / 2 like numpy, NaN for an empty selection
//   med_compact (r02) after the second pass the prefix buckets (exponent + 10 mantissa bits: a fraction of a percent of
//               the bins) are copied out once and passes 3-6 run on the copies (med_cpass) -- see MedCompact below
// Algorithmic bytes: 9 B/bin (float64 coverage + int8 GC); the implementation streams them 3 times (r01_v5: 6 times,
// 1.10 ms for 61.8 M bins; r01_v4: 9 times with 8-bit digits and a separate next-value pass, 1.92 ms).
#include "tdt_common.cuh"

namespace tdt {

constexpr int MD_THREADS = 256;
constexpr int MD_ITEMS = 8;
constexpr int MD_TILE = MD_THREADS * MD_ITEMS;
#ifndef TDT_MD_BITS
#define TDT_MD_BITS 11
#endif
constexpr int MD_BITS = TDT_MD_BITS;        // widest digit
constexpr int MD_NB = 1 << MD_BITS;         // histogram bins per segment
constexpr int MD_PASSES = (64 + MD_BITS - 1) / MD_BITS;
constexpr size_t MD_SMEM = (size_t)2 * MD_NB * sizeof(u32);
static_assert(MD_PASSES >= 2 && (64 - (MD_PASSES - 1) * MD_BITS) >= 5, "the last digit must have at least 5 bits (one digit per lane)");

struct MedState {
    u64 prefix;      // digits selected so far (in place, lower bits zero)
    u64 next;        // smallest key above the lower median
    int64_t count;   // selected bins
    int64_t rank;    // wanted rank among the bins matching the prefix
    int64_t less;    // bins below the prefix
    int64_t equal;   // (after the last pass) bins equal to the lower median
    int64_t bucket;  // bins inside the prefix bucket selected by the latest pass
};

// Compaction after the second digit pass (r02).  22 bits of a positive double are its exponent and the top 10 bits of
// its mantissa: the bins that share them with a segment's median are a fraction of a percent of the genome, so they are
// copied out ONCE (med_compact_kernel, the third and last streaming pass) and the remaining digit passes run on the
// copies -- three streaming passes over 9 B/bin instead of six.  If the buckets of all segments together exceed the
// reserved capacity (a genome whose bins are nearly all equal) the original passes run instead; `ok` decides, on the
// device, which of the two sets of kernels does anything.
constexpr int64_t MD_COMPACT_CAP = 4 << 20;   // keys (32 MB)
constexpr int MD_SPLIT_PASSES = 2;            // streaming digit passes before the compaction
constexpr int MD_MAX_SEG = 1023;              // segments (contigs + "all") the compact path supports

struct MedCompact {
    int32_t ok;                       // 1: the compact lists hold every bucket
    int32_t pad;
    int64_t total;                    // keys in all lists
    int64_t off[MD_MAX_SEG + 2];      // list offsets (segment s owns [off[s], off[s+1]))
    unsigned long long cursor[MD_MAX_SEG + 1];
};

struct MedParams {
    const double *bins;
    const int8_t *gc;
    const int64_t *bin_off;   // [C+1]
    int64_t n;
    int C;
    MedState *state;          // [C+1], segment C = all contigs
    u32 *hist;                // [C+1][MD_NB]
    MedCompact *cpt;          // compaction header
    u64 *ckeys;               // [MD_COMPACT_CAP] compacted keys, list by list
    double *medians;
    int64_t *counts;
};

__device__ __forceinline__ int md_find(const int64_t *__restrict__ off, int C, int64_t i) {
    int lo = 0, hi = C;
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// digit passes: digit = bits [shift, shift + width) of the key.  NEXT (the last pass): also the smallest key whose
// higher digits lie ABOVE the segment's prefix -- together with the next non-empty digit of the last histogram
// (med_select) that is the value following the lower median, which even counts need; no extra pass over the bins.
template <bool FIRST, bool NEXT>
__global__ void __launch_bounds__(MD_THREADS) med_pass_kernel(MedParams p, int shift, int width, int after_split) {
    if (after_split && p.cpt->ok) return;   // the compact lists carry the remaining passes
    extern __shared__ __align__(16) u32 md_smem[];
    u32 *hAll = md_smem, *hCtg = md_smem + MD_NB;
    __shared__ u64 sNextAll, sNextCtg;
    const int64_t tiles = (p.n + MD_TILE - 1) / MD_TILE;
    const int64_t per = (tiles + gridDim.x - 1) / gridDim.x;          // a contiguous run of tiles per CTA
    const int64_t tile_lo = (int64_t)blockIdx.x * per;
    const int64_t tile_hi = tile_lo + per < tiles ? tile_lo + per : tiles;
    if (tile_lo >= tile_hi) return;
    const MedState stAll = p.state[p.C];
    const u32 dmask = (1u << width) - 1u;
    const int nb = 1 << width;
    for (int d = threadIdx.x; d < nb; d += MD_THREADS) {
        hAll[d] = 0;
        hCtg[d] = 0;
    }
    if (threadIdx.x == 0) {
        sNextAll = ~0ull;
        sNextCtg = ~0ull;
    }
    // the contig whose histogram lives in shared memory, its end and its prefix: kept in registers so that the common
    // tile (whole inside one contig) issues its 72 bytes of loads per thread without a dependent look-up in front
    int cur = md_find(p.bin_off, p.C, tile_lo * MD_TILE);
    int64_t cur_end = p.bin_off[cur + 1];
    u64 cur_prefix = p.state[cur].prefix;
    __syncthreads();
    for (int64_t tile = tile_lo; tile < tile_hi; tile++) {
        const int64_t t0 = tile * MD_TILE;
        const int64_t i0 = t0 + (int64_t)threadIdx.x * MD_ITEMS;
        // this thread's 8 bins: 64 B of coverage + 8 B of GC, vector loads when whole and aligned
        double vv[MD_ITEMS];
        int8_t gg[MD_ITEMS];
        if (i0 + MD_ITEMS <= p.n && ((((uintptr_t)p.bins) & 15) == 0) && ((((uintptr_t)p.gc) & 7) == 0)) {
            const double2 *b2 = (const double2 *)(p.bins + i0);
#pragma unroll
            for (int k = 0; k < MD_ITEMS / 2; k++) {
                const double2 t = b2[k];
                vv[2 * k] = t.x;
                vv[2 * k + 1] = t.y;
            }
            const uint2 g2 = *(const uint2 *)(p.gc + i0);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                gg[k] = (int8_t)(g2.x >> (8 * k));
                gg[4 + k] = (int8_t)(g2.y >> (8 * k));
            }
        } else {
#pragma unroll
            for (int k = 0; k < MD_ITEMS; k++) {
                vv[k] = i0 + k < p.n ? p.bins[i0 + k] : 0.0;
                gg[k] = i0 + k < p.n ? p.gc[i0 + k] : (int8_t)-1;
            }
        }
        if (t0 >= cur_end) {   // the run of tiles moved on to another contig: hand the cached histogram over
            __syncthreads();
            if (NEXT && threadIdx.x == 0) {
                if (sNextCtg != ~0ull) atomicMin(&p.state[cur].next, sNextCtg);
                sNextCtg = ~0ull;
            }
            for (int d = threadIdx.x; d < nb; d += MD_THREADS) {
                const u32 b = hCtg[d];
                if (b) atomicAdd(p.hist + (int64_t)cur * MD_NB + d, b);
                hCtg[d] = 0;
            }
            while (t0 >= p.bin_off[cur + 1]) cur++;
            cur_end = p.bin_off[cur + 1];
            cur_prefix = p.state[cur].prefix;
            __syncthreads();
        }
        if (i0 < p.n) {
            int c = cur;
            int64_t cend = cur_end;
            u64 prefix = cur_prefix;
            if (i0 >= cend) {
                while (i0 >= p.bin_off[c + 1]) c++;
                cend = p.bin_off[c + 1];
                prefix = p.state[c].prefix;
            }
            // run-length aggregation: (target histogram, digit) of the previous bin
            int runA_d = -1, runC_d = -1, runC_c = c;
            u32 runA_n = 0, runC_n = 0;
            u64 nextA = ~0ull, nextC = ~0ull;
            auto flushC = [&]() {
                if (runC_n) {
                    if (runC_c == cur) atomicAdd(&hCtg[runC_d], runC_n);
                    else atomicAdd(p.hist + (int64_t)runC_c * MD_NB + runC_d, runC_n);
                }
                runC_n = 0;
            };
            auto flushNextC = [&]() {
                if (NEXT && nextC != ~0ull) {
                    if (c == cur) atomicMin(&sNextCtg, nextC);
                    else atomicMin(&p.state[c].next, nextC);
                }
                nextC = ~0ull;
            };
#pragma unroll
            for (int k = 0; k < MD_ITEMS; k++) {
                const int64_t i = i0 + k;
                if (i >= p.n) continue;
                if (i >= cend) {
                    flushC();
                    flushNextC();
                    while (i >= p.bin_off[c + 1]) c++;
                    cend = p.bin_off[c + 1];
                    prefix = p.state[c].prefix;
                    runC_c = c;
                    runC_d = -1;
                }
                const double v = vv[k];
                if (!(v > 0.0) || gg[k] == -1) continue;              // tiddit_coverage_analysis.pyx:18
                const u64 key = (u64)__double_as_longlong(v);
                const int d = (int)((u32)(key >> shift) & dmask);
                const u64 hi = FIRST ? 0ull : (key >> (shift + width));
                if (NEXT) {
                    if (hi > (stAll.prefix >> (shift + width)) && key < nextA) nextA = key;
                    if (hi > (prefix >> (shift + width)) && key < nextC) nextC = key;
                }
                if (FIRST || hi == (stAll.prefix >> (shift + width))) {
                    if (d == runA_d) runA_n++;
                    else {
                        if (runA_n) atomicAdd(&hAll[runA_d], runA_n);
                        runA_d = d;
                        runA_n = 1;
                    }
                }
                if (FIRST || hi == (prefix >> (shift + width))) {
                    if (d == runC_d) runC_n++;
                    else {
                        flushC();
                        runC_d = d;
                        runC_n = 1;
                    }
                }
            }
            if (runA_n) atomicAdd(&hAll[runA_d], runA_n);
            flushC();
            flushNextC();
            if (NEXT && nextA != ~0ull) atomicMin(&sNextAll, nextA);
        }
    }
    __syncthreads();
    if (NEXT && threadIdx.x == 0) {
        if (sNextAll != ~0ull) atomicMin(&p.state[p.C].next, sNextAll);
        if (sNextCtg != ~0ull) atomicMin(&p.state[cur].next, sNextCtg);
    }
    for (int d = threadIdx.x; d < nb; d += MD_THREADS) {
        const u32 a = hAll[d], b = hCtg[d];
        if (a) atomicAdd(p.hist + (int64_t)p.C * MD_NB + d, a);
        if (b) atomicAdd(p.hist + (int64_t)cur * MD_NB + d, b);
    }
}

// ---- compaction (after MD_SPLIT_PASSES digit passes) ----------------------------------------------------------------------
// list offsets from the bucket sizes; one CTA
__global__ void med_plan_kernel(MedParams p) {
    if (threadIdx.x != 0) return;
    MedCompact *c = p.cpt;
    int64_t run = 0;
    const bool fits = p.C <= MD_MAX_SEG - 1;
    for (int s = 0; s <= p.C && fits; s++) {
        c->off[s] = run;
        c->cursor[s] = 0ull;
        run += p.state[s].count > 0 ? p.state[s].bucket : 0;
    }
    if (fits) c->off[p.C + 1] = run;
    c->total = run;
    c->ok = fits && run <= MD_COMPACT_CAP ? 1 : 0;
}

// atomicMin that first looks: the minimum only ever decreases, so a candidate that is not below the value just read can
// be dropped without the atomic (a stale read only costs a redundant atomic) -- after a few updates nearly every call
// takes the cheap path, instead of millions of atomics serialising on one address per contig
__device__ __forceinline__ void med_min_u64(u64 *addr, u64 v) {
    if (v < ld_volatile_u64(addr)) atomicMin(addr, v);
}

// the third streaming pass: every qualifying bin whose top `top_bits` bits equal its contig's prefix is appended to the
// contig's list, likewise for "all"; the smallest key ABOVE either bucket is kept for the upper median (what the last
// pass kernel's NEXT logic finds among the bins the lists no longer contain)
__global__ void __launch_bounds__(MD_THREADS) med_compact_kernel(MedParams p, int low_bits) {
    if (!p.cpt->ok) return;
    __shared__ u64 sNextAll;
    if (threadIdx.x == 0) sNextAll = ~0ull;
    __syncthreads();
    const MedState stAll = p.state[p.C];
    const u64 topAll = stAll.prefix >> low_bits;
    u64 nextA = ~0ull, nextC = ~0ull;
    // a CTA owns a CONTIGUOUS run of tiles (like the digit passes): a thread stays inside one contig for almost all of
    // its bins, so the running minimum above the bucket lives in a register and reaches memory once per contig
    const int64_t tiles = (p.n + MD_TILE - 1) / MD_TILE;
    const int64_t per = (tiles + gridDim.x - 1) / gridDim.x;
    const int64_t tile_lo = (int64_t)blockIdx.x * per;
    const int64_t tile_hi = tile_lo + per < tiles ? tile_lo + per : tiles;
    int c = tile_lo < tile_hi ? md_find(p.bin_off, p.C, tile_lo * MD_TILE + (int64_t)threadIdx.x * MD_ITEMS < p.n
                                                             ? tile_lo * MD_TILE + (int64_t)threadIdx.x * MD_ITEMS : p.n - 1)
                              : 0;
    int64_t cend = p.bin_off[c + 1];
    u64 topC = p.state[c].prefix >> low_bits;
    for (int64_t tile = tile_lo; tile < tile_hi; tile++) {
        const int64_t i0 = tile * MD_TILE + (int64_t)threadIdx.x * MD_ITEMS;
        if (i0 >= p.n) break;
        double vv[MD_ITEMS];
        int8_t gg[MD_ITEMS];
        if (i0 + MD_ITEMS <= p.n && ((((uintptr_t)p.bins) & 15) == 0) && ((((uintptr_t)p.gc) & 7) == 0)) {
            const double2 *b2 = (const double2 *)(p.bins + i0);
#pragma unroll
            for (int k = 0; k < MD_ITEMS / 2; k++) {
                const double2 t = b2[k];
                vv[2 * k] = t.x;
                vv[2 * k + 1] = t.y;
            }
            const uint2 g2 = *(const uint2 *)(p.gc + i0);
#pragma unroll
            for (int k = 0; k < 4; k++) {
                gg[k] = (int8_t)(g2.x >> (8 * k));
                gg[4 + k] = (int8_t)(g2.y >> (8 * k));
            }
        } else {
#pragma unroll
            for (int k = 0; k < MD_ITEMS; k++) {
                vv[k] = i0 + k < p.n ? p.bins[i0 + k] : 0.0;
                gg[k] = i0 + k < p.n ? p.gc[i0 + k] : (int8_t)-1;
            }
        }
#pragma unroll
        for (int k = 0; k < MD_ITEMS; k++) {
            const int64_t i = i0 + k;
            if (i >= p.n) continue;
            if (i >= cend) {
                if (nextC != ~0ull) med_min_u64(&p.state[c].next, nextC);
                nextC = ~0ull;
                while (i >= p.bin_off[c + 1]) c++;
                cend = p.bin_off[c + 1];
                topC = p.state[c].prefix >> low_bits;
            }
            const double v = vv[k];
            if (!(v > 0.0) || gg[k] == -1) continue;
            const u64 key = (u64)__double_as_longlong(v);
            const u64 top = key >> low_bits;
            if (top == topC) {
                const unsigned long long slot = atomicAdd(&p.cpt->cursor[c], 1ull);
                if ((int64_t)slot < p.cpt->off[c + 1] - p.cpt->off[c]) p.ckeys[p.cpt->off[c] + (int64_t)slot] = key;
            } else if (top > topC && key < nextC) nextC = key;
            if (top == topAll) {
                const unsigned long long slot = atomicAdd(&p.cpt->cursor[p.C], 1ull);
                if ((int64_t)slot < p.cpt->off[p.C + 1] - p.cpt->off[p.C]) p.ckeys[p.cpt->off[p.C] + (int64_t)slot] = key;
            } else if (top > topAll && key < nextA) nextA = key;
        }
    }
    if (nextC != ~0ull) med_min_u64(&p.state[c].next, nextC);
    if (nextA != ~0ull) atomicMin(&sNextAll, nextA);
    __syncthreads();
    if (threadIdx.x == 0 && sNextAll != ~0ull) med_min_u64(&p.state[p.C].next, sNextAll);
}

// One CTA (MS_THREADS threads) picks the digit of histogram `h` (global or shared) that holds the wanted rank, narrows
// the segment's state and clears the histogram.  Every thread passes the same `st` and gets the updated one back.
// (r01: one WARP per segment, every lane walking 64 strided counters three times -- 25 us per call, six calls; now the
// 2048 counters are read once with coalesced 32-byte loads and a block scan finds the owner of the rank.)
constexpr int MS_THREADS = 256;
struct SelectSmem {
    u32 warp[MS_THREADS / 32];
    int gd, best;
    MedState st;
};

__device__ __forceinline__ MedState md_select_cta(SelectSmem &sm, u32 *h, int shift, int width, bool first, bool last,
                                                  MedState st) {
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const int nb = 1 << width;
    const int per = (nb + MS_THREADS - 1) / MS_THREADS;     // consecutive digits per thread (8 for 11 bits, 2 for 9)
    u32 v[MD_NB / MS_THREADS];
    u32 mine = 0;
#pragma unroll
    for (int k = 0; k < MD_NB / MS_THREADS; k++) {
        const int d = t * per + k;
        v[k] = (k < per && d < nb) ? h[d] : 0u;
        mine += v[k];
    }
    u32 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 x = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += x;
    }
    __syncthreads();                 // (sm is reused from call to call)
    if (lane == 31) sm.warp[warp] = inc;
    if (t == 0) {
        sm.gd = -1;
        sm.best = 1 << 30;
    }
    __syncthreads();
    u32 wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < MS_THREADS / 32; w++) {
        if (w < warp) wbase += sm.warp[w];
        total += sm.warp[w];
    }
    inc += wbase;
    if (first) {
        st.prefix = 0;
        st.next = ~0ull;
        st.count = total;                                  // histogram totals can exceed 2^32 only beyond 4G bins
        st.rank = total ? (int64_t)(total - 1) / 2 : 0;    // lower median
        st.less = 0;
        st.equal = 0;
        st.bucket = 0;
    }
    if (t == 0) sm.st = st;                                // what everybody gets back when nothing is selected
    const u32 before = inc - mine;   // bins in lower threads
    const bool owner = total != 0u && mine != 0u && (int64_t)before <= st.rank && st.rank < (int64_t)inc;
    __syncthreads();
    if (owner) {
        u32 cum = before;
        int d = 0;
#pragma unroll
        for (int k = 0; k < MD_NB / MS_THREADS; k++) {
            if (k < per) {
                if (st.rank >= (int64_t)cum + v[k]) cum += v[k];
                else {
                    d = k;
                    st.bucket = v[k];
                    if (last) st.equal = v[k];
                    break;
                }
            }
        }
        const int gd = t * per + d;
        st.prefix |= (u64)gd << shift;
        st.less += cum;
        st.rank -= cum;
        sm.gd = gd;
    }
    __syncthreads();
    if (last && total != 0u) {
        // the value after the lower median: the next non-empty digit of this histogram, else the smallest key above
        // the prefix bucket that the pass kernels left in st.next
        const int gd = sm.gd;
        int best = 1 << 30;
#pragma unroll
        for (int k = MD_NB / MS_THREADS - 1; k >= 0; k--) {
            const int d = t * per + k;
            if (k < per && d > gd && v[k]) best = d;
        }
        if (best < (1 << 30)) atomicMin(&sm.best, best);
    }
    __syncthreads();
    if (owner) {
        if (last && sm.best < (1 << 30)) {
            const u64 cand = (st.prefix & ~(((u64)(1u << width) - 1ull) << shift)) | ((u64)sm.best << shift);
            if (cand < st.next) st.next = cand;
        }
        sm.st = st;
    }
#pragma unroll
    for (int k = 0; k < MD_NB / MS_THREADS; k++) {
        const int d = t * per + k;
        if (k < per && d < nb) h[d] = 0u;
    }
    __syncthreads();
    return sm.st;
}

__global__ void __launch_bounds__(MS_THREADS) med_select_kernel(MedParams p, int shift, int width, int first, int last,
                                                                 int after_split) {
    __shared__ SelectSmem sm;
    const int seg = blockIdx.x;
    if (seg > p.C) return;
    if (after_split && p.cpt->ok) return;   // med_cfinish_kernel did the remaining passes of every segment
    const MedState st = md_select_cta(sm, p.hist + (int64_t)seg * MD_NB, shift, width, first != 0, last != 0, p.state[seg]);
    if (threadIdx.x == 0) p.state[seg] = st;
}

// the digit passes after the compaction, ALL of them, for one segment per CTA: its list is a fraction of a percent of
// the bins, so histogram (shared memory), select and the next pass follow each other inside one launch
__global__ void __launch_bounds__(MS_THREADS) med_cfinish_kernel(MedParams p, int hi0) {
    __shared__ SelectSmem sm;
    __shared__ u32 h[MD_NB];
    __shared__ u64 s_next;
    if (!p.cpt->ok) return;
    const int seg = blockIdx.x;
    if (seg > p.C) return;
    MedState st = p.state[seg];
    if (st.count == 0) return;
    const int64_t lo = p.cpt->off[seg], cnt = p.cpt->off[seg + 1] - lo;
    for (int d = threadIdx.x; d < MD_NB; d += MS_THREADS) h[d] = 0u;
    if (threadIdx.x == 0) s_next = ~0ull;
    u64 next = ~0ull;
    int hi = hi0;
    for (int pass = MD_SPLIT_PASSES; pass < MD_PASSES; pass++) {
        const bool last = pass == MD_PASSES - 1;
        const int width = hi - MD_BITS >= 0 ? (last ? hi : MD_BITS) : hi;
        const int shift = hi - width;
        const u32 dmask = (1u << width) - 1u;
        const u64 pup = st.prefix >> (shift + width);
        __syncthreads();
        // one CTA walks its list alone: eight independent loads per thread and round, or the walk is one L2 round
        // trip per key (82 us for the 40 k keys of the genome-wide list before the unrolling)
        for (int64_t e0 = threadIdx.x; e0 < cnt; e0 += MS_THREADS * 8) {
            u64 kk[8];
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int64_t e = e0 + (int64_t)j * MS_THREADS;
                kk[j] = e < cnt ? p.ckeys[lo + e] : 0ull;     // 0 is no key (coverage > 0): matches no prefix above
            }
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const u64 key = kk[j];
                if (key == 0ull) continue;
                const u64 up = key >> (shift + width);
                if (up == pup) atomicAdd(&h[(u32)(key >> shift) & dmask], 1u);
                else if (last && up > pup && key < next) next = key;
            }
        }
        __syncthreads();
        st = md_select_cta(sm, h, shift, width, false, last, st);
        hi = shift;
    }
    if (next != ~0ull) atomicMin(&s_next, next);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_next < st.next) st.next = s_next;
        p.state[seg] = st;
    }
}

__global__ void med_finish_kernel(MedParams p) {
    const int seg = blockIdx.x * blockDim.x + threadIdx.x;
    if (seg > p.C) return;
    const MedState st = p.state[seg];
    double m;
    if (st.count == 0) {
        m = __longlong_as_double(0x7ff8000000000000LL);   // numpy.median([]) -> nan
    } else {
        const double v1 = __longlong_as_double((long long)st.prefix);
        const int64_t r1 = (st.count - 1) / 2, r2 = st.count / 2;
        if (r2 == r1 || r2 < st.less + st.equal) m = v1;       // odd count, or the upper median equals the lower
        else m = (v1 + __longlong_as_double((long long)st.next)) / 2.0;   // numpy: mean of the two middle values
    }
    p.medians[seg] = m;
    p.counts[seg] = st.count;
}

}  // namespace tdt

using namespace tdt;

extern "C" {

size_t tdt_coverage_medians_workspace_bytes(int32_t C) {
    if (C < 0) return 0;
    return ((size_t)(C + 1) * sizeof(MedState) + 255) / 256 * 256 + ((size_t)(C + 1) * MD_NB * 4 + 255) / 256 * 256 +
           (sizeof(MedCompact) + 255) / 256 * 256 + (size_t)MD_COMPACT_CAP * 8 + 256;
}

int tdt_coverage_medians(const double *bins, const int8_t *gc, const int64_t *bin_off, int32_t C, int64_t n_bins,
                         double *medians_out, int64_t *counts_out, void *ws, size_t ws_bytes, void *stream) {
    if (C < 0 || n_bins < 0) return fail(TDT_E_ARG, "negative size");
    if (!medians_out || !counts_out) return fail(TDT_E_ARG, "null output pointer");
    if (n_bins > 0 && (!bins || !gc || !bin_off || C < 1)) return fail(TDT_E_ARG, "null input pointer or no contig");
    if (ws == nullptr || ws_bytes < tdt_coverage_medians_workspace_bytes(C))
        return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes,
                    tdt_coverage_medians_workspace_bytes(C));
    cudaStream_t st = (cudaStream_t)stream;
    MedParams p = {};
    p.bins = bins;
    p.gc = gc;
    p.bin_off = bin_off;
    p.n = n_bins;
    p.C = C;
    p.state = (MedState *)ws;
    const size_t state_bytes = ((size_t)(C + 1) * sizeof(MedState) + 255) / 256 * 256;
    const size_t hist_bytes = ((size_t)(C + 1) * MD_NB * 4 + 255) / 256 * 256;
    p.hist = (u32 *)((char *)ws + state_bytes);
    p.cpt = (MedCompact *)((char *)ws + state_bytes + hist_bytes);
    p.ckeys = (u64 *)((char *)p.cpt + (sizeof(MedCompact) + 255) / 256 * 256);
    p.medians = medians_out;
    p.counts = counts_out;
    TDT_CUDA(cudaMemsetAsync(ws, 0, state_bytes + hist_bytes + sizeof(MedCompact), st));
    int64_t tiles = (n_bins + MD_TILE - 1) / MD_TILE;
    static thread_local int per_sm = 0, sms = 0;
    if (!per_sm) {
        int dev = 0;
        TDT_CUDA(cudaGetDevice(&dev));
        TDT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        TDT_CUDA(cudaFuncSetAttribute(med_pass_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MD_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(med_pass_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MD_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(med_pass_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)MD_SMEM));
        TDT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, med_pass_kernel<false, false>, MD_THREADS, MD_SMEM));
        if (per_sm < 1) per_sm = 1;
    }
    const int64_t cap = (int64_t)sms * per_sm;
    const unsigned grid = (unsigned)(tiles < cap ? (tiles > 0 ? tiles : 1) : cap);
    const unsigned sel_grid = (unsigned)(C + 1);   // one CTA per segment
    ProfScope ps("coverage_medians", st);
    int hi = 64;
    for (int pass = 0; pass < MD_PASSES; pass++) {
        const int width = hi - MD_BITS >= 0 ? (pass == MD_PASSES - 1 ? hi : MD_BITS) : hi;
        const int shift = hi - width;
        const int after = pass >= MD_SPLIT_PASSES ? 1 : 0;
        if (n_bins > 0) {
            if (pass == MD_SPLIT_PASSES) {   // the buckets are narrow now: copy them out once and finish on the copies
                TDT_LAUNCH(med_plan_kernel, 1, 32, 0, st, p);
                TDT_LAUNCH(med_compact_kernel, grid, MD_THREADS, 0, st, p, hi);
                TDT_LAUNCH(med_cfinish_kernel, sel_grid, MS_THREADS, 0, st, p, hi);
            }
            if (pass == 0) TDT_LAUNCH((med_pass_kernel<true, false>), grid, MD_THREADS, MD_SMEM, st, p, shift, width, after);
            else if (pass < MD_PASSES - 1) TDT_LAUNCH((med_pass_kernel<false, false>), grid, MD_THREADS, MD_SMEM, st, p, shift, width, after);
            else TDT_LAUNCH((med_pass_kernel<false, true>), grid, MD_THREADS, MD_SMEM, st, p, shift, width, after);
        }
        TDT_LAUNCH(med_select_kernel, sel_grid, MS_THREADS, 0, st, p, shift, width, pass == 0, pass == MD_PASSES - 1, after);
        hi = shift;
    }
    TDT_LAUNCH(med_finish_kernel, (unsigned)((C + 1 + 255) / 256), 256, 0, st, p);
    return TDT_OK;
}

}  // extern "C"

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
// Algorithmic bytes: 9 B/bin (float64 coverage + int8 GC); the implementation streams them 6 times (r01_v4: 9 times
// with 8-bit digits and a separate next-value pass, 1.92 ms for 61.8 M bins).
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
};

struct MedParams {
    const double *bins;
    const int8_t *gc;
    const int64_t *bin_off;   // [C+1]
    int64_t n;
    int C;
    MedState *state;          // [C+1], segment C = all contigs
    u32 *hist;                // [C+1][MD_NB]
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
__global__ void __launch_bounds__(MD_THREADS) med_pass_kernel(MedParams p, int shift, int width) {
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

// one warp per segment: pick the digit holding the wanted rank, clear the histogram for the next pass
__global__ void __launch_bounds__(256) med_select_kernel(MedParams p, int shift, int width, int first, int last) {
    const int seg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (seg > p.C) return;
    u32 *h = p.hist + (int64_t)seg * MD_NB;
    const int per = (1 << width) / 32;      // lane owns digits [per*lane, per*lane + per); width >= 5
    u32 mine = 0;
    for (int k = 0; k < per; k++) mine += h[lane * per + k];
    u32 inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    const u32 total = __shfl_sync(0xffffffffu, inc, 31);
    MedState st = p.state[seg];
    if (first) {
        st.prefix = 0;
        st.next = ~0ull;
        st.count = total;                                  // histogram totals can exceed 2^32 only beyond 4G bins
        st.rank = total ? (int64_t)(total - 1) / 2 : 0;    // lower median
        st.less = 0;
        st.equal = 0;
    }
    if (total == 0) {
        if (lane == 0 && first) p.state[seg] = st;
        return;
    }
    const u32 before = inc - mine;   // bins in lower lanes
    const bool owner = (int64_t)before <= st.rank && st.rank < (int64_t)inc;
    int gd = 0;                      // the selected digit (owner lane)
    if (owner) {
        u32 cum = before;
        int d = 0;
        for (int k = 0; k < per; k++) {
            const u32 v = h[lane * per + k];
            if (st.rank >= (int64_t)cum + v) cum += v;
            else {
                d = k;
                if (last) st.equal = v;
                break;
            }
        }
        gd = lane * per + d;
        st.prefix |= (u64)gd << shift;
        st.less += cum;
        st.rank -= cum;
    }
    if (last) {
        // the value after the lower median: the next non-empty digit of this histogram, else the smallest key above
        // the prefix bucket that the pass kernel left in st.next
        const int src = __ffs(__ballot_sync(0xffffffffu, owner)) - 1;
        gd = __shfl_sync(0xffffffffu, gd, src);
        int best = 1 << 30;
        for (int k = per - 1; k >= 0; k--) {
            const int d = lane * per + k;
            if (d > gd && h[d]) best = d;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, o));
        if (owner && best < (1 << 30)) {
            const u64 cand = (st.prefix & ~(((u64)(1u << width) - 1ull) << shift)) | ((u64)best << shift);
            if (cand < st.next) st.next = cand;
        }
    }
    if (owner) p.state[seg] = st;
    __syncwarp();
    for (int k = 0; k < per; k++) h[lane * per + k] = 0;
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
    return ((size_t)(C + 1) * sizeof(MedState) + 255) / 256 * 256 + (size_t)(C + 1) * MD_NB * 4 + 256;
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
    p.hist = (u32 *)((char *)ws + ((size_t)(C + 1) * sizeof(MedState) + 255) / 256 * 256);
    p.medians = medians_out;
    p.counts = counts_out;
    TDT_CUDA(cudaMemsetAsync(ws, 0, tdt_coverage_medians_workspace_bytes(C) - 256, st));
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
    const unsigned sel_grid = (unsigned)(((int64_t)(C + 1) * 32 + 255) / 256);
    ProfScope ps("coverage_medians", st);
    int hi = 64;
    for (int pass = 0; pass < MD_PASSES; pass++) {
        const int width = hi - MD_BITS >= 0 ? (pass == MD_PASSES - 1 ? hi : MD_BITS) : hi;
        const int shift = hi - width;
        if (n_bins > 0) {
            if (pass == 0) TDT_LAUNCH((med_pass_kernel<true, false>), grid, MD_THREADS, MD_SMEM, st, p, shift, width);
            else if (pass < MD_PASSES - 1) TDT_LAUNCH((med_pass_kernel<false, false>), grid, MD_THREADS, MD_SMEM, st, p, shift, width);
            else TDT_LAUNCH((med_pass_kernel<false, true>), grid, MD_THREADS, MD_SMEM, st, p, shift, width);
        }
        TDT_LAUNCH(med_select_kernel, sel_grid, 256, 0, st, p, shift, width, pass == 0, pass == MD_PASSES - 1);
        hi = shift;
    }
    TDT_LAUNCH(med_finish_kernel, (unsigned)((C + 1 + 255) / 256), 256, 0, st, p);
    return TDT_OK;
}

}  // extern "C"

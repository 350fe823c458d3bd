// tdt_ploidy.cu -- masked coverage medians on B200 (sm_100a).
//
// Replaces the bin loops of tiddit/tiddit_coverage_analysis.pyx:14-29 (determine_ploidy): per contig the median of
// the coverage bins with coverage > 0 and GC != -1 (:17-22, numpy.median), and the same over all contigs (:26-27).
// numpy.median of k values is the middle value (k odd) or the mean of the two middle values (k even); both are
// found EXACTLY by a most-significant-digit radix select on the 64-bit patterns of the (positive) doubles:
//
//   med_hist    8 passes, 8 bits each: per segment (every contig + "all") a 256-bin histogram of the current digit
//               over the bins whose higher digits match the segment's prefix; shared-memory histograms for the
//               tile's contig and for "all", per-thread run-length aggregation (neighbouring bins mostly share the
//               high digits), one global atomic per non-empty bin per CTA
//   med_select  one warp per segment: the digit that holds the wanted rank, prefix and rank narrowed
//   med_next    the smallest value above the lower median (for even counts)
//   med_finish  (v1 + v2) / 2 like numpy, NaN for an empty selection
// Algorithmic bytes: 9 B/bin (float64 coverage + int8 GC); the implementation streams them 9 times.
#include "tdt_common.cuh"

namespace tdt {

constexpr int MD_THREADS = 256;
constexpr int MD_ITEMS = 8;
constexpr int MD_TILE = MD_THREADS * MD_ITEMS;

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
    u32 *hist;                // [C+1][256]
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

// PASS 0..7: digit = bits [56 - 8*PASS, 64 - 8*PASS) ; PASS 8: the "next value" pass
template <bool FIRST, bool NEXT>
__global__ void __launch_bounds__(MD_THREADS) med_pass_kernel(MedParams p, int shift) {
    __shared__ u32 hAll[256], hCtg[256];
    __shared__ u64 sNextAll, sNextCtg;
    const int64_t tiles = (p.n + MD_TILE - 1) / MD_TILE;
    const MedState stAll = p.state[p.C];
    for (int64_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        const int64_t t0 = tile * MD_TILE;
        const int c0 = md_find(p.bin_off, p.C, t0);   // the tile's first contig owns the shared histogram
        hAll[threadIdx.x] = 0;
        hCtg[threadIdx.x] = 0;
        if (threadIdx.x == 0) {
            sNextAll = ~0ull;
            sNextCtg = ~0ull;
        }
        __syncthreads();
        const int64_t i0 = t0 + (int64_t)threadIdx.x * MD_ITEMS;
        if (i0 < p.n) {
            int c = md_find(p.bin_off, p.C, i0);
            int64_t cend = p.bin_off[c + 1];
            MedState st = p.state[c];
            // run-length aggregation: (target histogram, digit) of the previous bin
            int runA_d = -1, runC_d = -1, runC_c = c;
            u32 runA_n = 0, runC_n = 0;
            u64 nextA = ~0ull, nextC = ~0ull;
            auto flushC = [&]() {
                if (runC_n) {
                    if (runC_c == c0) atomicAdd(&hCtg[runC_d], runC_n);
                    else atomicAdd(p.hist + (int64_t)runC_c * 256 + runC_d, runC_n);
                }
                runC_n = 0;
            };
            auto flushNextC = [&]() {
                if (NEXT && nextC != ~0ull) {
                    if (c == c0) atomicMin(&sNextCtg, nextC);
                    else atomicMin(&p.state[c].next, nextC);
                }
                nextC = ~0ull;
            };
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
#pragma unroll
            for (int k = 0; k < MD_ITEMS; k++) {
                const int64_t i = i0 + k;
                if (i >= p.n) continue;
                if (i >= cend) {
                    flushC();
                    flushNextC();
                    while (i >= p.bin_off[c + 1]) c++;
                    cend = p.bin_off[c + 1];
                    st = p.state[c];
                    runC_c = c;
                    runC_d = -1;
                }
                const double v = vv[k];
                if (!(v > 0.0) || gg[k] == -1) continue;              // tiddit_coverage_analysis.pyx:18
                const u64 key = (u64)__double_as_longlong(v);
                if (NEXT) {
                    if (key > stAll.prefix && key < nextA) nextA = key;
                    if (key > st.prefix && key < nextC) nextC = key;
                    continue;
                }
                const int d = (int)((key >> shift) & 255u);
                const u64 hi = FIRST ? 0ull : (key >> (shift + 8));
                if (FIRST || hi == (stAll.prefix >> (shift + 8))) {
                    if (d == runA_d) runA_n++;
                    else {
                        if (runA_n) atomicAdd(&hAll[runA_d], runA_n);
                        runA_d = d;
                        runA_n = 1;
                    }
                }
                if (FIRST || hi == (st.prefix >> (shift + 8))) {
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
        __syncthreads();
        if (NEXT) {
            if (threadIdx.x == 0) {
                if (sNextAll != ~0ull) atomicMin(&p.state[p.C].next, sNextAll);
                if (sNextCtg != ~0ull) atomicMin(&p.state[c0].next, sNextCtg);
            }
        } else {
            const u32 a = hAll[threadIdx.x], b = hCtg[threadIdx.x];
            if (a) atomicAdd(p.hist + (int64_t)p.C * 256 + threadIdx.x, a);
            if (b) atomicAdd(p.hist + (int64_t)c0 * 256 + threadIdx.x, b);
        }
        __syncthreads();
    }
}

// one warp per segment: pick the digit holding the wanted rank, clear the histogram for the next pass
__global__ void __launch_bounds__(256) med_select_kernel(MedParams p, int shift, int first, int last) {
    const int seg = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    const int lane = threadIdx.x & 31;
    if (seg > p.C) return;
    u32 *h = p.hist + (int64_t)seg * 256;
    u32 v[8];
    u32 mine = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) {          // lane owns digits [8*lane, 8*lane + 8)
        v[k] = h[lane * 8 + k];
        mine += v[k];
        h[lane * 8 + k] = 0;
    }
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
        if (lane == 0) {
            if (first) p.state[seg] = st;
        }
        return;
    }
    const u32 before = inc - mine;   // bins in lower lanes
    const bool owner = (int64_t)before <= st.rank && st.rank < (int64_t)inc;
    if (owner) {
        u32 cum = before;
        int d = 0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (st.rank >= (int64_t)cum + v[k]) cum += v[k];
            else {
                d = k;
                break;
            }
        }
        st.prefix |= (u64)(lane * 8 + d) << shift;
        st.less += cum;
        st.rank -= cum;
        if (last) st.equal = v[d];
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
    return ((size_t)(C + 1) * sizeof(MedState) + 255) / 256 * 256 + (size_t)(C + 1) * 256 * 4 + 256;
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
    const unsigned grid = (unsigned)(tiles < 148 * 8 ? (tiles > 0 ? tiles : 1) : 148 * 8);
    const unsigned sel_grid = (unsigned)(((int64_t)(C + 1) * 32 + 255) / 256);
    ProfScope ps("coverage_medians", st);
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        if (n_bins > 0) {
            if (pass == 0) TDT_LAUNCH((med_pass_kernel<true, false>), grid, MD_THREADS, 0, st, p, shift);
            else TDT_LAUNCH((med_pass_kernel<false, false>), grid, MD_THREADS, 0, st, p, shift);
        }
        TDT_LAUNCH(med_select_kernel, sel_grid, 256, 0, st, p, shift, pass == 0, pass == 7);
    }
    if (n_bins > 0) TDT_LAUNCH((med_pass_kernel<false, true>), grid, MD_THREADS, 0, st, p, 0);
    TDT_LAUNCH(med_finish_kernel, (unsigned)((C + 1 + 255) / 256), 256, 0, st, p);
    return TDT_OK;
}

}  // extern "C"

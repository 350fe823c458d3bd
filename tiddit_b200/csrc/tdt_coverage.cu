// tdt_coverage.cu -- fractional coverage bins on B200 (sm_100a).
//
// Replaces tiddit/tiddit_coverage.pyx:48-74 (update_coverage) applied to a batch of reads.  Every
// addend is the float32 quotient the reference computes (C `float` division, then promoted to
// double for the add) or the integer 1.0; all of them are multiples of a common power of two and
// the bin totals stay far below 2^53 of it, so float64 sums are exact in ANY order: per-thread
// pre-aggregation + RED.ADD.F64 into HBM is bit-identical to the reference's sequential `+=`.
//
// Fast tiles (full tile, one contig, table-sized bin -- all but ~25 tiles of a genome): every thread anchors a
// window of two adjacent bins at its lowest read and tests each read against it with two subtractions and
// unsigned compares -- no division; both addends are table look-ups whose INDEX is selected (0 -> adds 0.0), so a
// read costs ~25 instructions: two LDS.64 and two DADD.  A read outside the window flushes and re-anchors it.
//
// Layout of the work: a CTA stages 8192 reads (start[] and end[], 32 KB each) in shared memory with two 1-D TMA
// bulk copies; every thread then walks its OWN 32 consecutive reads.  Reads arrive coordinate-sorted (BAM
// order), so a run touches a handful of neighbouring bins: the thread keeps a window of two adjacent bins in
// registers and issues atomics only when the window moves; the lanes of a warp work 32 reads apart, so their
// atomics spread over bins instead of piling 32 deep onto one L2 address.  bin = pos / bin_size uses an exact multiply-shift; the quotients
// float(a) / float(bin_size) for a in [0, bin_size] come from a shared-memory table filled with the very
// division the reference performs, so no rounding differs.
#include "tdt_common.cuh"

namespace tdt {

constexpr int COV_THREADS = 256;
constexpr int COV_RUN = 32;            // consecutive reads one thread walks (power of two)
constexpr int COV_TILE = COV_THREADS * COV_RUN;  // reads staged in shared memory per CTA iteration
constexpr int COV_TABLE_MAX = 4096;    // bin sizes up to this use the quotient table (32 KB of doubles)

__device__ __forceinline__ int64_t floordiv_i64(int64_t a, int64_t b) {  // Python //, b > 0
    int64_t q = a / b;
    if ((a % b != 0) && (a < 0)) q--;
    return q;
}

struct CovContig {  // where one contig's reads and bins live (all-contig call)
    const int64_t *read_off;
    const int64_t *bin_off;
    const int32_t *end_bin_size;
    int C;
};

// A thread keeps a window of two adjacent bins [cb, cb+1] of the current contig in registers.  The common read
// (both of its fractional bins inside the window) is four predicated adds; a read that leaves the window
// flushes it (two atomics) and re-opens it at the read's first bin, bins further right go straight to memory.
struct CovAcc {
    int32_t cb;
    double s0, s1;
    double *bins;  // bins of the current contig
    __device__ __forceinline__ void flush() {
        if (cb >= 0) {
            if (s0 != 0.0) atomicAdd(bins + cb, s0);
            if (s1 != 0.0) atomicAdd(bins + cb + 1, s1);
        }
        cb = -1;
        s0 = s1 = 0.0;
    }
    // first bin fb gets v1, last bin eb >= fb gets v2 (0 when eb == fb), bins strictly between get 1.0
    __device__ __forceinline__ void add_read(int32_t fb, double v1, int32_t eb, double v2) {
        const uint32_t df = (uint32_t)(fb - cb), de = (uint32_t)(eb - cb);
        if (cb >= 0 && df <= 1u && de <= 1u) {
            s0 += (df == 0u ? v1 : 0.0) + (de == 0u ? v2 : 0.0);
            s1 += (df == 1u ? v1 : 0.0) + (de == 1u ? v2 : 0.0);
            return;
        }
        flush();
        cb = fb;
        s0 = v1;
        if (eb == fb + 1) {
            s1 = v2;
        } else if (eb > fb + 1) {
            s1 = 1.0;
            for (int32_t b = fb + 2; b < eb; b++) atomicAdd(bins + b, 1.0);
            atomicAdd(bins + eb, v2);
        }
    }
};

// negative start or empty / inverted read: the reference's arithmetic literally, incl. Python's floor division
// and negative-index wrap-around (tiddit_coverage.pyx:50-72); rare, so straight atomics
__device__ __noinline__ void cov_odd_read(int64_t rs, int64_t re, int32_t bin_size, int32_t ebs, double *bins,
                                          int64_t n_bins, int64_t r, unsigned long long *first_bad) {
    const float fbin = (float)bin_size;
    const int64_t fb = floordiv_i64(rs, bin_size), eb = floordiv_i64(re - 1, bin_size);
    if (fb < -n_bins || fb >= n_bins || eb < -n_bins || eb >= n_bins) {
        atomicMin(first_bad, (unsigned long long)r);
        return;
    }
    const int64_t wfb = fb < 0 ? fb + n_bins : fb, web = eb < 0 ? eb + n_bins : eb;
    if (eb == fb) {
        atomicAdd(bins + wfb, (double)__fdiv_rn((float)(re - rs), fbin));
        return;
    }
    atomicAdd(bins + wfb, (double)__fdiv_rn((float)((fb + 1) * (int64_t)bin_size - rs), fbin));
    const float last = (float)((re - 1) - eb * (int64_t)bin_size);
    atomicAdd(bins + web, (double)__fdiv_rn(last, eb < n_bins - 1 ? fbin : (float)ebs));
    for (int64_t b = fb + 1; b < eb; b++) atomicAdd(bins + (b < 0 ? b + n_bins : b), 1.0);
}

// ---- fast tiles ------------------------------------------------------------------------------------------
struct CovWin {      // two adjacent bins [cb, cb+1] of the contig, sums in registers
    double s0, s1;
    int32_t cb;
    uint32_t lim;    // a read with (end-1) - cb*bin < lim stays inside the window: 2*bin, or bin when bin cb+1 is the
                     // contig's last bin (its divisor differs, tiddit_coverage.pyx:66-69) or does not exist
};

struct CovGeom {
    int32_t bin_size, n_bins, ebs;
    uint32_t magic;
    int magic_shift;
    const double *q_table;
    double *bins;    // bins of the tile's contig
};

__device__ __forceinline__ uint32_t cov_lim(int32_t cb, const CovGeom &g) {
    return cb + 1 < g.n_bins - 1 ? 2u * (uint32_t)g.bin_size : (uint32_t)g.bin_size;
}

__device__ __forceinline__ void cov_win_flush(const CovWin &w, double *bins) {
    if (w.s0 != 0.0) atomicAdd(bins + w.cb, w.s0);
    if (w.s1 != 0.0) atomicAdd(bins + w.cb + 1, w.s1);
}

// a read that is not inside the window: flush, then the reference's arithmetic with the exact division and the
// window re-anchored at the read's first bin
__device__ __noinline__ CovWin cov_miss(CovWin w, int32_t rs, int32_t re, int64_t r, CovGeom g,
                                        unsigned long long *first_bad) {
    cov_win_flush(w, g.bins);
    w.s0 = w.s1 = 0.0;
    if (!(rs >= 0 && re > rs)) {
        cov_odd_read(rs, re, g.bin_size, g.ebs, g.bins, g.n_bins, r, first_bad);
        return w;
    }
    const int32_t fb = (int32_t)(((uint32_t)rs + __umulhi((uint32_t)rs, g.magic)) >> g.magic_shift);            // :50
    const int32_t eb = (int32_t)(((uint32_t)(re - 1) + __umulhi((uint32_t)(re - 1), g.magic)) >> g.magic_shift);  // :51
    if (eb >= g.n_bins) {  // boundscheck: IndexError in the reference
        atomicMin(first_bad, (unsigned long long)r);
        return w;
    }
    const bool same = eb == fb;
    const int a1 = same ? re - rs : (fb + 1) * g.bin_size - rs;
    const int a2 = same ? 0 : (re - 1) - eb * g.bin_size;
    double v2 = g.q_table[a2];
    if (!same && eb == g.n_bins - 1) v2 = (double)__fdiv_rn((float)a2, (float)g.ebs);
    w.cb = fb;
    w.lim = cov_lim(fb, g);
    w.s0 = g.q_table[a1];
    if (eb == fb + 1) {
        w.s1 = v2;
    } else if (eb > fb + 1) {
        w.s1 = 1.0;
        for (int32_t b = fb + 2; b < eb; b++) atomicAdd(g.bins + b, 1.0);
        atomicAdd(g.bins + eb, v2);
    }
    return w;
}

__device__ __forceinline__ void cov_tile_fast(const int32_t *s_start, const int32_t *s_end, int64_t t0, const CovGeom &g,
                                              unsigned long long *first_bad) {
    const int t = threadIdx.x, lane = t & 31;
    const uint32_t bin = (uint32_t)g.bin_size;
    // anchor the window at the first bin of the thread's lowest read (reads are coordinate-sorted)
    const int32_t rs0 = s_start[t * COV_RUN];
    CovWin w;
    w.s0 = w.s1 = 0.0;
    w.cb = rs0 > 0 ? (int32_t)(((uint32_t)rs0 + __umulhi((uint32_t)rs0, g.magic)) >> g.magic_shift) : 0;
    if (w.cb > g.n_bins - 1) w.cb = g.n_bins - 1;
    w.lim = cov_lim(w.cb, g);
    uint32_t wb = (uint32_t)w.cb * bin;
    const int4 *S4 = (const int4 *)s_start + t * (COV_RUN / 4);
    const int4 *E4 = (const int4 *)s_end + t * (COV_RUN / 4);
#pragma unroll 2
    for (int k4 = 0; k4 < COV_RUN / 4; k4++) {
        // 16-byte shared loads, rotated by the lane so that the 8 lanes of a quarter warp hit 8 different bank groups
        const int q = (k4 + lane) & (COV_RUN / 4 - 1);
        const int4 S = S4[q], E = E4[q];
        const int32_t rsv[4] = {S.x, S.y, S.z, S.w}, rev[4] = {E.x, E.y, E.z, E.w};
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const int32_t rs = rsv[j], re = rev[j];
            const uint32_t os = (uint32_t)rs - wb, oe = (uint32_t)re - 1u - wb;
            if (os <= oe && oe < w.lim) {
                const bool df = os >= bin, de = oe >= bin;        // in the window's second bin?
                const uint32_t len = (uint32_t)(re - rs);
                const uint32_t i0 = df ? 0u : (de ? bin - os : len);   // :56 / :61 -> bin cb
                const uint32_t i1 = de ? (df ? len : oe - bin) : 0u;   // :56 / :63 -> bin cb+1   (q_table[0] = 0)
                w.s0 += g.q_table[i0];
                w.s1 += g.q_table[i1];
            } else {
                w = cov_miss(w, rs, re, t0 + t * COV_RUN + q * 4 + j, g, first_bad);
                wb = (uint32_t)w.cb * bin;
            }
        }
    }
    cov_win_flush(w, g.bins);
}

template <bool MULTI>
__global__ void __launch_bounds__(COV_THREADS) coverage_kernel(const int32_t *__restrict__ start,
                                                               const int32_t *__restrict__ end, int64_t n_reads,
                                                               int32_t bin_size, int32_t end_bin_size_one,
                                                               double *__restrict__ bins, int64_t n_bins_one,
                                                               CovContig cc, uint32_t magic, int magic_shift,
                                                               bool vec_ok, unsigned long long *first_bad) {
    extern __shared__ __align__(128) unsigned char cov_smem[];
    __shared__ __align__(8) uint64_t mbar;
    int32_t *s_start = (int32_t *)cov_smem;                 // [COV_TILE]
    int32_t *s_end = s_start + COV_TILE;                    // [COV_TILE]
    double *q_table = (double *)(s_end + COV_TILE);         // q_table[a] = (double)((float)a / (float)bin_size)
    const bool use_table = bin_size <= COV_TABLE_MAX;
    const float fbin = (float)bin_size;
    if (use_table)
        for (int a = threadIdx.x; a <= bin_size; a += COV_THREADS) q_table[a] = (double)__fdiv_rn((float)a, fbin);
    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    CovAcc acc;
    acc.cb = -1;
    acc.s0 = acc.s1 = 0.0;
    acc.bins = bins;
    int32_t n_bins = (int32_t)n_bins_one, ebs = end_bin_size_one;
    int64_t c_first = 0, c_next = MULTI ? 0 : INT64_MAX;  // reads [c_first, c_next) belong to the current contig
    const int lane = threadIdx.x & 31;
    uint32_t phase = 0;

    const int64_t n_tiles = (n_reads + COV_TILE - 1) / COV_TILE;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t t0 = tile * COV_TILE;
        const int cnt = n_reads - t0 < COV_TILE ? (int)(n_reads - t0) : COV_TILE;
        // stage the tile: two 32 KB TMA bulk copies (full, 16-byte aligned tiles) or plain coalesced loads
        if (vec_ok && cnt == COV_TILE) {
            if (threadIdx.x == 0) {
                mbar_expect_tx(&mbar, 2u * COV_TILE * 4u);
                tma_load_1d(s_start, start + t0, COV_TILE * 4u, &mbar);
                tma_load_1d(s_end, end + t0, COV_TILE * 4u, &mbar);
            }
            mbar_wait(&mbar, phase);
            phase ^= 1u;
        } else {
            for (int i = threadIdx.x; i < cnt; i += COV_THREADS) {
                s_start[i] = start[t0 + i];
                s_end[i] = end[t0 + i];
            }
            __syncthreads();
        }
        bool fast = use_table && cnt == COV_TILE;
        if (MULTI && fast && !(t0 >= c_first && t0 + cnt <= c_next)) {
            // the contig of the tile's first read (the same look-up in every thread)
            acc.flush();
            int lo = 0, hi = cc.C;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (cc.read_off[mid] <= t0) lo = mid; else hi = mid;
            }
            c_first = cc.read_off[lo];
            c_next = cc.read_off[lo + 1];
            const int64_t base = cc.bin_off[lo];
            n_bins = (int32_t)(cc.bin_off[lo + 1] - base);
            ebs = cc.end_bin_size[lo];
            acc.bins = bins + base;
            fast = t0 + cnt <= c_next;
        }
        if (fast && n_bins >= 1) {
            CovGeom g;
            g.bin_size = bin_size;
            g.n_bins = n_bins;
            g.ebs = ebs;
            g.magic = magic;
            g.magic_shift = magic_shift;
            g.q_table = q_table;
            g.bins = acc.bins;
            cov_tile_fast(s_start, s_end, t0, g, first_bad);
            __syncthreads();
            continue;
        }
        // a thread walks its own COV_RUN consecutive reads (rotated by the lane so that the 32 lanes hit 32
        // different banks): its register window absorbs nearly all adds, and the lanes of a warp work COV_RUN
        // reads apart, so their atomics go to different bins instead of piling up on one L2 address
#pragma unroll 2
        for (int k = 0; k < COV_RUN; k++) {
            const int idx = threadIdx.x * COV_RUN + ((k + lane) & (COV_RUN - 1));
            if (idx >= cnt) continue;
            const int64_t r = t0 + idx;
            if (MULTI && (r >= c_next || r < c_first)) {
                // another contig: close the window, look the contig up (largest c with read_off[c] <= r)
                acc.flush();
                int lo = 0, hi = cc.C;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (cc.read_off[mid] <= r) lo = mid; else hi = mid;
                }
                c_first = cc.read_off[lo];
                c_next = cc.read_off[lo + 1];
                const int64_t base = cc.bin_off[lo];
                n_bins = (int32_t)(cc.bin_off[lo + 1] - base);
                ebs = cc.end_bin_size[lo];
                acc.bins = bins + base;
            }
            const int32_t rs = s_start[idx], re = s_end[idx];
            if (rs >= 0 && re > rs) {
                // the usual read: exact multiply-shift division (n < 2^31), table quotients
                const int32_t fb = (int32_t)(((uint32_t)rs + __umulhi((uint32_t)rs, magic)) >> magic_shift);      // :50
                const int32_t eb =
                    (int32_t)(((uint32_t)(re - 1) + __umulhi((uint32_t)(re - 1), magic)) >> magic_shift);        // :51
                if (eb >= n_bins) {  // boundscheck: IndexError in the reference
                    atomicMin(first_bad, (unsigned long long)r);
                } else {
                    const bool same = eb == fb;                                               // :55
                    const int a1 = same ? re - rs : (fb + 1) * bin_size - rs;                 // :56 / :61
                    const int a2 = same ? 0 : (re - 1) - eb * bin_size;                       // :63 (q_table[0] = 0)
                    double v1, v2;
                    if (use_table) {
                        v1 = q_table[a1];
                        v2 = q_table[a2];
                    } else {
                        v1 = (double)__fdiv_rn((float)a1, fbin);
                        v2 = (double)__fdiv_rn((float)a2, fbin);
                    }
                    if (!same && eb == n_bins - 1) v2 = (double)__fdiv_rn((float)a2, (float)ebs);   // :69
                    acc.add_read(fb, v1, eb, v2);
                }
            } else {
                cov_odd_read(rs, re, bin_size, ebs, acc.bins, n_bins, r, first_bad);
            }
        }
        acc.flush();
        __syncthreads();  // the tile is consumed: the next bulk copy may overwrite it
    }
}

// n / d == (n + umulhi(n, magic)) >> shift for every 0 <= n < 2^31: Granlund-Montgomery with
// M = ceil(2^(32+L) / d) = 2^32 + magic, L = ceil(log2 d) (n + hi(n * magic) < 2^32 because n < 2^31)
static void magic_for(uint32_t d, uint32_t &magic, int &shift) {
    int L = 0;
    while ((1ull << L) < d) L++;
    shift = L;
    const unsigned __int128 one = (unsigned __int128)1 << (32 + L);
    const unsigned __int128 M = (one + d - 1) / d;
    magic = (uint32_t)(M - ((unsigned __int128)1 << 32));
}

static int launch_coverage(bool multi, const int32_t *start, const int32_t *end, int64_t n_reads, int32_t bin_size,
                           int32_t end_bin_size, double *bins, int64_t n_bins, CovContig cc, int64_t *first_bad,
                           cudaStream_t st) {
    const bool vec_ok = (((uintptr_t)start | (uintptr_t)end) & 15) == 0;
    uint32_t magic;
    int shift;
    magic_for((uint32_t)bin_size, magic, shift);
    int64_t blocks = (n_reads + COV_TILE - 1) / COV_TILE;
    if (blocks > 148 * 3) blocks = 148 * 3;
    const size_t smem = (size_t)COV_TILE * 8 + (bin_size <= COV_TABLE_MAX ? ((size_t)bin_size + 1) * sizeof(double) : 8);
    static thread_local bool configured = false;
    if (!configured) {
        const int max_smem = COV_TILE * 8 + (COV_TABLE_MAX + 1) * 8;
        TDT_CUDA(cudaFuncSetAttribute(coverage_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        TDT_CUDA(cudaFuncSetAttribute(coverage_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem));
        configured = true;
    }
    ProfScope ps("coverage", st);
    if (multi)
        TDT_LAUNCH(coverage_kernel<true>, (unsigned)blocks, COV_THREADS, smem, st, start, end, n_reads, bin_size, 0, bins,
                   (int64_t)0, cc, magic, shift, vec_ok, (unsigned long long *)first_bad);
    else
        TDT_LAUNCH(coverage_kernel<false>, (unsigned)blocks, COV_THREADS, smem, st, start, end, n_reads, bin_size,
                   end_bin_size, bins, n_bins, cc, magic, shift, vec_ok, (unsigned long long *)first_bad);
    return TDT_OK;
}

}  // namespace tdt

using namespace tdt;

extern "C" {

int tdt_coverage_accumulate(const int32_t *start, const int32_t *end, int64_t n_reads, int32_t bin_size,
                            int32_t end_bin_size, double *bins, int64_t n_bins, int64_t *first_bad, void *stream) {
    if (n_reads < 0 || n_bins < 0) return fail(TDT_E_ARG, "negative size");
    if (bin_size <= 0) return fail(TDT_E_ARG, "bin_size = %d (the reference divides by it)", bin_size);
    if (n_reads == 0) return TDT_OK;
    if (!start || !end || !bins || !first_bad) return fail(TDT_E_ARG, "null pointer argument");
    CovContig cc = {nullptr, nullptr, nullptr, 0};
    return launch_coverage(false, start, end, n_reads, bin_size, end_bin_size, bins, n_bins, cc, first_bad,
                           (cudaStream_t)stream);
}

int tdt_coverage_accumulate_contigs(const int32_t *start, const int32_t *end, int64_t n_reads, const int64_t *read_off,
                                    const int64_t *bin_off, const int32_t *end_bin_size, int32_t C, int32_t bin_size,
                                    double *bins, int64_t n_bins_total, int64_t *first_bad, void *stream) {
    if (n_reads < 0 || n_bins_total < 0 || C < 0) return fail(TDT_E_ARG, "negative size");
    if (bin_size <= 0) return fail(TDT_E_ARG, "bin_size = %d (the reference divides by it)", bin_size);
    if (n_reads == 0) return TDT_OK;
    if (C < 1) return fail(TDT_E_ARG, "C = %d contigs for %lld reads", C, (long long)n_reads);
    if (!start || !end || !bins || !first_bad || !read_off || !bin_off || !end_bin_size)
        return fail(TDT_E_ARG, "null pointer argument");
    CovContig cc = {read_off, bin_off, end_bin_size, C};
    return launch_coverage(true, start, end, n_reads, bin_size, 0, bins, 0, cc, first_bad, (cudaStream_t)stream);
}

}  // extern "C"

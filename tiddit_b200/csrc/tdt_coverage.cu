// tdt_coverage.cu -- fractional coverage bins on B200 (sm_100a).
//
// Replaces tiddit/tiddit_coverage.pyx:48-74 (update_coverage) applied to a batch of reads.  Every
// addend is the float32 quotient the reference computes (C `float` division, then promoted to
// double for the add) or the integer 1.0; all of them are multiples of a common power of two and
// the bin totals stay far below 2^53 of it, so float64 sums are exact in ANY order: warp-level
// pre-aggregation + RED.ADD.F64 into HBM is bit-identical to the reference's sequential `+=`.
//
// Reads arrive coordinate-sorted (BAM order), so the 32 reads of a warp touch one or two distinct
// bins: lanes with the same bin form runs, each run is summed with a segmented shuffle scan and its
// last lane issues ONE atomic -- ~4 atomics per 32 reads instead of 64+.
#include "tdt_common.cuh"

namespace tdt {

__device__ __forceinline__ int64_t floordiv_i64(int64_t a, int64_t b) {  // Python //, b > 0
    int64_t q = a / b;
    if ((a % b != 0) && (a < 0)) q--;
    return q;
}

// one atomic per run of equal `bin` among adjacent active lanes; v summed exactly in double
__device__ __forceinline__ void warp_run_add(double *bins, int64_t bin, double v, bool active) {
    const int lane = threadIdx.x & 31;
    const u32 act = __ballot_sync(0xffffffffu, active);
    if (act == 0u) return;
    const int64_t prev = __shfl_up_sync(0xffffffffu, bin, 1);
    const bool head = active && (lane == 0 || !((act >> (lane - 1)) & 1u) || prev != bin);
    const u32 heads = __ballot_sync(0xffffffffu, head);
    const int hl = 31 - __clz((int)(heads & lanemask_le()));  // my run's first lane (valid when active)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, v, o);
        if (active && lane - o >= hl) v += t;
    }
    const bool tail = active && (lane == 31 || !((act >> (lane + 1)) & 1u) || ((heads >> (lane + 1)) & 1u));
    if (tail) atomicAdd(bins + bin, v);
}

// the same for addends that are all 1.0: the run sum is the run length
__device__ __forceinline__ void warp_run_add_one(double *bins, int64_t bin, bool active) {
    const int lane = threadIdx.x & 31;
    const u32 act = __ballot_sync(0xffffffffu, active);
    if (act == 0u) return;
    const int64_t prev = __shfl_up_sync(0xffffffffu, bin, 1);
    const bool head = active && (lane == 0 || !((act >> (lane - 1)) & 1u) || prev != bin);
    const u32 heads = __ballot_sync(0xffffffffu, head);
    const int hl = 31 - __clz((int)(heads & lanemask_le()));
    const bool tail = active && (lane == 31 || !((act >> (lane + 1)) & 1u) || ((heads >> (lane + 1)) & 1u));
    if (tail) atomicAdd(bins + bin, (double)(lane - hl + 1));
}

struct CovContig {  // where one contig's reads and bins live (all-contig call)
    const int64_t *read_off;
    const int64_t *bin_off;
    const int32_t *end_bin_size;
    int C;
};

template <bool MULTI>
__global__ void __launch_bounds__(256) coverage_kernel(const int32_t *__restrict__ start,
                                                       const int32_t *__restrict__ end, int64_t n_reads,
                                                       int32_t bin_size, int32_t end_bin_size_one,
                                                       double *__restrict__ bins, int64_t n_bins_one, CovContig cc,
                                                       unsigned long long *first_bad) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t n_round = (n_reads + 31) & ~(int64_t)31;  // whole warps stay converged
    const float fbin = (float)bin_size;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n_round; r += stride) {
        bool valid = r < n_reads;
        int64_t fb = 0, eb = 0, base = 0, n_bins = n_bins_one;
        double v1 = 0.0, v2 = 0.0;
        bool two = false;
        if (valid) {
            const int64_t rs = start[r], re = end[r];
            int32_t ebs = end_bin_size_one;
            if (MULTI) {
                int lo = 0, hi = cc.C;  // contig of read r: largest c with read_off[c] <= r
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (cc.read_off[mid] <= r) lo = mid; else hi = mid;
                }
                base = cc.bin_off[lo];
                n_bins = cc.bin_off[lo + 1] - base;
                ebs = cc.end_bin_size[lo];
            }
            fb = floordiv_i64(rs, bin_size);        // tiddit_coverage.pyx:50
            eb = floordiv_i64(re - 1, bin_size);    // :51
            if (eb == fb) {                         // :55-57
                v1 = (double)__fdiv_rn((float)(re - rs), fbin);
            } else {                                // :61-69
                two = true;
                v1 = (double)__fdiv_rn((float)((fb + 1) * (int64_t)bin_size - rs), fbin);
                const float last = (float)((re - 1) - eb * (int64_t)bin_size);
                v2 = (double)__fdiv_rn(last, eb < n_bins - 1 ? fbin : (float)ebs);
            }
            // boundscheck + wraparound are on for this function in the reference: i in [-n, n) only
            if (fb < -n_bins || fb >= n_bins || eb < -n_bins || eb >= n_bins) {
                atomicMin(first_bad, (unsigned long long)r);
                valid = false;
            }
        }
        const int64_t wfb = base + (fb < 0 ? fb + n_bins : fb);
        const int64_t web = base + (eb < 0 ? eb + n_bins : eb);
        warp_run_add(bins, wfb, v1, valid);
        warp_run_add(bins, web, v2, valid && two);
        // bins strictly between first and last get 1.0 each (:71-72)
        for (int64_t k = 1;; k++) {
            const int64_t i = fb + k;
            const bool mid = valid && two && i < eb;
            if (!__any_sync(0xffffffffu, mid)) break;
            warp_run_add_one(bins, base + (i < 0 ? i + n_bins : i), mid);
        }
    }
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
    int64_t blocks = (n_reads + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    CovContig cc = {nullptr, nullptr, nullptr, 0};
    ProfScope ps("coverage", (cudaStream_t)stream);
    TDT_LAUNCH(coverage_kernel<false>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, start, end, n_reads, bin_size,
               end_bin_size, bins, n_bins, cc, (unsigned long long *)first_bad);
    return TDT_OK;
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
    int64_t blocks = (n_reads + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    CovContig cc = {read_off, bin_off, end_bin_size, C};
    ProfScope ps("coverage", (cudaStream_t)stream);
    TDT_LAUNCH(coverage_kernel<true>, (unsigned)blocks, 256, 0, (cudaStream_t)stream, start, end, n_reads, bin_size, 0,
               bins, 0, cc, (unsigned long long *)first_bad);
    return TDT_OK;
}

}  // extern "C"

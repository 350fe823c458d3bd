// tdt_gc.cu -- GC-content bins on B200 (sm_100a).
//
// Replaces tiddit/tiddit_gc.pyx:6-33 (binned_gc) on a contig sequence resident in HBM, one byte per
// base.  Per bin: n = #{N,n}, gc = #{C,c,G,g}, chars = bin length (short only for the last bin);
// out = -1 if n / bin_size > n_cutoff else rint(100 * gc / chars)   (Python round() = half-to-even).
//
// Bytes are classified four at a time in 32-bit words (case folded by |0x20, exact zero-byte test),
// so the kernel stays HBM-bound at 1 byte per base:
//   small bins (<= GC_SMALL_MAX bytes): a CTA stages ~40 KB of sequence (a multiple of 16 bins) in
//       shared memory with 1-D TMA bulk copies and every thread walks its own bins word by word;
//   large bins: one warp per bin, 16-byte coalesced loads straight from HBM.
#include "tdt_common.cuh"

namespace tdt {

#ifndef TDT_GC_EDGE_MASKS
#define TDT_GC_EDGE_MASKS 0
#endif
constexpr int GC_THREADS = 256;
constexpr int GC_SMALL_MAX = 192;      // thread-per-bin up to this bin size
constexpr int GC_TILE_BYTES = 40 * 1024;  // + static shared memory stays under the 48 KB default limit
constexpr uint32_t GC_TMA_CHUNK = 32768;

// 0x80 in every byte of t that is zero, 0 elsewhere (exact, no borrow between bytes)
__device__ __forceinline__ u32 zero_bytes(u32 t) { return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu); }

// valid: 0x80 in every byte that belongs to the bin
__device__ __forceinline__ void classify_word(u32 w, u32 valid, int &n, int &gc) {
    const u32 u = w | 0x20202020u;  // 'N'|0x20 = 'n' ...; only the two cases of a letter map onto it
    n += __popc(zero_bytes(u ^ 0x6e6e6e6eu) & valid);
    gc += __popc((zero_bytes(u ^ 0x63636363u) | zero_bytes(u ^ 0x67676767u)) & valid);
}

// 0x80 flags for bytes [lo, hi) of a word, 0 <= lo <= hi <= 4
__device__ __forceinline__ u32 byte_range_mask(int lo, int hi) {
    const u32 upto_hi = hi >= 4 ? 0xffffffffu : ((1u << (8 * hi)) - 1u);
    const u32 upto_lo = lo >= 4 ? 0xffffffffu : ((1u << (8 * lo)) - 1u);
    return (upto_hi & ~upto_lo) & 0x80808080u;
}

__device__ __forceinline__ int8_t gc_value(int n, int gc, int chars, int bin_size, double n_cutoff) {
    if (__ddiv_rn((double)n, (double)bin_size) > n_cutoff) return (int8_t)-1;      // tiddit_gc.pyx:27-28
    return (int8_t)(int)rint(__ddiv_rn((double)(100 * gc), (double)chars));        // :30
}

__global__ void __launch_bounds__(GC_THREADS) gc_small_kernel(const uint8_t *__restrict__ seq, int64_t len,
                                                              int32_t bin_size, double n_cutoff, int64_t n_bins,
                                                              int bins_per_cta, int8_t *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char tile[];
    __shared__ __align__(8) uint64_t mbar;
    const int64_t bin0 = (int64_t)blockIdx.x * bins_per_cta;  // a multiple of 16: tiles start 16-byte aligned
    const int64_t byte0 = bin0 * bin_size;
    int64_t bytes = (int64_t)bins_per_cta * bin_size;
    if (byte0 + bytes > len) bytes = len - byte0;
    if (threadIdx.x == 0) {
        mbar_init(&mbar, 1);
        fence_mbar_init();
        const uint32_t padded = (uint32_t)((bytes + 15) & ~(int64_t)15);  // the allocation is padded to 16
        mbar_expect_tx(&mbar, padded);
        for (uint32_t off = 0; off < padded; off += GC_TMA_CHUNK) {
            const uint32_t l = padded - off < GC_TMA_CHUNK ? padded - off : GC_TMA_CHUNK;
            tma_load_1d(tile + off, seq + byte0 + off, l, &mbar);
        }
    }
    __syncthreads();
    mbar_wait(&mbar, 0);
    const u32 *words = (const u32 *)tile;
    for (int b = threadIdx.x; b < bins_per_cta; b += GC_THREADS) {
        const int64_t bin = bin0 + b;
        if (bin >= n_bins) break;
        const int lo = b * bin_size;
        int hi = lo + bin_size;
        if ((int64_t)hi > bytes) hi = (int)bytes;
        int n = 0, gc = 0;
#if TDT_GC_EDGE_MASKS
        // PREPARED FOR THE NEXT ROUND, NOT YET RUN ON A GPU (default off): only the first and the last word of a bin
        // need a byte mask; the words between them are whole.  The r01 kernel is bound by these instructions (SM 77 %
        // busy, 33 % of the HBM peak), and the mask arithmetic is about a third of them.
        const int w0 = lo >> 2, w1 = (hi - 1) >> 2;
        {
            const int z = hi < w0 * 4 + 4 ? hi - w0 * 4 : 4;
            classify_word(words[w0], byte_range_mask(lo - w0 * 4, z), n, gc);
        }
        for (int w = w0 + 1; w < w1; w++) classify_word(words[w], 0x80808080u, n, gc);
        if (w1 > w0) classify_word(words[w1], byte_range_mask(0, hi - w1 * 4), n, gc);
#else
        for (int w = lo >> 2; w <= (hi - 1) >> 2; w++) {
            const int wlo = w * 4;
            const int a = lo > wlo ? lo - wlo : 0;
            const int z = hi < wlo + 4 ? hi - wlo : 4;
            classify_word(words[w], byte_range_mask(a, z), n, gc);
        }
#endif
        out[bin] = gc_value(n, gc, hi - lo, bin_size, n_cutoff);
    }
}

__global__ void __launch_bounds__(GC_THREADS) gc_large_kernel(const uint8_t *__restrict__ seq, int64_t len,
                                                              int32_t bin_size, double n_cutoff, int64_t n_bins,
                                                              int8_t *__restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * GC_THREADS + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * GC_THREADS) >> 5;
    for (int64_t bin = warp0; bin < n_bins; bin += nwarps) {
        const int64_t lo = bin * bin_size;
        int64_t hi = lo + bin_size;
        if (hi > len) hi = len;
        int n = 0, gc = 0;
        // 16-byte chunks covering [lo, hi); seq is 16-byte aligned and padded
        for (int64_t c = (lo & ~(int64_t)15) + 16 * (int64_t)lane; c < hi; c += 16 * 32) {
            const uint4 q = *reinterpret_cast<const uint4 *>(seq + c);
            const u32 ws[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int64_t wlo = c + 4 * k;
                const int a = lo > wlo ? (lo - wlo > 4 ? 4 : (int)(lo - wlo)) : 0;
                const int z = hi < wlo + 4 ? (hi > wlo ? (int)(hi - wlo) : 0) : 4;
                if (z > a) classify_word(ws[k], byte_range_mask(a, z), n, gc);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            n += __shfl_xor_sync(0xffffffffu, n, o);
            gc += __shfl_xor_sync(0xffffffffu, gc, o);
        }
        if (lane == 0) out[bin] = gc_value(n, gc, (int)(hi - lo), bin_size, n_cutoff);
    }
}

}  // namespace tdt

using namespace tdt;

extern "C" {

int tdt_gc_bins(const uint8_t *seq, int64_t len, int32_t bin_size, double n_cutoff, int8_t *out, void *stream) {
    if (len < 0) return fail(TDT_E_ARG, "len = %lld is negative", (long long)len);
    if (bin_size <= 0) return fail(TDT_E_ARG, "bin_size = %d (the reference divides by it)", bin_size);
    if (len == 0) return TDT_OK;
    if (!seq || !out) return fail(TDT_E_ARG, "null pointer argument");
    if (((uintptr_t)seq & 15) != 0) return fail(TDT_E_ARG, "seq must be 16-byte aligned (and padded to 16 bytes)");
    const int64_t n_bins = (len + bin_size - 1) / bin_size;
    cudaStream_t st = (cudaStream_t)stream;
    ProfScope ps("gc_bins", st);
    if (bin_size <= GC_SMALL_MAX) {
        const int bpc = (GC_TILE_BYTES / bin_size) & ~15;  // >= 208 bins for bin_size <= 192
        const size_t smem = (size_t)bpc * bin_size + 16;
        const int64_t blocks = (n_bins + bpc - 1) / bpc;
        TDT_LAUNCH(gc_small_kernel, (unsigned)blocks, GC_THREADS, smem, st, seq, len, bin_size, n_cutoff, n_bins, bpc,
                   out);
    } else {
        int64_t blocks = (n_bins + 7) / 8;
        if (blocks > 148 * 16) blocks = 148 * 16;
        TDT_LAUNCH(gc_large_kernel, (unsigned)blocks, GC_THREADS, 0, st, seq, len, bin_size, n_cutoff, n_bins, out);
    }
    return TDT_OK;
}

}  // extern "C"

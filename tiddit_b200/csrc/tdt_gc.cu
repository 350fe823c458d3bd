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

constexpr int GC_THREADS = 256;
constexpr int GC_SMALL_MAX = 192;      // thread-per-bin up to this bin size
constexpr int GC_TILE_BYTES = 40 * 1024;  // + static shared memory stays under the 48 KB default limit
constexpr uint32_t GC_TMA_CHUNK = 32768;

// 0x80 in every byte of t that is zero, 0 elsewhere (exact, no borrow between bytes)
__device__ __forceinline__ u32 zero_bytes(u32 t) { return ~(((t & 0x7f7f7f7fu) + 0x7f7f7f7fu) | t | 0x7f7f7f7fu); }

// valid: 0x80 in every byte that belongs to the bin (large-bin kernel: every word is masked)
__device__ __forceinline__ void classify_word(u32 w, u32 valid, int &n, int &gc) {
    const u32 u = w | 0x20202020u;  // 'N'|0x20 = 'n' ...; only the two cases of a letter map onto it
    n += __popc(zero_bytes(u ^ 0x6e6e6e6eu) & valid);
    gc += __popc((zero_bytes(u ^ 0x63636363u) | zero_bytes(u ^ 0x67676767u)) & valid);
}

// The small-bin kernel's classification: 0x80 in every byte of w that is N/n (zn) resp. C/c/G/g (zg).
//   (w ^ 'N') & 0x5f  is zero exactly for 'N', 'n' and their copies with bit 7 set; the masked value is <= 0x5f, so
//   adding 0x7f never carries into the next byte and sets bit 7 iff it is non-zero; OR-ing w itself rules the
//   bit-7 copies out.  'C' 'c' 'G' 'g' differ only in bits 5 and 2 -> one test with mask 0x5b.
// Three instructions per class (LOP3, IADD, LOP3) + POPC: the r01 kernel spent ~22 per word and was bound by them
// (SM 77 % busy at 33 % of the HBM peak).
__device__ __forceinline__ void classify_flags(u32 w, u32 &zn, u32 &zg) {
    const u32 sn = ((w ^ 0x4e4e4e4eu) & 0x5f5f5f5fu) + 0x7f7f7f7fu;
    const u32 sg = ((w ^ 0x43434343u) & 0x5b5b5b5bu) + 0x7f7f7f7fu;
    zn = ~(sn | w) & 0x80808080u;
    zg = ~(sg | w) & 0x80808080u;
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

// Thread per bin over a TMA-staged tile.  Only the first and the last word of a bin carry a byte mask, the words
// between them are whole; the two double divisions of gc_value are tabulated per CTA (a full bin has chars =
// bin_size, so both depend on one small integer each) -- only the contig's short last bin evaluates them directly.
__global__ void __launch_bounds__(GC_THREADS) gc_small_kernel(const uint8_t *__restrict__ seq, int64_t len,
                                                              int32_t bin_size, double n_cutoff, int64_t n_bins,
                                                              int bins_per_cta, int8_t *__restrict__ out) {
    extern __shared__ __align__(128) unsigned char tile[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int8_t s_val[GC_SMALL_MAX + 1];     // rint(100 * gc / bin_size)
    __shared__ uint8_t s_masked[GC_SMALL_MAX + 1]; // n / bin_size > n_cutoff
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
    if ((int)threadIdx.x <= bin_size) {   // the tables, while the bulk copy is in flight (bin_size <= 192 < 256)
        const int t = threadIdx.x;
        s_masked[t] = __ddiv_rn((double)t, (double)bin_size) > n_cutoff ? 1 : 0;
        s_val[t] = (int8_t)(int)rint(__ddiv_rn((double)(100 * t), (double)bin_size));
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
        const int w0 = lo >> 2, w1 = (hi - 1) >> 2;
        u32 zn, zg;
        classify_flags(words[w0], zn, zg);
        const u32 first = byte_range_mask(lo - w0 * 4, hi < w0 * 4 + 4 ? hi - w0 * 4 : 4);
        int n = __popc(zn & first), gc = __popc(zg & first);
#pragma unroll 4
        for (int w = w0 + 1; w < w1; w++) {
            classify_flags(words[w], zn, zg);
            n += __popc(zn);
            gc += __popc(zg);
        }
        if (w1 > w0) {
            classify_flags(words[w1], zn, zg);
            const u32 last = byte_range_mask(0, hi - w1 * 4);
            n += __popc(zn & last);
            gc += __popc(zg & last);
        }
        out[bin] = hi - lo == bin_size ? (s_masked[n] ? (int8_t)-1 : s_val[gc])
                                       : gc_value(n, gc, hi - lo, bin_size, n_cutoff);
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
        int bpc = GC_TILE_BYTES / bin_size;   // >= 213 bins for bin_size <= 192
        bpc = bpc >= GC_THREADS ? bpc / GC_THREADS * GC_THREADS : (bpc & ~15);   // whole rounds of the CTA; 16-byte aligned tiles
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

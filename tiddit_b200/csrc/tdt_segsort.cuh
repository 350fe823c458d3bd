// tdt_segsort.cuh -- hand-written segmented, stable LSD radix sort of (u32 key, i32 value) pairs (sm_100a).
//
// The clustering path sorts twice, and both sorts are SEGMENTED: signals by posA inside every
// (chrA,chrB) pair (tiddit_cluster.pyx:152) and x-cluster members by posB inside every x-cluster
// (DBSCAN.py:79-81).  Because the segment is implicit in the position (segment s owns
// [off[s], off[s+1]) before and after the sort) the keys stay 32 bits -- the coordinate alone -- instead of
// the 64-bit (segment, coordinate) composites a flat device-wide sort needs.
//
//   small segments (<= SS_LOCAL_MAX elements; nearly all x-clusters): CTA k takes the segments that START in
//       element window [k*W, (k+1)*W) -- at most W-1+LOCAL_MAX contiguous elements -- and sorts them together
//       in shared memory by (local segment index, key): LSD passes over the key digits, then over the
//       segment index, one global read and one global write per element;
//   large segments: onesweep-style passes over 4096-element tiles that never straddle a segment: digit
//       histograms of all passes up front, then per pass a stable in-tile ranking, a per-digit decoupled
//       look-back over the EARLIER TILES OF THE SAME SEGMENT and a digit-ordered write through shared memory.
//
// Stable ranking (both paths): a warp owns a contiguous run of the tile and walks it 32 elements at a time;
// __match_any_sync groups equal digits, the lowest lane of a group bumps the warp-private digit counter.
#pragma once
#include "tdt_common.cuh"

namespace tdt {

constexpr int SS_THREADS = 256;
constexpr int SS_WARPS = SS_THREADS / 32;
constexpr int SS_LOCAL_MAX = 2048;
constexpr int SS_WINDOW = 2048;
constexpr int SS_LOCAL_CAP = 4096;  // >= SS_WINDOW - 1 + SS_LOCAL_MAX
constexpr int SS_TILE = 4096;
constexpr int SS_MAX_PASSES = 4;
constexpr int SS_ERR_KEY_RANGE = 1;

struct SSLarge {
    int64_t start, size;
    int32_t tile_base, pad;
};

struct SSCounters {
    int32_t n_large, n_tiles;
    uint32_t ticket[SS_MAX_PASSES];
    int32_t pad[2];
};

struct SSLayout {  // carved out of the caller's temp storage
    SSCounters *cnt;
    SSLarge *large;
    int32_t *tile_seg;
    int32_t *win_lo, *win_hi;
    uint32_t *ghist;   // [n_large][SS_MAX_PASSES][256]
    uint32_t *status;  // [n_tiles][256]
    int64_t nlarge_max, tiles_max, nwin;
    size_t zero_bytes;  // leading region (counters + windows + status) that one memset initialises
};

static inline size_t ss_align(size_t b) { return (b + 255) & ~(size_t)255; }

static inline int64_t ss_nlarge_max(int64_t n, int64_t nseg_max) {
    int64_t a = n / (SS_LOCAL_MAX + 1) + 1;
    return a < nseg_max ? a : (nseg_max > 0 ? nseg_max : 1);
}

static inline size_t segsort_temp_bytes(int64_t n, int64_t nseg_max) {
    const int64_t nl = ss_nlarge_max(n, nseg_max);
    const int64_t tiles = n / SS_TILE + nl + 1;
    const int64_t nwin = (n + SS_WINDOW - 1) / SS_WINDOW + 1;
    return ss_align(sizeof(SSCounters)) + ss_align((size_t)nl * sizeof(SSLarge)) + ss_align((size_t)tiles * 4) +
           2 * ss_align((size_t)nwin * 4) + ss_align((size_t)nl * SS_MAX_PASSES * 256 * 4) +
           ss_align((size_t)tiles * 256 * 4) + 1024;
}

static inline SSLayout ss_layout(void *temp, int64_t n, int64_t nseg_max) {
    SSLayout L;
    L.nlarge_max = ss_nlarge_max(n, nseg_max);
    L.tiles_max = n / SS_TILE + L.nlarge_max + 1;
    L.nwin = (n + SS_WINDOW - 1) / SS_WINDOW + 1;
    char *p = (char *)temp;
    L.cnt = (SSCounters *)p;
    p += ss_align(sizeof(SSCounters));
    L.status = (uint32_t *)p;
    p += ss_align((size_t)L.tiles_max * 256 * 4);
    L.win_hi = (int32_t *)p;  // 0xff.. = -1 : "no segment starts here"
    p += ss_align((size_t)L.nwin * 4);
    L.win_lo = (int32_t *)p;  // 0x7f.. : +inf
    p += ss_align((size_t)L.nwin * 4);
    L.large = (SSLarge *)p;
    p += ss_align((size_t)L.nlarge_max * sizeof(SSLarge));
    L.tile_seg = (int32_t *)p;
    p += ss_align((size_t)L.tiles_max * 4);
    L.ghist = (uint32_t *)p;
    L.zero_bytes = (size_t)((char *)L.win_hi - (char *)temp);
    return L;
}

struct SSArgs {
    const uint32_t *keys_in;
    const int32_t *vals_in;  // nullptr: value = element index
    uint32_t *keys_out;
    int32_t *vals_out;
    uint32_t *keys_tmp;
    int32_t *vals_tmp;
    const int64_t *off;   // [nseg + 1], device
    const int64_t *dims;  // device: dims[0] = n, dims[1] = nseg (actual values; the grids use upper bounds)
    int key_bits, n_passes, bits_per_pass;
    SSLayout L;
    int *err;
};

__device__ __forceinline__ uint32_t ss_digit(uint32_t key, int pass, int bits) {
    return (key >> (pass * bits)) & ((1u << bits) - 1u);
}

// ---- classification: windows of the small segments, tile ranges of the large ones -----------------
__global__ void segsort_classify_kernel(SSArgs a) {
    const int64_t nseg = a.dims[1];
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nseg) return;
    const int64_t q = a.off[s], size = a.off[s + 1] - q;
    if (size <= 0) return;
    const int64_t k = q / SS_WINDOW;
    atomicMin(a.L.win_lo + k, (int32_t)s);
    atomicMax(a.L.win_hi + k, (int32_t)s);
    if (size > SS_LOCAL_MAX) {
        const int32_t idx = atomicAdd(&a.L.cnt->n_large, 1);
        const int32_t nt = (int32_t)((size + SS_TILE - 1) / SS_TILE);
        const int32_t tb = atomicAdd(&a.L.cnt->n_tiles, nt);
        SSLarge rec;
        rec.start = q;
        rec.size = size;
        rec.tile_base = tb;
        rec.pad = 0;
        a.L.large[idx] = rec;
        for (int32_t t = 0; t < nt; t++) a.L.tile_seg[tb + t] = idx;
        uint32_t *h = a.L.ghist + (size_t)idx * SS_MAX_PASSES * 256;
        for (int i = 0; i < SS_MAX_PASSES * 256; i++) h[i] = 0u;
    }
}

// ---- one stable LSD pass over `count` elements held in shared memory ---------------------------------
// cur -> alt; the digit comes from the key (use_lid false) or from the local segment index (true).
struct SSLocalBufs {
    uint32_t *K[2];
    int32_t *V[2];
    uint16_t *Lid[2];
    uint32_t (*wh)[256];  // [SS_WARPS][256]
    uint32_t *bin;        // [256] scratch for the digit scan
};

__device__ __forceinline__ uint32_t ss_block_excl_scan_256(uint32_t v, uint32_t *scratch) {
    // exclusive scan over the 256 threads of the CTA (thread = digit); scratch: >= 8 words
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) scratch[warp] = inc;
    __syncthreads();
    uint32_t wbase = 0;
#pragma unroll
    for (int w = 0; w < SS_WARPS; w++)
        if (w < warp) wbase += scratch[w];
    __syncthreads();
    return wbase + inc - v;
}

// count phase: warp-private digit histograms of the warp's own run of the tile
template <typename DigitFn>
__device__ __forceinline__ void ss_count(int count, int epw, uint32_t (*wh)[256], DigitFn digit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < epw; c += 32) {
        const int e = warp * epw + c + lane;
        const bool valid = e < count;
        const uint32_t d = valid ? digit(e) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        if (valid && lane == __ffs(peers) - 1) wh[warp][d] += __popc(peers);
        __syncwarp();
    }
}

// scatter phase: stable position of every element; wh[warp][d] holds the running base of (warp, digit)
template <typename DigitFn, typename EmitFn>
__device__ __forceinline__ void ss_scatter(int count, int epw, uint32_t (*wh)[256], DigitFn digit, EmitFn emit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < epw; c += 32) {
        const int e = warp * epw + c + lane;
        const bool valid = e < count;
        const uint32_t d = valid ? digit(e) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(0xffffffffu, d);
        uint32_t base = 0;
        if (valid) base = wh[warp][d];
        __syncwarp();
        if (valid && lane == __ffs(peers) - 1) wh[warp][d] = base + __popc(peers);
        __syncwarp();
        if (valid) emit(e, base + __popc(peers & lanemask_lt()));
    }
}

// turns the per-warp counts into running bases: wh[w][d] = (#elements with a smaller digit) + (#elements with
// digit d in earlier warps); returns this thread's digit total and exclusive digit base
__device__ __forceinline__ void ss_digit_bases(uint32_t (*wh)[256], uint32_t *scratch, uint32_t &total,
                                               uint32_t &excl) {
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll
    for (int w = 0; w < SS_WARPS; w++) {
        const uint32_t t = wh[w][d];
        wh[w][d] = run;
        run += t;
    }
    total = run;
    excl = ss_block_excl_scan_256(run, scratch);
#pragma unroll
    for (int w = 0; w < SS_WARPS; w++) wh[w][d] += excl;
}

// ---- small segments: whole sort in shared memory ---------------------------------------------------
constexpr size_t SS_LOCAL_SMEM = (size_t)SS_LOCAL_CAP * (4 + 4 + 2) * 2 + SS_WARPS * 256 * 4 + 256 * 4 + 128 * 4 + 64;

__global__ void __launch_bounds__(SS_THREADS) segsort_local_kernel(SSArgs a) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    SSLocalBufs B;
    unsigned char *p = ss_smem;
    B.K[0] = (uint32_t *)p; p += SS_LOCAL_CAP * 4;
    B.K[1] = (uint32_t *)p; p += SS_LOCAL_CAP * 4;
    B.V[0] = (int32_t *)p; p += SS_LOCAL_CAP * 4;
    B.V[1] = (int32_t *)p; p += SS_LOCAL_CAP * 4;
    B.Lid[0] = (uint16_t *)p; p += SS_LOCAL_CAP * 2;
    B.Lid[1] = (uint16_t *)p; p += SS_LOCAL_CAP * 2;
    B.wh = (uint32_t(*)[256])p; p += SS_WARPS * 256 * 4;
    B.bin = (uint32_t *)p; p += 256 * 4;
    uint32_t *heads = (uint32_t *)p;  // [128] head bits, then reused as word prefixes
    __shared__ int64_t s_range[2];
    __shared__ int s_nheads;

    const int64_t n = a.dims[0];
    const int64_t k = blockIdx.x;
    if (k * SS_WINDOW >= n) return;
    const int32_t lo_s = a.L.win_lo[k], hi_s = a.L.win_hi[k];
    if (hi_s < 0) return;  // no segment starts in this window
    if (threadIdx.x == 0) {
        const int64_t begin = a.off[lo_s];
        const int64_t last_start = a.off[hi_s], last_end = a.off[hi_s + 1];
        s_range[0] = begin;
        s_range[1] = (last_end - last_start > SS_LOCAL_MAX) ? last_start : last_end;  // a large one is not ours
    }
    if (threadIdx.x < 128) heads[threadIdx.x] = 0u;
    __syncthreads();
    const int64_t begin = s_range[0], end = s_range[1];
    const int count = (int)(end - begin);
    if (count <= 0) return;

    for (int64_t s = (int64_t)lo_s + threadIdx.x; s <= hi_s; s += SS_THREADS) {
        const int64_t q = a.off[s];
        if (a.off[s + 1] > q && q < end) atomicOr(&heads[(q - begin) >> 5], 1u << ((q - begin) & 31));
    }
    for (int e = threadIdx.x; e < count; e += SS_THREADS) {
        const uint32_t key = a.keys_in[begin + e];
        if (a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
        B.K[0][e] = key;
        B.V[0][e] = a.vals_in ? a.vals_in[begin + e] : (int32_t)(begin + e);
    }
    __syncthreads();
    // local segment index = (#heads at or before e) - 1
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t c[4], t = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            c[i] = __popc(heads[lane * 4 + i]);
            t += c[i];
        }
        uint32_t inc = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        uint32_t ex = inc - t;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            B.bin[lane * 4 + i] = ex;  // heads before word lane*4+i
            ex += c[i];
        }
        if (lane == 31) s_nheads = (int)inc;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < count; e += SS_THREADS)
        B.Lid[0][e] = (uint16_t)(B.bin[e >> 5] + __popc(heads[e >> 5] & (0xffffffffu >> (31 - (e & 31)))) - 1u);
    const int nheads = s_nheads;
    __syncthreads();

    const int epw = ((count + SS_THREADS - 1) / SS_THREADS) * 32;  // elements per warp, multiple of 32
    int cur = 0;
    const int lid_bits = nheads > 1 ? (32 - __clz(nheads - 1)) : 0;
    const int lid_passes = (lid_bits + 7) / 8;
    const int total_passes = a.n_passes + lid_passes;
    for (int pass = 0; pass < total_passes; pass++) {
        for (int i = threadIdx.x; i < SS_WARPS * 256; i += SS_THREADS) (&B.wh[0][0])[i] = 0u;
        __syncthreads();
        const bool on_lid = pass >= a.n_passes;
        const uint32_t *Kc = B.K[cur];
        const uint16_t *Lc = B.Lid[cur];
        const int kp = pass, lp = pass - a.n_passes, bits = a.bits_per_pass;
        auto digit = [&](int e) -> uint32_t {
            return on_lid ? (((uint32_t)Lc[e] >> (8 * lp)) & 255u) : ss_digit(Kc[e], kp, bits);
        };
        ss_count(count, epw, B.wh, digit);
        __syncthreads();
        uint32_t total, excl;
        ss_digit_bases(B.wh, B.bin, total, excl);
        __syncthreads();
        uint32_t *Ka = B.K[cur ^ 1];
        int32_t *Va = B.V[cur ^ 1];
        uint16_t *La = B.Lid[cur ^ 1];
        const int32_t *Vc = B.V[cur];
        ss_scatter(count, epw, B.wh, digit, [&](int e, uint32_t pos) {
            Ka[pos] = Kc[e];
            Va[pos] = Vc[e];
            La[pos] = Lc[e];
        });
        __syncthreads();
        cur ^= 1;
    }
    for (int e = threadIdx.x; e < count; e += SS_THREADS) {
        a.keys_out[begin + e] = B.K[cur][e];
        a.vals_out[begin + e] = B.V[cur][e];
    }
}

// ---- large segments --------------------------------------------------------------------------------
__global__ void __launch_bounds__(SS_THREADS) segsort_hist_kernel(SSArgs a) {
    __shared__ uint32_t h[SS_MAX_PASSES][256];
    const int tile = blockIdx.x;
    if (tile >= a.L.cnt->n_tiles) return;
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

constexpr size_t SS_PASS_SMEM = (size_t)SS_TILE * 4 * 4 + SS_WARPS * 256 * 4 + 256 * 4 + 256 * 8 + 64;

__global__ void __launch_bounds__(SS_THREADS) segsort_pass_kernel(SSArgs a, int pass, const uint32_t *src_k,
                                                                  const int32_t *src_v, uint32_t *dst_k,
                                                                  int32_t *dst_v) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned char *p = ss_smem;
    uint32_t *K = (uint32_t *)p; p += SS_TILE * 4;
    int32_t *V = (int32_t *)p; p += SS_TILE * 4;
    uint32_t *K2 = (uint32_t *)p; p += SS_TILE * 4;
    int32_t *V2 = (int32_t *)p; p += SS_TILE * 4;
    uint32_t(*wh)[256] = (uint32_t(*)[256])p; p += SS_WARPS * 256 * 4;
    uint32_t *bin = (uint32_t *)p; p += 256 * 4;
    int64_t *gbase = (int64_t *)p;
    __shared__ int s_tile;

    if (threadIdx.x == 0) s_tile = (int)atomicAdd(&a.L.cnt->ticket[pass], 1u);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= a.L.cnt->n_tiles) return;
    const int seg = a.L.tile_seg[tile];
    const SSLarge L = a.L.large[seg];
    const int lt = tile - L.tile_base;
    const int64_t t0 = L.start + (int64_t)lt * SS_TILE;
    const int64_t rem = L.start + L.size - t0;
    const int cnt = rem < SS_TILE ? (int)rem : SS_TILE;
    const int bits = a.bits_per_pass;

    for (int e = threadIdx.x; e < cnt; e += SS_THREADS) {
        K[e] = src_k[t0 + e];
        V[e] = src_v ? src_v[t0 + e] : (int32_t)(t0 + e);
    }
    for (int i = threadIdx.x; i < SS_WARPS * 256; i += SS_THREADS) (&wh[0][0])[i] = 0u;
    __syncthreads();
    const int epw = ((cnt + SS_THREADS - 1) / SS_THREADS) * 32;
    auto digit = [&](int e) -> uint32_t { return ss_digit(K[e], pass, bits); };
    ss_count(cnt, epw, wh, digit);
    __syncthreads();
    uint32_t total, excl;
    ss_digit_bases(wh, bin, total, excl);
    {   // per-digit chained scan over the earlier tiles of this segment (thread = digit)
        const int d = threadIdx.x;
        const uint32_t epoch = (uint32_t)pass + 1u;
        uint32_t *row = a.L.status + (size_t)tile * 256 + d;
        uint32_t before = 0;
        if (lt == 0) {
            st_volatile_u32(row, ss_pack(2u, epoch, total));
        } else {
            st_volatile_u32(row, ss_pack(1u, epoch, total));
            for (int t = tile - 1;; t--) {
                const uint32_t *prow = a.L.status + (size_t)t * 256 + d;
                uint32_t s;
                do {
                    s = ld_volatile_u32(prow);
                } while ((s >> 30) == 0u || ((s >> 26) & 15u) != epoch);
                before += s & 0x3ffffffu;
                if ((s >> 30) == 2u) break;
            }
            st_volatile_u32(row, ss_pack(2u, epoch, before + total));
        }
        const uint32_t gh = a.L.ghist[((size_t)seg * SS_MAX_PASSES + pass) * 256 + d];
        gbase[d] = L.start + (int64_t)gh + (int64_t)before - (int64_t)excl;
    }
    __syncthreads();
    ss_scatter(cnt, epw, wh, digit, [&](int e, uint32_t pos) {
        K2[pos] = K[e];
        V2[pos] = V[e];
    });
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += SS_THREADS) {
        const uint32_t key = K2[i];
        const int64_t g = gbase[ss_digit(key, pass, bits)] + i;
        dst_k[g] = key;
        dst_v[g] = V2[i];
    }
}

// ---- host launcher ---------------------------------------------------------------------------------
// Sorts every segment [off[s], off[s+1]) of keys_in/vals_in by key (stable) into keys_out/vals_out.
// n_max / nseg_max: host-side upper bounds that size the grids; the actual n / nseg are read on the device
// from dims[0] / dims[1].  keys_tmp / vals_tmp: scratch of n_max elements (only touched for large segments).
static int segsort_pairs(const uint32_t *keys_in, const int32_t *vals_in, uint32_t *keys_out, int32_t *vals_out,
                         uint32_t *keys_tmp, int32_t *vals_tmp, const int64_t *off, const int64_t *dims, int64_t n_max,
                         int64_t nseg_max, int key_bits, void *temp, size_t temp_bytes, int *err, cudaStream_t st) {
    if (n_max <= 0 || nseg_max <= 0) return TDT_OK;
    if (temp_bytes < segsort_temp_bytes(n_max, nseg_max))
        return fail(TDT_E_WORKSPACE, "segmented sort needs %zu bytes of temporary storage, %zu reserved",
                    segsort_temp_bytes(n_max, nseg_max), temp_bytes);
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
    a.key_bits = key_bits;
    a.n_passes = (key_bits + 7) / 8;
    a.bits_per_pass = (key_bits + a.n_passes - 1) / a.n_passes;
    a.L = ss_layout(temp, n_max, nseg_max);
    a.err = err;
    TDT_CUDA(cudaMemsetAsync(temp, 0, a.L.zero_bytes, st));
    TDT_CUDA(cudaMemsetAsync(a.L.win_hi, 0xff, (size_t)a.L.nwin * 4, st));
    TDT_CUDA(cudaMemsetAsync(a.L.win_lo, 0x7f, (size_t)a.L.nwin * 4, st));
    TDT_LAUNCH(segsort_classify_kernel, (unsigned)((nseg_max + 255) / 256), 256, 0, st, a);
    static thread_local bool configured = false;
    if (!configured) {
        TDT_CUDA(cudaFuncSetAttribute(segsort_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SS_LOCAL_SMEM));
        TDT_CUDA(cudaFuncSetAttribute(segsort_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)SS_PASS_SMEM));
        configured = true;
    }
    const unsigned nwin = (unsigned)((n_max + SS_WINDOW - 1) / SS_WINDOW);
    TDT_LAUNCH(segsort_local_kernel, nwin, SS_THREADS, SS_LOCAL_SMEM, st, a);
    const unsigned tiles = (unsigned)a.L.tiles_max;
    TDT_LAUNCH(segsort_hist_kernel, tiles, SS_THREADS, 0, st, a);
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
    return TDT_OK;
}

}  // namespace tdt

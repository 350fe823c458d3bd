// tdt_common.cuh -- shared plumbing for libtdt_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <atomic>

#include "../../include/tdt_b200.h"

typedef unsigned long long u64;
typedef unsigned int u32;

namespace tdt {

extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

int fail(int code, const char *fmt, ...);

#define TDT_CUDA(expr)                                                                              \
    do {                                                                                            \
        cudaError_t e__ = (expr);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
            return tdt::fail(TDT_E_CUDA, "%s:%d %s -> %s", __FILE__, __LINE__, #expr,               \
                             cudaGetErrorString(e__));                                              \
    } while (0)

// every kernel launch goes through this so that tdt_launch_count() is honest
#define TDT_LAUNCH(kernel, grid, block, smem, stream, ...)                                          \
    do {                                                                                            \
        tdt::prof_kernel_begin(#kernel, (stream));                                                  \
        kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);                                 \
        tdt::prof_kernel_end((stream));                                                             \
        tdt::g_launches.fetch_add(1, std::memory_order_relaxed);                                    \
        TDT_CUDA(cudaGetLastError());                                                               \
    } while (0)

// optional per-stage timing (tdt_profile_begin / tdt_profile_end): CUDA events on the caller's stream around
// each stage of a call; off by default, not thread-safe (one profiled caller at a time)
void prof_kernel_begin(const char *name, cudaStream_t st);   // per-launch timing (TDT_PROF_DETAIL=1)
void prof_kernel_end(cudaStream_t st);
void prof_stage_begin(const char *name, cudaStream_t st);
void prof_stage_end(cudaStream_t st);
struct ProfScope {
    cudaStream_t st;
    ProfScope(const char *name, cudaStream_t s) : st(s) { prof_stage_begin(name, s); }
    ~ProfScope() { prof_stage_end(st); }
};

// bump allocator over the caller's workspace
struct Arena {
    char *base;
    size_t cap, off;
    Arena(void *p, size_t bytes) : base((char *)p), cap(bytes), off(0) {}
    template <typename T>
    T *take(size_t count) {
        size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
        size_t at = off;
        off += bytes;
        if (base == nullptr || off > cap) return nullptr;
        return (T *)(base + at);
    }
};

static inline int bit_width_u32(uint32_t v) {
    int b = 0;
    while (v) { b++; v >>= 1; }
    return b;
}

// ----------------------------------------------------------------------------------------------
// TMA (1-D bulk async copy) + mbarrier, raw PTX.  SASS: UBLKCP / SYNCS.
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    // make the barrier initialisation visible to the async (TMA) proxy
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "TDT_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra TDT_DONE_%=;\n\t"
        "bra TDT_WAIT_%=;\n\t"
        "TDT_DONE_%=:\n\t"
        "}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ u32 lanemask_lt() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ u32 lanemask_le() {
    u32 m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}

}  // namespace tdt

// ----------------------------------------------------------------------------------------------
// Single-pass chained scan ("decoupled look-back") over tiles, two 31-bit sums per tile.
//   status[t] = flag(2 bits) | a(31) | b(31) in ONE 64-bit word, so a reader always sees a
//   consistent (flag, a, b) without fences.  flag 0 = not yet published, 1 = tile aggregate,
//   2 = inclusive prefix.
//   FORWARD PROGRESS.  A tile publishes its own aggregate BEFORE it waits on anybody and only ever waits on tiles
//   with a SMALLER index.  Two ways of numbering tiles are in use:
//     * an atomic ticket (group_extra_scan_kernel, the aggregation scans): every smaller index belongs to a CTA
//       that is already running -- safe whatever order the hardware starts CTAs in;
//     * blockIdx.x (window_runs_kernel / window_runs_small_kernel<., false>, segsort_pass_kernel, s2_pass_kernel):
//       this ASSUMES the hardware starts the CTAs of a grid in blockIdx order (the assumption CUB's DeviceScan made
//       for years): a waiting tile's predecessors then have a lower blockIdx and are resident or finished.  The
//       persistent sort passes (grid capped at the resident-CTA count, CTA b takes tiles b, b + grid, ...) keep it:
//       tile t waits on tiles < t, which belong to CTAs with a smaller blockIdx in the same or an earlier round.
//       The cap is computed for the pass kernel alone; when the small-segment kernels of the same sort run on the
//       forked side stream they may delay, never block, the start of higher-numbered CTAs -- those only wait on
//       lower-numbered ones, which were scheduled first and drain without needing anybody.
// ----------------------------------------------------------------------------------------------
namespace tdt {

__device__ __forceinline__ u64 lb_pack(u32 flag, u32 a, u32 b) { return ((u64)flag << 62) | ((u64)a << 31) | (u64)b; }
__device__ __forceinline__ u32 lb_flag(u64 s) { return (u32)(s >> 62); }
__device__ __forceinline__ u32 lb_a(u64 s) { return (u32)(s >> 31) & 0x7fffffffu; }
__device__ __forceinline__ u32 lb_b(u64 s) { return (u32)s & 0x7fffffffu; }

__device__ __forceinline__ u64 ld_volatile_u64(const u64 *p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile_u64(u64 *p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ __forceinline__ u32 warp_sum(u32 v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Called by ONE full warp of the CTA that owns tile `tile`.  (aggA, aggB): the tile's own sums
// (identical in all lanes).  Returns in (exA, exB) the sums over all earlier tiles.
__device__ __forceinline__ void lookback(u64 *status, int tile, u32 aggA, u32 aggB, u32 &exA, u32 &exB) {
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) st_volatile_u64(status, lb_pack(2, aggA, aggB));
        exA = 0;
        exB = 0;
        return;
    }
    if (lane == 0) st_volatile_u64(status + tile, lb_pack(1, aggA, aggB));
    u32 accA = 0, accB = 0;
    int base = tile - 1;
    while (true) {
        const int idx = base - lane;  // lane 0 looks at the nearest predecessor
        u64 s;
        if (idx < 0) {
            s = lb_pack(2, 0, 0);  // virtual tile before tile 0: inclusive prefix 0
        } else {
            do {
                s = ld_volatile_u64(status + idx);
            } while (lb_flag(s) == 0);
        }
        const u32 incl = __ballot_sync(0xffffffffu, lb_flag(s) == 2);
        const int first = incl ? (__ffs(incl) - 1) : 31;  // nearest tile that already holds a full prefix
        u32 a = lane <= first ? lb_a(s) : 0u;
        u32 b = lane <= first ? lb_b(s) : 0u;
        accA += warp_sum(a);
        accB += warp_sum(b);
        if (incl) break;
        base -= 32;
    }
    if (lane == 0) st_volatile_u64(status + tile, lb_pack(2, accA + aggA, accB + aggB));
    exA = accA;
    exB = accB;
}

}  // namespace tdt

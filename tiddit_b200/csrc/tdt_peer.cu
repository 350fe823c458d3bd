// tdt_peer.cu -- the label exchange of the sharded clustering path over NVLink peer memory (sm_100a).
//
// north_star: "work shards naturally by (chrA,chrB) pair across the 8 GPUs of one box with a single all-gather of
// the final label array over NVLink".  SURVEY 8(b) proposes tdt_cluster_labels_sharded(..., ncclComm_t, ...); the
// exchange itself needs no communicator: every rank owns ONE peer-mapped buffer
//
//     [ gather area: nranks * slot int32 | flags: nranks x 128 B | control: 128 B ]
//
// allocated by tdt_peer_alloc (plain cudaMalloc, exportable with cudaIpcGetMemHandle) and mapped into every other
// rank with tdt_peer_open.  The clustering call writes its labels straight into this rank's slot of its own buffer;
// tdt_peer_allgather then launches ONE kernel that
//   1. copies the slot into the same slot of every peer's buffer with 16-byte stores over NVLink (every thread loads
//      a vector once and stores it nranks-1 times, peers visited in a rank-rotated order so the 8 x 7 streams of an
//      8-GPU box spread over the NVSwitch ports instead of converging on one GPU),
//   2. fences system-wide and counts finished CTAs; the last CTA stores this step's epoch into flag[rank] of every
//      peer (release, system scope) and
//   3. waits (acquire, system scope) until every peer's epoch has arrived in its own flags -- the peers' labels are
//      then complete in the local gather area for everything later in the stream.
// The epoch lives in the buffer's control block (read at kernel start, advanced by the last CTA), so the launch has
// no per-call arguments and can be captured in a CUDA graph.  A peer that never arrives trips a 5 s timeout that
// reports through the caller's status word instead of hanging the device.
//
// The caller keeps ranks in step: a rank must not start exchange k+1 before every rank has CONSUMED the result of
// exchange k (any collective or barrier between steps does; bench.py and engine.sharded_labels have one).
#include "tdt_common.cuh"

namespace tdt {

constexpr int PX_THREADS = 512;
#ifndef TDT_PX_UNROLL
#define TDT_PX_UNROLL 4
#endif
constexpr int PX_UNROLL = TDT_PX_UNROLL;
constexpr int PX_MAX_RANKS = 16;
constexpr int PX_FLAG_STRIDE = 32;   // u32 words (128 B) between two flags
constexpr int PX_ERR_TIMEOUT = 77;

struct PeerArgs {
    int32_t *buf[PX_MAX_RANKS];   // every rank's buffer as mapped HERE (buf[rank] = the local one)
    int64_t slot;                 // int32 elements per rank (multiple of 4)
    int32_t rank, nranks;
    int32_t *status_accum;        // optional
};

__device__ __forceinline__ u32 *px_flags(int32_t *buf, int64_t slot, int nranks) { return (u32 *)(buf + slot * nranks); }
__device__ __forceinline__ u32 *px_ctrl(int32_t *buf, int64_t slot, int nranks) {
    return px_flags(buf, slot, nranks) + (size_t)nranks * PX_FLAG_STRIDE;
}
__device__ __forceinline__ void st_release_sys(u32 *p, u32 v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ u32 ld_acquire_sys(const u32 *p) {
    u32 v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long px_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(PX_THREADS) peer_allgather_kernel(const PeerArgs a) {
    __shared__ u32 s_last;
    int32_t *mine = a.buf[a.rank];
    u32 *ctrl = px_ctrl(mine, a.slot, a.nranks);      // [0] epoch of the last finished exchange, [1] CTA counter
    const u32 epoch = *(volatile u32 *)ctrl + 1u;     // every CTA reads it before the last CTA advances it
    const int64_t nvec = a.slot / 4;
    const uint4 *src = reinterpret_cast<const uint4 *>(mine + a.slot * a.rank);
    const int64_t stride = (int64_t)gridDim.x * PX_THREADS;
    // PX_UNROLL vectors per thread and round: all loads first, then the (nranks - 1) * PX_UNROLL remote stores -- posted
    // writes, so the more of them are in flight per thread the closer the NVLink ports run to their rate
    for (int64_t i0 = (int64_t)blockIdx.x * PX_THREADS + threadIdx.x; i0 < nvec; i0 += stride * PX_UNROLL) {
        uint4 v[PX_UNROLL];
#pragma unroll
        for (int u = 0; u < PX_UNROLL; u++) {
            const int64_t i = i0 + (int64_t)u * stride;
            if (i < nvec) v[u] = src[i];
        }
#pragma unroll 1
        for (int d = 1; d < a.nranks; d++) {
            const int p = (a.rank + d) % a.nranks;
            uint4 *dst = reinterpret_cast<uint4 *>(a.buf[p] + a.slot * a.rank);
#pragma unroll
            for (int u = 0; u < PX_UNROLL; u++) {
                const int64_t i = i0 + (int64_t)u * stride;
                if (i < nvec) dst[i] = v[u];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(ctrl + 1, 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (!s_last) return;
    // the last CTA: every CTA's stores are fenced; publish, then wait for the peers
    __threadfence_system();
    if (threadIdx.x < a.nranks && (int)threadIdx.x != a.rank)
        st_release_sys(px_flags(a.buf[threadIdx.x], a.slot, a.nranks) + (size_t)a.rank * PX_FLAG_STRIDE, epoch);
    if (threadIdx.x < a.nranks && (int)threadIdx.x != a.rank) {
        const u32 *f = px_flags(mine, a.slot, a.nranks) + (size_t)threadIdx.x * PX_FLAG_STRIDE;
        const unsigned long long t0 = px_now_ns();
        while ((int32_t)(ld_acquire_sys(f) - epoch) < 0) {
            if (px_now_ns() - t0 > 5000000000ull) {
                if (a.status_accum) atomicMax(a.status_accum, PX_ERR_TIMEOUT);
                break;
            }
            __nanosleep(200);
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        ctrl[1] = 0u;
        __threadfence();
        *(volatile u32 *)ctrl = epoch;
    }
}

}  // namespace tdt

using namespace tdt;

extern "C" {

size_t tdt_peer_buffer_bytes(int64_t slot_elems, int32_t nranks) {
    if (slot_elems < 0 || nranks < 1) return 0;
    return (size_t)slot_elems * 4 * (size_t)nranks + (size_t)(nranks + 1) * PX_FLAG_STRIDE * 4 + 256;
}

int tdt_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64_h) {
    if (!dev_ptr || !handle64_h) return fail(TDT_E_ARG, "null pointer argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
    void *p = nullptr;
    TDT_CUDA(cudaMalloc(&p, bytes ? bytes : 256));
    TDT_CUDA(cudaMemset(p, 0, bytes ? bytes : 256));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        return fail(TDT_E_CUDA, "cudaIpcGetMemHandle -> %s", cudaGetErrorString(e));
    }
    memcpy(handle64_h, &h, 64);
    *dev_ptr = p;
    return TDT_OK;
}

int tdt_peer_open(const unsigned char *handle64_h, void **dev_ptr) {
    if (!dev_ptr || !handle64_h) return fail(TDT_E_ARG, "null pointer argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64_h, 64);
    TDT_CUDA(cudaIpcOpenMemHandle(dev_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return TDT_OK;
}

int tdt_peer_close(void *dev_ptr) {
    if (dev_ptr) TDT_CUDA(cudaIpcCloseMemHandle(dev_ptr));
    return TDT_OK;
}

int tdt_peer_free(void *dev_ptr) {
    if (dev_ptr) TDT_CUDA(cudaFree(dev_ptr));
    return TDT_OK;
}

int tdt_peer_allgather(void *const *bufs_h, int64_t slot_elems, int32_t rank, int32_t nranks, int32_t *status_accum,
                       void *stream) {
    if (nranks < 1 || nranks > PX_MAX_RANKS) return fail(TDT_E_ARG, "nranks = %d (1..%d supported)", nranks, PX_MAX_RANKS);
    if (rank < 0 || rank >= nranks) return fail(TDT_E_ARG, "rank = %d of %d", rank, nranks);
    if (slot_elems < 0 || (slot_elems & 3)) return fail(TDT_E_ARG, "slot_elems = %lld must be a non-negative multiple of 4", (long long)slot_elems);
    if (!bufs_h) return fail(TDT_E_ARG, "null pointer argument");
    PeerArgs a = {};
    for (int r = 0; r < nranks; r++) {
        if (!bufs_h[r]) return fail(TDT_E_ARG, "buffer of rank %d is null", r);
        a.buf[r] = (int32_t *)bufs_h[r];
    }
    a.slot = slot_elems;
    a.rank = rank;
    a.nranks = nranks;
    a.status_accum = status_accum;
    if (nranks == 1) return TDT_OK;
    static thread_local int sms = 0;
    if (!sms) {
        int dev = 0;
        TDT_CUDA(cudaGetDevice(&dev));
        TDT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    int64_t blocks = (slot_elems / 4 + PX_THREADS - 1) / PX_THREADS;
    if (blocks > sms) blocks = sms;
    if (blocks < 1) blocks = 1;
    TDT_LAUNCH(peer_allgather_kernel, (unsigned)blocks, PX_THREADS, 0, (cudaStream_t)stream, a);
    return TDT_OK;
}

}  // extern "C"

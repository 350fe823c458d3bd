// feasibility probe: global "placement" partition -- count (RED) + scatter (ATOM slot + 8-byte store) + per-bucket finish
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
typedef unsigned int u32;
#define CK(x) do { cudaError_t err__ = (x); if (err__ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(err__)); exit(1); } } while (0)

constexpr int SEG = 1 << 20;       // elements per segment
constexpr int NSEG = 20;
constexpr int NB = 256;            // coarse intervals per segment
constexpr int F = SEG / NB / 16;   // fine buckets per interval (mean occupancy 16)
constexpr int NF = NB * F;         // fine buckets per segment

__device__ __forceinline__ u32 fine_bucket(const u32 *spl, u32 key) {
    u32 b = 0;
#pragma unroll
    for (int step = 128; step > 0; step >>= 1)
        if (spl[b + step - 1] <= key) b += step;
    const u32 lo = b ? spl[b - 1] : 0u, hi = b < 255 ? spl[b] : 0x10000000u;
    u32 f = (u32)((float)(key - lo) * ((float)F / ((float)(hi - lo) + 1.0f)));
    f = f > F - 1 ? F - 1 : f;
    return b * F + f;
}

__global__ void __launch_bounds__(256) count_kernel(const u32 *keys, const u32 *splitters, u32 *cnt, u32 *fbk) {
    __shared__ u32 spl[256];
    const int seg = blockIdx.x / (SEG / 4096);
    spl[threadIdx.x] = splitters[seg * 256 + threadIdx.x];
    __syncthreads();
    const size_t t0 = (size_t)blockIdx.x * 4096;
    u32 k[16], fb[16];
#pragma unroll
    for (int i = 0; i < 16; i++) k[i] = keys[t0 + i * 256 + threadIdx.x];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        fb[i] = seg * NF + fine_bucket(spl, k[i]);
        atomicAdd(cnt + fb[i], 1u);   // RED (result unused)
    }
#pragma unroll
    for (int i = 0; i < 16; i++) fbk[t0 + i * 256 + threadIdx.x] = fb[i];
}

__global__ void scan_kernel(const u32 *cnt, u32 *base) {   // one CTA per segment, exclusive scan of NF counters
    __shared__ u32 part[1024];
    const int seg = blockIdx.x, per = NF / 1024;
    u32 s = 0;
    for (int i = 0; i < per; i++) s += cnt[seg * NF + threadIdx.x * per + i];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        u32 v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    u32 run = part[threadIdx.x] - s;
    for (int i = 0; i < per; i++) {
        base[seg * NF + threadIdx.x * per + i] = run;
        run += cnt[seg * NF + threadIdx.x * per + i];
    }
}

__global__ void __launch_bounds__(256) scatter_kernel(const u32 *keys, const u32 *fbk, const u32 *base, u32 *cursor, uint2 *out) {
    const int seg = blockIdx.x / (SEG / 4096);
    const size_t t0 = (size_t)blockIdx.x * 4096;
    u32 k[16], fb[16], slot[16];
#pragma unroll
    for (int i = 0; i < 16; i++) {
        k[i] = keys[t0 + i * 256 + threadIdx.x];
        fb[i] = fbk[t0 + i * 256 + threadIdx.x];
    }
#pragma unroll
    for (int i = 0; i < 16; i++) slot[i] = atomicAdd(cursor + fb[i], 1u);
#pragma unroll
    for (int i = 0; i < 16; i++)
        out[(size_t)seg * SEG + base[fb[i]] + slot[i]] = make_uint2(k[i], (u32)(t0 + i * 256 + threadIdx.x));
}

__global__ void __launch_bounds__(256) finish_kernel(const uint2 *in, const u32 *cnt, const u32 *base, uint2 *out) {
    const size_t b = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (b >= (size_t)NSEG * NF) return;
    const int seg = (int)(b / NF);
    const u32 c = cnt[b];
    const size_t lo = (size_t)seg * SEG + base[b];
    if (c > 64) return;   // (big buckets: another path)
    for (u32 i = 0; i < c; i++) {
        const uint2 e = in[lo + i];
        u32 r = 0;
        for (u32 j = 0; j < c; j++) {
            const uint2 o = in[lo + j];
            r += (o.x < e.x) || (o.x == e.x && o.y < e.y);
        }
        out[lo + r] = e;
    }
}

int main() {
    const size_t n = (size_t)SEG * NSEG;
    std::vector<u32> h(n), spl(NSEG * 256);
    srand(1);
    for (size_t i = 0; i < n; i++) h[i] = ((u32)rand() * 2654435761u) >> 4;   // 28-bit keys
    // a hotspot in every segment: 5 % of the keys inside 2000 positions
    for (int s = 0; s < NSEG; s++)
        for (int i = 0; i < SEG / 20; i++) h[(size_t)s * SEG + (rand() % SEG)] = 100000000u + rand() % 2000;
    for (int s = 0; s < NSEG; s++) {   // equi-depth splitters from a sample
        std::vector<u32> smp(2048);
        for (int i = 0; i < 2048; i++) smp[i] = h[(size_t)s * SEG + (size_t)i * SEG / 2048];
        std::sort(smp.begin(), smp.end());
        for (int k = 0; k < 256; k++) spl[s * 256 + k] = k < 255 ? smp[(k + 1) * 8] : 0xffffffffu;
    }
    u32 *keys, *dspl, *cnt, *base, *cursor, *fbk;
    uint2 *tmp, *out;
    CK(cudaMalloc(&keys, n * 4)); CK(cudaMalloc(&fbk, n * 4)); CK(cudaMalloc(&dspl, spl.size() * 4));
    CK(cudaMalloc(&cnt, (size_t)NSEG * NF * 4)); CK(cudaMalloc(&base, (size_t)NSEG * NF * 4)); CK(cudaMalloc(&cursor, (size_t)NSEG * NF * 4));
    CK(cudaMalloc(&tmp, n * 8)); CK(cudaMalloc(&out, n * 8));
    CK(cudaMemcpy(keys, h.data(), n * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dspl, spl.data(), spl.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t ev[6];
    for (auto &evt : ev) CK(cudaEventCreate(&evt));
    char *flush; CK(cudaMalloc(&flush, 256 << 20));
    for (int it = 0; it < 4; it++) {
        CK(cudaMemset(flush, it, 256 << 20));
        CK(cudaMemset(cnt, 0, (size_t)NSEG * NF * 4)); CK(cudaMemset(cursor, 0, (size_t)NSEG * NF * 4));
        CK(cudaEventRecord(ev[0]));
        count_kernel<<<n / 4096, 256>>>(keys, dspl, cnt, fbk);
        CK(cudaEventRecord(ev[1]));
        scan_kernel<<<NSEG, 1024>>>(cnt, base);
        CK(cudaEventRecord(ev[2]));
        scatter_kernel<<<n / 4096, 256>>>(keys, fbk, base, cursor, tmp);
        CK(cudaEventRecord(ev[3]));
        finish_kernel<<<(NSEG * NF + 255) / 256, 256>>>(tmp, cnt, base, out);
        CK(cudaEventRecord(ev[4]));
        CK(cudaDeviceSynchronize());
        float t[4];
        for (int i = 0; i < 4; i++) CK(cudaEventElapsedTime(&t[i], ev[i], ev[i + 1]));
        printf("n=%zu count %.1f us  scan %.1f us  scatter %.1f us  finish %.1f us  total %.1f us\n", n, t[0] * 1e3, t[1] * 1e3, t[2] * 1e3, t[3] * 1e3, (t[0] + t[1] + t[2] + t[3]) * 1e3);
    }
    // correctness of the small buckets + occupancy statistics
    std::vector<u32> hc((size_t)NSEG * NF);
    CK(cudaMemcpy(hc.data(), cnt, hc.size() * 4, cudaMemcpyDeviceToHost));
    u32 mx = 0; size_t big = 0, bigel = 0;
    for (u32 c : hc) { mx = std::max(mx, c); if (c > 64) { big++; bigel += c; } }
    printf("fine buckets %zu, max occupancy %u, buckets > 64: %zu holding %zu elements\n", hc.size(), mx, big, bigel);
    std::vector<uint2> ho(n);
    CK(cudaMemcpy(ho.data(), out, n * 8, cudaMemcpyDeviceToHost));
    std::vector<u32> hb((size_t)NSEG * NF);
    CK(cudaMemcpy(hb.data(), base, hb.size() * 4, cudaMemcpyDeviceToHost));
    size_t bad = 0;
    for (int s = 0; s < NSEG; s++)
        for (int f = 0; f < NF; f++) {
            const u32 c = hc[(size_t)s * NF + f];
            if (c > 64) continue;
            const size_t lo = (size_t)s * SEG + hb[(size_t)s * NF + f];
            for (u32 i = 1; i < c; i++) if (ho[lo + i - 1].x > ho[lo + i].x) bad++;
            if (f + 1 < NF && c && hc[(size_t)s * NF + f + 1] && hc[(size_t)s * NF + f + 1] <= 64 && ho[lo + c - 1].x > ho[lo + c].x) bad++;
        }
    printf("order violations: %zu\n", bad);
    return 0;
}

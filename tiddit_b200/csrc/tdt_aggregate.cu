// tdt_aggregate.cu -- candidate aggregation on B200 (sm_100a): cluster labels -> per-candidate statistics.
//
// Replaces the per-signal Python fold of tiddit/tiddit_cluster.pyx:156-254 and the per-candidate
// statistics of :258-336 for ALL (chrA,chrB) pairs in one call.  Per candidate the reference keeps
//   N_discordants / N_splits / N_contigs   sizes of the SETS of read names                       (:261-263)
//   posA, posB                             Counter(...).most_common(1) of the split / contig / discordant
//                                          positions (first-inserted value wins ties) or the orientation-aware
//                                          min / max of the discordant positions                 (:265-330)
//   startA, endA, startB, endB             min / max over all members                            (:332-336)
// and numbers candidates by DBSCAN id, noise surviving only as short intra-chromosomal assembly contigs with
// fresh ids len(pair) + k (:162-168).  Candidates are reported in the reference's dict insertion order (first
// appearance in insertion order), the order tiddit_variant numbers SVs by.
//
// Pipeline (no host synchronisation, data-dependent sizes stay on the device):
//   agg_key_scan         chained scan over insertion order: candidate key per signal (DBSCAN id | surviving noise
//                        contig | dropped), the kept signals compacted (noise never reaches the sort)
//   segsort #1           kept signals grouped by (pair, candidate key), stable => members in insertion order
//   agg_mark_scan        chained scan over the sorted order: group heads, group offsets
//   agg_rank_scan        chained scan over insertion order: rank of every candidate by first appearance
//   agg_gather           per member: (kind,posA) (kind,posB) (kind,name) sort keys; min/max/orientation sums by a
//                        segmented warp reduction, one atomic per (candidate, warp)
//   agg_direct           candidates of <= direct_max members (nearly all; default 256): every member counts its equals
//                        among the candidate's members directly -- mode with first-inserted tie-break as a warp-
//                        segmented max of (count : ~position), distinct names as the members without an earlier
//                        equal.  The members of LARGER candidates were compacted into the "big" arrays by agg_gather;
//                        their sorts run on high-priority side streams NEXT TO this kernel
//   segsort #2..#4       inside every big candidate by (kind,posA), (kind,posB), (kind,name)
//   agg_runs             runs of equal keys: mode with first-inserted tie-break (64-bit atomicMax of
//                        count:~first), distinct-name counts
//                        (TDT_AGG_DIRECT=0: no direct kernel, every candidate goes through the three sorts)
//   agg_finalize         the branch logic of :265-330, one 16-int row per candidate
#include "tdt_common.cuh"
#include "tdt_segsort.cuh"

namespace tdt {

constexpr int AG_THREADS = 256;
#ifndef TDT_AG_ITEMS
#define TDT_AG_ITEMS 8
#endif
constexpr int AG_ITEMS = TDT_AG_ITEMS;   // signals per thread of the three chained scans (a multiple of 4 that divides 32).
                                         // B200, 30X set, whole call: 8 -> 1.655 ms, 16 -> 1.640 (within noise), 32 -> 1.764
static_assert(AG_ITEMS % 4 == 0 && 32 % AG_ITEMS == 0, "AG_ITEMS: a multiple of 4 that divides 32");
constexpr int AG_TILE = AG_THREADS * AG_ITEMS;
constexpr int AG_ACC = 12;    // int32 accumulators per candidate
constexpr int AG_MODES = 6;   // u64 (count:~first) per side x kind
enum { AG_ERR_LABEL = 8, AG_ERR_POS = 9, AG_ERR_NAME = 10 };
constexpr int AG_DIRECT_RB = 10;                       // bits of the in-candidate position in a packed (count : ~position) word
constexpr int AG_DIRECT_LIMIT = 1 << AG_DIRECT_RB;     // largest candidate the direct kernel can be given
#ifndef TDT_AGG_DIRECT_DEFAULT
#define TDT_AGG_DIRECT_DEFAULT 256   // O(k) work per member.  B200, whole call (tools/agg_direct_ab.py), 30X set / tumour set:
                                     // 0 (all sorts) 2.23 / 3.60 ms, 32: 2.03 / 3.79, 64: 1.80 / 3.44, 128: 1.79 / 3.00,
                                     // 256: 1.79 / 2.82, 1024: 1.79 / 2.87
#endif

struct AggDims {
    int64_t n, nseg;
};

struct AggSmall {                // one 256-byte aligned record of device-side scalars
    AggDims d1;                  // {kept signals, P}
    AggDims d2;                  // {members kept, candidates}
    u32 ticket[4];
    int err;
    int pad[3];
    AggDims d3;                  // {members of the big candidates, big candidates}: what the sub-sorts see in direct mode
    u64 big_ctr;                 // (big candidates << 32) | their members: ONE atomic reserves a slot and a member range
    u64 pad2;
};

struct AggParams {
    // inputs
    const int32_t *labels, *posA, *posB, *name_id;
    const int4 *span;
    const uint8_t *flags, *same_chrom;
    const int64_t *seg_off;
    int64_t n;
    int32_t P, max_ins_len, is_mp, min_reads;
    int pos_bits, name_bits;
    u32 sentinel;
    // scratch
    AggSmall *small;
    u32 *pair_base;      // [P]   first survivor group of the pair
    u32 *pair_heads;     // bit j: kept signal j is the first of its pair
    u32 *key1, *key1s;   // [M]   candidate keys of the kept signals, insertion / sorted order
    int32_t *val1;       // [M]   insertion index of the kept signals (sorted order: member_idx)
    int32_t *val1s;      // [M]   sub-sort values
    int64_t *coff;       // [P+1] pair offsets into the kept signals
    int32_t *gfirst;     // [n]   (group + 1) at the insertion index of a group's first member, else 0
    int32_t *c_grp;      // [M]   group of every kept signal (sorted order)
    int64_t *goff;       // [G+1] group offsets into the compact order
    int32_t *gpair, *gcid, *slot;   // [G]
    u32 *keyA, *keyB, *keyN;        // [M] sub-sort keys
    int32_t *acc;        // [G][AG_ACC]
    u64 *modes;          // [G][AG_MODES]
    u32 *ncnt;           // [G][4]
    u64 *status1, *status2, *status3;
    // direct mode: candidates beyond direct_max members ("big"), compacted for the sub-sorts
    int direct_max;      // 0: no direct kernel
    int2 *big;           // [G]   {first position in the big arrays, big slot}; written for big candidates only
    int64_t *goff_big;   // [B+1] offsets of the big candidates into the big arrays
    int32_t *bigg;       // [B]   candidate of a big slot
    u32 *bkA, *bkB, *bkN;           // [M] sub-sort keys of the big candidates' members
    int32_t *bgrp, *borig;          // [M] big slot / compact position of the same
    // outputs
    int32_t *cand_out, *member_idx;
    int64_t *counts_out;
};

struct ScanSmem {
    int tile;
    u32 warpA[AG_THREADS / 32], warpB[AG_THREADS / 32];
    u32 exA, exB;
};

__device__ __forceinline__ int take_tile(ScanSmem &s, u32 *ticket) {
    if (threadIdx.x == 0) s.tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    return s.tile;
}

// (sumA, sumB): the thread's own totals.  Returns the exclusive prefixes over all earlier threads of all earlier
// tiles in (preA, preB).  All threads of the CTA must call.
__device__ __forceinline__ void block_scan2(ScanSmem &s, u64 *status, int tile, u32 sumA, u32 sumB, u32 &preA, u32 &preB) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    u32 incA = sumA, incB = sumB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, incA, o);
        const u32 b = __shfl_up_sync(0xffffffffu, incB, o);
        if (lane >= o) {
            incA += a;
            incB += b;
        }
    }
    if (lane == 31) {
        s.warpA[warp] = incA;
        s.warpB[warp] = incB;
    }
    __syncthreads();
    if (warp == 0) {
        const u32 wa = lane < AG_THREADS / 32 ? s.warpA[lane] : 0u;
        const u32 wb = lane < AG_THREADS / 32 ? s.warpB[lane] : 0u;
        u32 ia = wa, ib = wb;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, ia, o);
            const u32 b = __shfl_up_sync(0xffffffffu, ib, o);
            if (lane >= o) {
                ia += a;
                ib += b;
            }
        }
        const u32 aggA = __shfl_sync(0xffffffffu, ia, 31), aggB = __shfl_sync(0xffffffffu, ib, 31);
        if (lane < AG_THREADS / 32) {
            s.warpA[lane] = ia - wa;
            s.warpB[lane] = ib - wb;
        }
        u32 exA, exB;
        lookback(status, tile, aggA, aggB, exA, exB);
        if (lane == 0) {
            s.exA = exA;
            s.exB = exB;
        }
    }
    __syncthreads();
    preA = s.exA + s.warpA[warp] + (incA - sumA);
    preB = s.exB + s.warpB[warp] + (incB - sumB);
}

// largest p with seg_off[p] <= i (i < seg_off[P]): the non-empty pair that owns element i
__device__ __forceinline__ int find_pair(const int64_t *__restrict__ off, int P, int64_t i) {
    int lo = 0, hi = P;  // invariant: off[lo] <= i < off[hi]
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid;
    }
    return lo;
}

// ---- candidate keys + compaction (scan 1, insertion order).  Noise is dropped here, before the sort; surviving noise
// contigs (tiddit_cluster.pyx:163-168) share one key below the sentinel: the stable sort keeps them in insertion
// order behind the pair's clusters, each becomes a group of its own, and its id len(pair) + k follows from its rank
// among the pair's survivor groups. -------------------------------------------------------------------------------------
__device__ __forceinline__ u32 candidate_key(const AggParams &a, int64_t i, int p, int64_t np) {
    const int32_t lab = a.labels[i];
    if (lab >= 0) {
        if (lab >= np) {   // DBSCAN ids are < len(pair)
            atomicMax(&a.small->err, (int)AG_ERR_LABEL);
            return a.sentinel;
        }
        return (u32)lab;
    }
    if (lab != -1) {
        atomicMax(&a.small->err, (int)AG_ERR_LABEL);
        return a.sentinel;
    }
    if ((a.flags[i] & 3) == 2 && a.same_chrom[p] &&
        ((int64_t)a.posB[i] - (int64_t)a.posA[i]) < 2 * (int64_t)a.max_ins_len)
        return a.sentinel - 1u;
    return a.sentinel;
}

__global__ void __launch_bounds__(AG_THREADS) agg_key_scan_kernel(AggParams a) {
    __shared__ ScanSmem s;
    __shared__ int s_p0;
    const int tile = take_tile(s, &a.small->ticket[0]);
    const int64_t t0 = (int64_t)tile * AG_TILE;
    if (threadIdx.x == 0) s_p0 = find_pair(a.seg_off, a.P, t0);
    __syncthreads();
    const int64_t i0 = t0 + (int64_t)threadIdx.x * AG_ITEMS;
    u32 keys[AG_ITEMS];
    u32 keptm = 0;
    int p = s_p0;
    if (i0 < a.n) {
        // the pair of the thread's first signal: a short walk from the tile's pair, a bisection when pairs are tiny
        int steps = 0;
        while (i0 >= a.seg_off[p + 1] && steps < 8) {
            p++;
            steps++;
        }
        if (i0 >= a.seg_off[p + 1]) p = find_pair(a.seg_off, a.P, i0);
        int q = p;
        int64_t qend = a.seg_off[q + 1], qn = qend - a.seg_off[q];
#pragma unroll
        for (int k = 0; k < AG_ITEMS; k++) {
            const int64_t i = i0 + k;
            if (i < a.n) {
                while (i >= qend) {
                    q++;
                    qend = a.seg_off[q + 1];
                    qn = qend - a.seg_off[q];
                }
                keys[k] = candidate_key(a, i, q, qn);
                if (keys[k] != a.sentinel) keptm |= 1u << k;
            }
        }
    }
    u32 pre, unused;
    block_scan2(s, a.status1, tile, (u32)__popc(keptm), 0u, pre, unused);
    if (i0 < a.n) {
        int q = p;
#pragma unroll
        for (int k = 0; k < AG_ITEMS; k++) {
            const int64_t i = i0 + k;
            if (i < a.n) {
                while (i >= a.seg_off[q + 1]) q++;
                if (i == a.seg_off[q]) {   // first signal of pair q (and of the empty pairs right before it)
                    for (int e = q; e >= 0 && a.seg_off[e] == i; e--) a.coff[e] = (int64_t)pre;
                }
                if (keptm & (1u << k)) {
                    a.key1[pre] = keys[k];
                    a.val1[pre] = (int32_t)i;
                    pre++;
                }
                if (i == a.n - 1) {   // totals; trailing empty pairs
                    for (int e = a.P; e > q; e--) a.coff[e] = (int64_t)pre;
                    a.small->d1.n = (int64_t)pre;
                    a.small->d1.nseg = a.P;
                    a.small->d2.n = (int64_t)pre;
                    a.counts_out[1] = (int64_t)pre;
                }
            }
        }
    }
}

__global__ void agg_heads_kernel(AggParams a) {
    const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s < a.P) {
        const int64_t q = a.coff[s];
        if (a.coff[s + 1] > q) atomicOr(a.pair_heads + (q >> 5), 1u << (q & 31));
    }
}

// ---- scan 2: group heads over the (pair, key)-sorted kept signals ------------------------------------------------------
__global__ void __launch_bounds__(AG_THREADS) agg_mark_scan_kernel(AggParams a) {
    __shared__ ScanSmem s;
    const int tile = take_tile(s, &a.small->ticket[1]);
    const int64_t M = a.small->d2.n;
    const int64_t j0 = (int64_t)tile * AG_TILE + (int64_t)threadIdx.x * AG_ITEMS;
    if ((int64_t)tile * AG_TILE >= M) {
        if (M == 0 && tile == 0 && threadIdx.x == 0) a.goff[0] = 0;
        return;   // nobody waits on a tile behind the last signal
    }
    u32 keys[AG_ITEMS];
    u32 headm = 0, firstsurv = 0;
    if (j0 < M) {
        u32 prev = j0 > 0 ? a.key1s[j0 - 1] : 0u;
        u32 hw = a.pair_heads[j0 >> 5] >> (j0 & 31);   // j0 is a multiple of AG_ITEMS, which divides 32: the bits sit in one word
#pragma unroll
        for (int k = 0; k < AG_ITEMS; k++) {
            const int64_t j = j0 + k;
            if (j < M) {
                const u32 key = a.key1s[j];
                keys[k] = key;
                const bool ph = (hw >> k) & 1u;
                if (ph || key != prev || key == a.sentinel - 1u) headm |= 1u << k;
                if (key == a.sentinel - 1u && (ph || key != prev)) firstsurv |= 1u << k;
                prev = key;
            }
        }
    }
    u32 preH, unused;
    block_scan2(s, a.status2, tile, (u32)__popc(headm), 0u, preH, unused);
    if (j0 < M) {
        int p = -1;
#pragma unroll
        for (int k = 0; k < AG_ITEMS; k++) {
            const int64_t j = j0 + k;
            if (j < M) {
                if (headm & (1u << k)) {
                    if (p < 0) p = find_pair(a.coff, a.P, j);
                    while (j >= a.coff[p + 1]) p++;
                    a.goff[preH] = j;
                    a.gpair[preH] = p;
                    a.gcid[preH] = (int32_t)keys[k];
                    if (firstsurv & (1u << k)) a.pair_base[p] = preH;   // first survivor group of the pair
                    a.gfirst[a.member_idx[j]] = (int32_t)preH + 1;
                    preH++;
                }
                a.c_grp[j] = (int32_t)preH - 1;
                if (j == M - 1) {   // totals
                    a.goff[preH] = M;
                    a.small->d2.nseg = (int64_t)preH;
                    a.counts_out[0] = (int64_t)preH;
                }
            }
        }
    }
}

// ---- scan 3: candidates ranked by first appearance (dict insertion order) ---------------------------------------------
__global__ void __launch_bounds__(AG_THREADS) agg_rank_scan_kernel(AggParams a) {
    __shared__ ScanSmem s;
    const int tile = take_tile(s, &a.small->ticket[2]);
    const int64_t i0 = (int64_t)tile * AG_TILE + (int64_t)threadIdx.x * AG_ITEMS;
    int32_t g[AG_ITEMS];
    u32 cnt = 0;
    if (i0 + AG_ITEMS <= a.n) {
#pragma unroll
        for (int k = 0; k < AG_ITEMS; k += 4) {
            const int4 v = *(const int4 *)(a.gfirst + i0 + k);
            g[k] = v.x; g[k + 1] = v.y; g[k + 2] = v.z; g[k + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < AG_ITEMS; k++) g[k] = i0 + k < a.n ? a.gfirst[i0 + k] : 0;
    }
#pragma unroll
    for (int k = 0; k < AG_ITEMS; k++) cnt += g[k] != 0;
    u32 pre, unused;
    block_scan2(s, a.status3, tile, cnt, 0u, pre, unused);
#pragma unroll
    for (int k = 0; k < AG_ITEMS; k++)
        if (g[k]) a.slot[g[k] - 1] = (int32_t)pre++;
}

// ---- accumulators ------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) agg_init_kernel(AggParams a) {
    const int64_t G = a.small->d2.nseg;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int64_t t = t0; t < G * AG_ACC; t += stride) {
        const int k = (int)(t % AG_ACC);
        a.acc[t] = k >= 8 ? 0 : ((k & 1) ? (int32_t)0x80000000 : 0x7fffffff);   // even: min, odd: max, then sums
    }
    for (int64_t t = t0; t < G * AG_MODES; t += stride) a.modes[t] = 0ull;
    for (int64_t t = t0; t < G; t += stride) ((uint4 *)a.ncnt)[t] = make_uint4(0u, 0u, 0u, 0u);
    if (a.direct_max > 0) {
        // big candidates reserve a slot and a member range in the big arrays (any order: the sub-sorts only need the
        // ranges to tile [0, members)); slot s + 1 starts where slot s ends, so both writers of goff_big[s + 1] agree
        for (int64_t t = t0; t < G; t += stride) {
            const int64_t k = a.goff[t + 1] - a.goff[t];
            if (k > a.direct_max) {
                const u64 old = atomicAdd(&a.small->big_ctr, (1ull << 32) | (u64)k);
                const int32_t slot = (int32_t)(old >> 32);
                const int64_t start = (int64_t)(old & 0xffffffffull);
                a.big[t] = make_int2((int32_t)start, slot);
                a.goff_big[slot] = start;
                a.goff_big[slot + 1] = start + k;
                a.bigg[slot] = (int32_t)t;
            }
        }
    }
}

__device__ __forceinline__ int32_t seg_min(int32_t v, int32_t g, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t w = __shfl_up_sync(0xffffffffu, v, o);
        const int32_t h = __shfl_up_sync(0xffffffffu, g, o);
        if (lane >= o && h == g) v = min(v, w);
    }
    return v;
}
__device__ __forceinline__ int32_t seg_max(int32_t v, int32_t g, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t w = __shfl_up_sync(0xffffffffu, v, o);
        const int32_t h = __shfl_up_sync(0xffffffffu, g, o);
        if (lane >= o && h == g) v = max(v, w);
    }
    return v;
}
__device__ __forceinline__ u32 seg_add(u32 v, int32_t g, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 w = __shfl_up_sync(0xffffffffu, v, o);
        const int32_t h = __shfl_up_sync(0xffffffffu, g, o);
        if (lane >= o && h == g) v += w;
    }
    return v;
}

__global__ void __launch_bounds__(256) agg_gather_kernel(AggParams a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t M = a.small->d2.n;
    const int lane = threadIdx.x & 31;
    if (c == 0) {   // agg_init_kernel has finished: the totals of the big candidates become the sub-sorts' dims
        const u64 ctr = a.small->big_ctr;
        a.small->d3.n = (int64_t)(ctr & 0xffffffffull);
        a.small->d3.nseg = (int64_t)(ctr >> 32);
    }
    if (c - lane >= M) return;   // whole warp beyond the end
    const bool live = c < M;
    int32_t g = -1;
    int32_t sA = 0x7fffffff, eA = (int32_t)0x80000000, sB = 0x7fffffff, eB = (int32_t)0x80000000;
    int32_t dAmin = 0x7fffffff, dAmax = (int32_t)0x80000000, dBmin = 0x7fffffff, dBmax = (int32_t)0x80000000;
    u32 ori = 0;   // four 8-bit counters: revA, fwdA, revB, fwdB (a warp adds at most 32 to each)
    if (live) {
        const int32_t idx = a.member_idx[c];
        g = a.c_grp[c];
        const u32 f = a.flags[idx];
        const u32 kind = f & 3u;
        const int32_t x = a.posA[idx], y = a.posB[idx], nm = a.name_id[idx];
        const int4 sp = a.span[idx];
        if (kind == 3u) atomicMax(&a.small->err, (int)AG_ERR_LABEL);
        if (x < 0 || y < 0 || (a.pos_bits < 31 && ((x >> a.pos_bits) || (y >> a.pos_bits))))
            atomicMax(&a.small->err, (int)AG_ERR_POS);
        if (nm < 0 || (nm >> a.name_bits)) atomicMax(&a.small->err, (int)AG_ERR_NAME);
        const u32 kA = (kind << a.pos_bits) | (u32)x, kB = (kind << a.pos_bits) | (u32)y, kN = (kind << a.name_bits) | (u32)nm;
        a.keyA[c] = kA;
        a.keyB[c] = kB;
        a.keyN[c] = kN;
        if (a.direct_max > 0) {   // members of a big candidate: a second copy, compacted for the sub-sorts
            const int64_t lo = a.goff[g];
            if (a.goff[g + 1] - lo > (int64_t)a.direct_max) {
                const int2 b = a.big[g];
                const int64_t j = (int64_t)b.x + (c - lo);
                a.bkA[j] = kA;
                a.bkB[j] = kB;
                a.bkN[j] = kN;
                a.bgrp[j] = b.y;
                a.borig[j] = (int32_t)c;
            }
        }
        sA = sp.x; eA = sp.y; sB = sp.z; eB = sp.w;
        if (kind == 0u) {
            dAmin = dAmax = x;
            dBmin = dBmax = y;
            ori = ((f >> 2) & 1u) | (((f >> 3) & 1u) << 8) | (((f >> 4) & 1u) << 16) | (((f >> 5) & 1u) << 24);
        }
    }
    sA = seg_min(sA, g, lane); eA = seg_max(eA, g, lane);
    sB = seg_min(sB, g, lane); eB = seg_max(eB, g, lane);
    dAmin = seg_min(dAmin, g, lane); dAmax = seg_max(dAmax, g, lane);
    dBmin = seg_min(dBmin, g, lane); dBmax = seg_max(dBmax, g, lane);
    ori = seg_add(ori, g, lane);
    const int32_t gnext = __shfl_down_sync(0xffffffffu, g, 1);
    if (live && (lane == 31 || gnext != g)) {   // last lane of the candidate's run inside this warp
        int32_t *acc = a.acc + (int64_t)g * AG_ACC;
        atomicMin(acc + 0, sA); atomicMax(acc + 1, eA);
        atomicMin(acc + 2, sB); atomicMax(acc + 3, eB);
        if (dAmin <= dAmax) {
            atomicMin(acc + 4, dAmin); atomicMax(acc + 5, dAmax);
            atomicMin(acc + 6, dBmin); atomicMax(acc + 7, dBmax);
            if (ori & 0xffu) atomicAdd(acc + 8, (int)(ori & 0xffu));
            if ((ori >> 8) & 0xffu) atomicAdd(acc + 9, (int)((ori >> 8) & 0xffu));
            if ((ori >> 16) & 0xffu) atomicAdd(acc + 10, (int)((ori >> 16) & 0xffu));
            if (ori >> 24) atomicAdd(acc + 11, (int)(ori >> 24));
        }
    }
}

// ---- direct modes / distinct names of the candidates with <= direct_max members -------------------------------------------
// One thread per member (compact order: a candidate's members are neighbours, in insertion order).  The member walks its
// candidate's keys (neighbouring lanes read the same words: broadcasts out of L1) and counts the members with its own
// (kind, posA) and (kind, posB) and whether an EARLIER member carries its (kind, name).  What the sort + run kernels
// produce follows without a sort: the mode is the maximum over the members of (count : ~position) -- among the members
// of the most frequent value the first-inserted one wins, exactly the run head the stable sort would present -- and
// the distinct names are the members without an earlier equal.  Both are reduced over the warp's lanes of the same
// candidate first (one hoisted set of segment predicates, 32-bit packed words), then one atomic per (candidate, warp,
// kind).  The members of big candidates are left to the sorts.
__device__ __forceinline__ u64 direct_mode_word(u32 v, int64_t lo) {
    const u64 count = (u64)(v >> AG_DIRECT_RB);
    const u32 first = (u32)lo + ((u32)(AG_DIRECT_LIMIT - 1) - (v & (u32)(AG_DIRECT_LIMIT - 1)));
    return (count << 32) | (u64)(0xffffffffu - first);
}

__global__ void __launch_bounds__(256) agg_direct_kernel(AggParams a) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t M = a.small->d2.n;
    const int lane = threadIdx.x & 31;
    if (c - lane >= M) return;   // whole warp beyond the end
    const bool live = c < M;
    int32_t g = -1;
    int64_t lo = 0;
    u32 vA0 = 0, vA1 = 0, vA2 = 0, vB0 = 0, vB1 = 0, vB2 = 0, one = 0;
    if (live) {
        g = a.c_grp[c];
        lo = a.goff[g];
        const int64_t hi = a.goff[g + 1];
        if (hi - lo <= (int64_t)a.direct_max) {   // (a big candidate's members were compacted for the sorts by agg_gather_kernel)
            const u32 kA = a.keyA[c], kB = a.keyB[c], kN = a.keyN[c];
            const int jl = (int)lo, jh = (int)hi, jc = (int)c;
            const u32 *__restrict__ pA = a.keyA, *__restrict__ pB = a.keyB, *__restrict__ pN = a.keyN;
            u32 cntA = 0, cntB = 0, dup = 0;
#pragma unroll 4
            for (int j = jl; j < jh; j++) {
                cntA += __ldg(pA + j) == kA ? 1u : 0u;
                cntB += __ldg(pB + j) == kB ? 1u : 0u;
                dup |= (__ldg(pN + j) == kN && j < jc) ? 1u : 0u;
            }
            const u32 kind = (kA >> a.pos_bits) & 3u;   // 3 is a data error flagged by agg_gather_kernel
            const u32 rel = (u32)(AG_DIRECT_LIMIT - 1 - (jc - jl));
            const u32 wa = (cntA << AG_DIRECT_RB) | rel, wb = (cntB << AG_DIRECT_RB) | rel;
            vA0 = kind == 0u ? wa : 0u; vA1 = kind == 1u ? wa : 0u; vA2 = kind == 2u ? wa : 0u;
            vB0 = kind == 0u ? wb : 0u; vB1 = kind == 1u ? wb : 0u; vB2 = kind == 2u ? wb : 0u;
            one = dup ? 0u : (1u << (8u * kind));
        }
    }
    // inclusive segmented scans over the lanes of the same candidate (max of the six mode words, sum of the name counters:
    // four 8-bit fields, a warp adds at most 32 to each)
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int32_t h = __shfl_up_sync(0xffffffffu, g, o);
        const bool same = lane >= o && h == g;
        const u32 a0 = __shfl_up_sync(0xffffffffu, vA0, o), a1 = __shfl_up_sync(0xffffffffu, vA1, o);
        const u32 a2 = __shfl_up_sync(0xffffffffu, vA2, o), b0 = __shfl_up_sync(0xffffffffu, vB0, o);
        const u32 b1 = __shfl_up_sync(0xffffffffu, vB1, o), b2 = __shfl_up_sync(0xffffffffu, vB2, o);
        const u32 w = __shfl_up_sync(0xffffffffu, one, o);
        if (same) {
            vA0 = max(vA0, a0); vA1 = max(vA1, a1); vA2 = max(vA2, a2);
            vB0 = max(vB0, b0); vB1 = max(vB1, b1); vB2 = max(vB2, b2);
            one += w;
        }
    }
    const int32_t gnext = __shfl_down_sync(0xffffffffu, g, 1);
    if (live && (lane == 31 || gnext != g)) {   // last lane of the candidate's run inside this warp
        u64 *md = a.modes + (int64_t)g * AG_MODES;
        if (vA0) atomicMax(md + 0, direct_mode_word(vA0, lo));
        if (vA1) atomicMax(md + 1, direct_mode_word(vA1, lo));
        if (vA2) atomicMax(md + 2, direct_mode_word(vA2, lo));
        if (vB0) atomicMax(md + 3, direct_mode_word(vB0, lo));
        if (vB1) atomicMax(md + 4, direct_mode_word(vB1, lo));
        if (vB2) atomicMax(md + 5, direct_mode_word(vB2, lo));
#pragma unroll
        for (int t = 0; t < 4; t++) {
            const u32 f = (one >> (8 * t)) & 0xffu;
            if (f) atomicAdd(a.ncnt + (int64_t)g * 4 + t, f);
        }
    }
}

// ---- runs of equal keys inside every candidate (after the sub-sort) -----------------------------------------------------
// WHAT 0 / 1: mode of (kind, posA) / (kind, posB); 2: distinct (kind, name).  BIG: the sorted arrays hold the members of
// the big candidates only (direct mode): segment = big slot, positions translated back through bigg / borig.
// One element of the sorted arrays per thread; called by whole warps (c - lane < M).
template <int WHAT, bool BIG>
__device__ __forceinline__ void agg_runs_element(const AggParams &a, const u32 *__restrict__ keys,
                                                 const int32_t *__restrict__ vals, int64_t c, int64_t M, int lane) {
    const bool live = c < M;
    const int64_t *__restrict__ goff = BIG ? a.goff_big : a.goff;
    int32_t g = 0, gr = 0;       // segment of the sorted arrays / the candidate it is
    int64_t lo = 0;
    u32 key = 0;
    bool head = false;
    if (live) {
        g = BIG ? a.bgrp[c] : a.c_grp[c];
        gr = BIG ? a.bigg[g] : g;
        lo = goff[g];
        key = keys[c];
        head = c == lo || keys[c - 1] != key;   // first element of a run of equal keys
    }
    if (WHAT == 2) {
        // distinct names per (candidate, kind): run heads counted per warp, one atomic per (candidate, kind, warp)
        const u32 kind = key >> a.name_bits;
        const u32 ck = live ? (((u32)g << 2) | kind) : (0xffffffffu - (u32)lane);
        const u32 prevk = __shfl_up_sync(0xffffffffu, ck, 1);
        const u32 nextk = __shfl_down_sync(0xffffffffu, ck, 1);
        const u32 bnd = __ballot_sync(0xffffffffu, lane == 0 || prevk != ck);
        const u32 heads = __ballot_sync(0xffffffffu, head);
        if (live && (lane == 31 || nextk != ck)) {
            const int start = 31 - __clz(bnd & lanemask_le());
            const u32 cnt = __popc(heads & lanemask_le() & ~((1u << start) - 1u));
            if (cnt) atomicAdd(a.ncnt + (int64_t)gr * 4 + kind, cnt);
        }
        return;
    }
    if (!head) return;
    // run end: gallop, then bisect (keys ascend inside the candidate)
    const int64_t hi = goff[g + 1];
    int64_t left = c, step = 1;            // keys[left] == key
    int64_t right = hi;                    // keys[right] > key (or right == hi)
    while (left + step < hi) {
        if (keys[left + step] == key) {
            left += step;
            step <<= 1;
        } else {
            right = left + step;
            break;
        }
    }
    while (right - left > 1) {
        const int64_t mid = (left + right) >> 1;
        if (keys[mid] == key) left = mid; else right = mid;
    }
    const u64 count = (u64)(right - c);
    // compact position of the first occurrence (the sort is stable)
    const u32 first = BIG ? (u32)a.borig[vals[c]] : (u32)vals[c];
    const u32 kind = key >> a.pos_bits;
    atomicMax(a.modes + (int64_t)gr * AG_MODES + WHAT * 3 + kind, (count << 32) | (u64)(0xffffffffu - first));
}

// Grid-stride over the sorted elements: the BIG launches cover "every member is a big one" with a capped grid (a grid
// of n / 256 CTAs that only find out that nothing is big would be dispatched ahead of the direct kernel's CTAs).
template <int WHAT, bool BIG>
__global__ void __launch_bounds__(256) agg_runs_kernel(AggParams a, const u32 *__restrict__ keys,
                                                       const int32_t *__restrict__ vals) {
    const int64_t M = BIG ? a.small->d3.n : a.small->d2.n;
    const int lane = threadIdx.x & 31;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c - lane < M; c += stride)
        agg_runs_element<WHAT, BIG>(a, keys, vals, c, M, lane);
}

// ---- one row per candidate (tiddit_cluster.pyx:258-336) ------------------------------------------------------------------
__global__ void __launch_bounds__(256) agg_finalize_kernel(AggParams a) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t G = a.small->d2.nseg;
    if (g >= G) return;
    const int32_t *acc = a.acc + g * AG_ACC;
    const u64 *md = a.modes + g * AG_MODES;
    const uint4 nc = *(const uint4 *)(a.ncnt + g * 4);
    const int64_t lo = a.goff[g], hi = a.goff[g + 1];
    int32_t row[16];
    const int p = a.gpair[g];
    row[0] = p;
    row[1] = (u32)a.gcid[g] == a.sentinel - 1u
                 ? (int32_t)(a.seg_off[p + 1] - a.seg_off[p] + (g - (int64_t)a.pair_base[p]))   // len(pair) + k (:166)
                 : a.gcid[g];
    row[2] = a.member_idx[lo];
    row[3] = (int32_t)lo;
    row[4] = (int32_t)(hi - lo);
    row[5] = (int32_t)nc.x;
    row[6] = (int32_t)nc.y;
    row[7] = (int32_t)nc.z;
    int use_kind = -1, rule;
    if (nc.y && (int64_t)a.min_reads <= (int64_t)nc.y) { use_kind = 1; rule = 0; }
    else if (nc.z) { use_kind = 2; rule = 1; }
    else if (nc.y) { use_kind = 1; rule = 2; }
    else {
        const int64_t revA = acc[8], fwdA = acc[9], revB = acc[10], fwdB = acc[11];
        if ((revA >= 5 * fwdA || revA * 5 <= fwdA) && (revB >= 5 * fwdB || revB * 5 <= fwdB)) {
            const bool A_rev = revA > fwdA, B_rev = revB > fwdB;
            const bool maxA = a.is_mp ? A_rev : !A_rev, maxB = a.is_mp ? B_rev : !B_rev;
            row[8] = maxA ? acc[5] : acc[4];
            row[9] = maxB ? acc[7] : acc[6];
            rule = 3;
        } else {
            use_kind = 0;
            rule = 4;
        }
    }
    if (use_kind >= 0) {
        const u32 fa = 0xffffffffu - (u32)md[use_kind];
        const u32 fb = 0xffffffffu - (u32)md[3 + use_kind];
        row[8] = a.posA[a.member_idx[fa]];
        row[9] = a.posB[a.member_idx[fb]];
    }
    row[10] = acc[0]; row[11] = acc[1]; row[12] = acc[2]; row[13] = acc[3];
    row[14] = rule;
    row[15] = 0;
    int4 *out = (int4 *)(a.cand_out + (int64_t)a.slot[g] * 16);
    out[0] = make_int4(row[0], row[1], row[2], row[3]);
    out[1] = make_int4(row[4], row[5], row[6], row[7]);
    out[2] = make_int4(row[8], row[9], row[10], row[11]);
    out[3] = make_int4(row[12], row[13], row[14], row[15]);
}

__global__ void agg_status_kernel(AggParams a) {
    a.counts_out[2] = (int64_t)a.small->err;
}

// ---- workspace -----------------------------------------------------------------------------------------------------------
struct AggPlan {
    size_t sort, total;
    int64_t tiles;
};

static AggPlan agg_plan(int64_t n, int32_t P) {
    AggPlan pl;
    const int64_t nseg = n > P ? n : P;
    pl.sort = (segsort_temp_bytes(n, nseg) + 255) & ~(size_t)255;
    pl.tiles = (n + AG_TILE - 1) / AG_TILE + 1;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    size_t t = 0;
    t += al(sizeof(AggSmall));
    t += 3 * al((size_t)pl.tiles * 8);                 // scan status
    t += al((size_t)(P + 1) * 4);                      // pair_base
    t += al((size_t)(P + 2) * 8);                      // coff
    t += al((size_t)(n + 4) * 4);                      // val1
    t += al((size_t)(n / 32 + 2) * 4);                 // pair_heads
    t += 11 * al((size_t)(n + 4) * 4);                 // key1 key1s val1s tmpK tmpV c_grp keyA keyB keyN gfirst(slot alias no) gpair
    t += 2 * al((size_t)(n + 4) * 4);                  // gcid slot
    t += al((size_t)(n + 2) * 8);                      // goff
    t += al((size_t)n * AG_ACC * 4) + al((size_t)n * AG_MODES * 8) + al((size_t)n * 16);
    t += 3 * (pl.sort + 256);                          // one sort scratch per parallel sub-sort branch
    t += 2 * 4 * al((size_t)(n + 4) * 4);              // key1s / val1s / tmpK / tmpV of branches 1 and 2
    t += al((size_t)(n + 4) * 8) + al((size_t)(n + 2) * 8) + 6 * al((size_t)(n + 4) * 4);   // big goff_big | bigg bkA bkB bkN bgrp borig
    pl.total = t;
    return pl;
}

static int aggregate_impl(AggParams a, void *ws, size_t ws_bytes, cudaStream_t st) {
    const int64_t n = a.n;
    const int32_t P = a.P;
    const AggPlan pl = agg_plan(n, P);
    if (ws == nullptr || ws_bytes < pl.total)
        return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);
    Arena ar(ws, ws_bytes);
    a.small = ar.take<AggSmall>(1);
    a.status1 = ar.take<u64>(pl.tiles);
    a.status2 = ar.take<u64>(pl.tiles);
    a.status3 = ar.take<u64>(pl.tiles);
    a.pair_heads = ar.take<u32>(n / 32 + 2);
    a.gfirst = ar.take<int32_t>(n + 4);
    char *zero_end = ar.base + ar.off;               // everything up to here is zeroed by one memset
    a.pair_base = ar.take<u32>(P + 1);
    a.coff = ar.take<int64_t>(P + 2);
    a.val1 = ar.take<int32_t>(n + 4);
    a.key1 = ar.take<u32>(n + 4);
    a.key1s = ar.take<u32>(n + 4);
    a.val1s = ar.take<int32_t>(n + 4);
    u32 *tmpK = ar.take<u32>(n + 4);
    int32_t *tmpV = ar.take<int32_t>(n + 4);
    a.c_grp = ar.take<int32_t>(n + 4);
    a.keyA = ar.take<u32>(n + 4);
    a.keyB = ar.take<u32>(n + 4);
    a.keyN = ar.take<u32>(n + 4);
    a.gpair = ar.take<int32_t>(n + 4);
    a.gcid = ar.take<int32_t>(n + 4);
    a.slot = ar.take<int32_t>(n + 4);
    a.goff = ar.take<int64_t>(n + 2);
    a.acc = ar.take<int32_t>((size_t)n * AG_ACC);
    a.modes = ar.take<u64>((size_t)n * AG_MODES);
    a.ncnt = ar.take<u32>((size_t)n * 4);
    a.big = ar.take<int2>(n + 4);
    a.goff_big = ar.take<int64_t>(n + 2);
    a.bigg = ar.take<int32_t>(n + 4);
    a.bkA = ar.take<u32>(n + 4);
    a.bkB = ar.take<u32>(n + 4);
    a.bkN = ar.take<u32>(n + 4);
    a.bgrp = ar.take<int32_t>(n + 4);
    a.borig = ar.take<int32_t>(n + 4);
    void *sort_temp = ar.take<char>(pl.sort);
    // the three sub-sorts (by posA, posB, name inside every candidate) are independent: branches 1 and 2 get their own
    // outputs, ping-pong buffers and sort scratch and run next to branch 0
    u32 *b_keys[3] = {a.key1s, ar.take<u32>(n + 4), ar.take<u32>(n + 4)};
    int32_t *b_vals[3] = {a.val1s, ar.take<int32_t>(n + 4), ar.take<int32_t>(n + 4)};
    u32 *b_tmpK[3] = {tmpK, ar.take<u32>(n + 4), ar.take<u32>(n + 4)};
    int32_t *b_tmpV[3] = {tmpV, ar.take<int32_t>(n + 4), ar.take<int32_t>(n + 4)};
    void *b_temp[3] = {sort_temp, ar.take<char>(pl.sort), ar.take<char>(pl.sort)};
    if (!sort_temp || !b_temp[2]) return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);

    TDT_CUDA(cudaMemsetAsync(ws, 0, (size_t)(zero_end - (char *)ws), st));
    TDT_CUDA(cudaMemsetAsync(a.counts_out, 0, 4 * sizeof(int64_t), st));
    const unsigned tiles = (unsigned)((n + AG_TILE - 1) / AG_TILE);
    const unsigned per_elem = (unsigned)((n + 255) / 256);
    {
        ProfScope ps("agg_keys", st);
        TDT_LAUNCH(agg_key_scan_kernel, tiles, AG_THREADS, 0, st, a);
        TDT_LAUNCH(agg_heads_kernel, (unsigned)((P + 255) / 256), 256, 0, st, a);
    }
    {
        ProfScope ps("agg_sort_groups", st);
        const int key_bits = bit_width_u32(a.sentinel);
        int rc = segsort_pairs(a.key1, a.val1, a.key1s, a.member_idx, tmpK, tmpV, a.coff, (const int64_t *)&a.small->d1,
                               nullptr, n, P, key_bits, sort_temp, pl.sort, &a.small->err, st);
        if (rc) return rc;
    }
    {
        ProfScope ps("agg_groups", st);
        TDT_LAUNCH(agg_mark_scan_kernel, tiles, AG_THREADS, 0, st, a);
        TDT_LAUNCH(agg_rank_scan_kernel, tiles, AG_THREADS, 0, st, a);
        TDT_LAUNCH(agg_init_kernel, 148 * 8, 256, 0, st, a);
    }
    {
        ProfScope ps("agg_gather", st);
        TDT_LAUNCH(agg_gather_kernel, per_elem, 256, 0, st, a);
    }
    {
        // Small candidates: one direct kernel.  What is left for the sorts (all candidates with TDT_AGG_DIRECT=0) runs
        // as three parallel branches: the main stream and two side streams forked here and joined before agg_finalize
        // (inside a CUDA-graph capture they become parallel graph branches).  Every sort is a chain of small,
        // latency-bound kernels that leaves most of the machine idle; measured one after the other they took
        // 0.43 + 0.43 + 0.36 ms of the 2.46 ms call on the 30X set.
        ProfScope ps("agg_modes_names", st);
        const bool direct = a.direct_max > 0;
        // TDT_AGG_OVERLAP=0: direct kernel first, then the sorts of the big candidates (measurement aid)
        const char *ov = getenv("TDT_AGG_OVERLAP");
        const bool overlap = direct && !(ov && ov[0] == '0');
        if (direct && !overlap) TDT_LAUNCH(agg_direct_kernel, per_elem, 256, 0, st, a);
        const int64_t nseg_max = direct ? n / ((int64_t)a.direct_max + 1) + 1 : n;
        const int64_t *sub_dims = direct ? (const int64_t *)&a.small->d3 : (const int64_t *)&a.small->d2;
        // the run kernels of the big candidates: capped grid, grid-stride (usually there is next to nothing to do)
        const unsigned runs_grid = direct && per_elem > 148u * 8u ? 148u * 8u : per_elem;
        static thread_local cudaStream_t br[16][3] = {};
        static thread_local cudaEvent_t ev_fork[16] = {}, ev_join[16][3] = {};
        int dev = 0;
        TDT_CUDA(cudaGetDevice(&dev));
        const bool par = dev >= 0 && dev < 16;
        if (par && !br[dev][0]) {
            // high priority: the chains are short kernels that depend on each other; next to the direct kernel's tens
            // of thousands of CTAs they must not queue behind them
            int pr_least = 0, pr_greatest = 0;
            TDT_CUDA(cudaDeviceGetStreamPriorityRange(&pr_least, &pr_greatest));
            for (int i = 0; i < 3; i++) {
                TDT_CUDA(cudaStreamCreateWithPriority(&br[dev][i], cudaStreamNonBlocking, pr_greatest));
                TDT_CUDA(cudaEventCreateWithFlags(&ev_join[dev][i], cudaEventDisableTiming));
            }
            TDT_CUDA(cudaEventCreateWithFlags(&ev_fork[dev], cudaEventDisableTiming));
        }
        if (par) TDT_CUDA(cudaEventRecord(ev_fork[dev], st));
        for (int what = 0; what < 3; what++) {
            // overlap: all three chains on side streams, the caller's stream keeps the direct kernel; otherwise the
            // first chain stays on the caller's stream
            cudaStream_t bs = !par ? st : (overlap ? br[dev][what] : (what > 0 ? br[dev][what - 1] : st));
            if (bs != st) TDT_CUDA(cudaStreamWaitEvent(bs, ev_fork[dev], 0));
            const u32 *kin = direct ? (what == 0 ? a.bkA : (what == 1 ? a.bkB : a.bkN))
                                    : (what == 0 ? a.keyA : (what == 1 ? a.keyB : a.keyN));
            const int bits = (what == 2 ? a.name_bits : a.pos_bits) + 2;
            segsort_set_branch(what + 1);
            // (big candidates exceed the sort's counting path: no per-element segment ids needed)
            int rc = segsort_pairs(kin, nullptr, b_keys[what], b_vals[what], b_tmpK[what], b_tmpV[what],
                                   direct ? a.goff_big : a.goff, sub_dims, direct ? nullptr : a.c_grp, n, nseg_max, bits,
                                   b_temp[what], pl.sort, &a.small->err, bs);
            segsort_set_branch(0);
            if (rc) return rc;
            if (direct) {
                if (what == 0) TDT_LAUNCH((agg_runs_kernel<0, true>), runs_grid, 256, 0, bs, a, b_keys[what], b_vals[what]);
                else if (what == 1) TDT_LAUNCH((agg_runs_kernel<1, true>), runs_grid, 256, 0, bs, a, b_keys[what], b_vals[what]);
                else TDT_LAUNCH((agg_runs_kernel<2, true>), runs_grid, 256, 0, bs, a, b_keys[what], b_vals[what]);
            } else {
                if (what == 0) TDT_LAUNCH((agg_runs_kernel<0, false>), runs_grid, 256, 0, bs, a, b_keys[what], b_vals[what]);
                else if (what == 1) TDT_LAUNCH((agg_runs_kernel<1, false>), runs_grid, 256, 0, bs, a, b_keys[what], b_vals[what]);
                else TDT_LAUNCH((agg_runs_kernel<2, false>), runs_grid, 256, 0, bs, a, b_keys[what], b_vals[what]);
            }
            if (bs != st) {
                const int bi = overlap ? what : what - 1;
                TDT_CUDA(cudaEventRecord(ev_join[dev][bi], bs));
                if (!overlap) TDT_CUDA(cudaStreamWaitEvent(st, ev_join[dev][bi], 0));
            }
        }
        if (overlap) {
            // launched behind the chains: their first kernels are already queued when its CTAs start to fill the SMs
            TDT_LAUNCH(agg_direct_kernel, per_elem, 256, 0, st, a);
            if (par)
                for (int i = 0; i < 3; i++) TDT_CUDA(cudaStreamWaitEvent(st, ev_join[dev][i], 0));
        }
    }
    {
        ProfScope ps("agg_finalize", st);
        TDT_LAUNCH(agg_finalize_kernel, per_elem, 256, 0, st, a);
        TDT_LAUNCH(agg_status_kernel, 1, 1, 0, st, a);
    }
    return TDT_OK;
}

}  // namespace tdt

using namespace tdt;

extern "C" {

size_t tdt_aggregate_workspace_bytes(int64_t n, int32_t P) {
    if (n <= 0) return 0;
    return agg_plan(n, P < 1 ? 1 : P).total;
}

int tdt_cluster_aggregate(const int32_t *labels, const int32_t *posA, const int32_t *posB, const int32_t *span,
                          const int32_t *name_id, const uint8_t *flags, const int64_t *seg_off,
                          const uint8_t *same_chrom, int64_t n, int32_t P, int32_t max_ins_len, int32_t is_mp,
                          int32_t min_reads, int32_t max_pos, int32_t n_names, int32_t *cand_out,
                          int32_t *member_idx_out, int64_t *counts_out, void *ws, size_t ws_bytes, void *stream) {
    if (n < 0) return fail(TDT_E_ARG, "n = %lld is negative", (long long)n);
    if (n >= (1LL << 30)) return fail(TDT_E_ARG, "n = %lld exceeds the 2^30 signals one call supports", (long long)n);
    if (!counts_out) return fail(TDT_E_ARG, "counts_out is null");
    if (max_pos < 0 || n_names < 0) return fail(TDT_E_ARG, "max_pos / n_names must not be negative");
    if (max_pos >= (1 << 30)) return fail(TDT_E_ARG, "max_pos = %d: positions must stay below 2^30", max_pos);
    if (n == 0) {
        TDT_CUDA(cudaMemsetAsync(counts_out, 0, 4 * sizeof(int64_t), (cudaStream_t)stream));
        return TDT_OK;
    }
    if (P < 1) return fail(TDT_E_ARG, "P = %d pairs for %lld signals", P, (long long)n);
    if (!labels || !posA || !posB || !span || !name_id || !flags || !seg_off || !same_chrom || !cand_out ||
        !member_idx_out)
        return fail(TDT_E_ARG, "null pointer argument");
    if (((uintptr_t)span & 15) || ((uintptr_t)cand_out & 15))
        return fail(TDT_E_ARG, "span and cand_out must be 16-byte aligned");
    AggParams a = {};
    a.labels = labels; a.posA = posA; a.posB = posB; a.name_id = name_id;
    a.span = (const int4 *)span;
    a.flags = flags; a.same_chrom = same_chrom; a.seg_off = seg_off;
    a.n = n; a.P = P; a.max_ins_len = max_ins_len; a.is_mp = is_mp != 0; a.min_reads = min_reads;
    a.pos_bits = max_pos ? bit_width_u32((uint32_t)max_pos) : 30;
    a.name_bits = n_names ? bit_width_u32((uint32_t)n_names) : 30;
    if (a.pos_bits < 1) a.pos_bits = 1;
    if (a.name_bits < 1) a.name_bits = 1;
    a.sentinel = (1u << bit_width_u32((uint32_t)(2 * n))) - 1u;
    a.cand_out = cand_out; a.member_idx = member_idx_out; a.counts_out = counts_out;
    // TDT_AGG_DIRECT (read at every call): 0 = every candidate through the sub-sorts; k = direct kernel up to k members
    a.direct_max = TDT_AGG_DIRECT_DEFAULT;
    if (const char *e = getenv("TDT_AGG_DIRECT")) {
        if (e[0] >= '0' && e[0] <= '9') a.direct_max = atoi(e);
    }
    if (a.direct_max < 0) a.direct_max = 0;
    if (a.direct_max > AG_DIRECT_LIMIT) a.direct_max = AG_DIRECT_LIMIT;
    return aggregate_impl(a, ws, ws_bytes, (cudaStream_t)stream);
}

}  // extern "C"

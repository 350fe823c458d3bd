// tdt_cluster.cu -- TIDDIT signal clustering on B200 (sm_100a).
//
// Replaces, for all (chrA,chrB) pairs at once, the reference's per-pair
//     sorted(key=posA) -> DBSCAN.main -> labels back in insertion order
// (tiddit/tiddit_cluster.pyx:140-160, tiddit/DBSCAN.py:33-129).  The reference "DBSCAN" is a
// sliding-window run detector on the posA-sorted signals followed by the same detector on posB inside
// every x-cluster; the kernels below evaluate its closed form:
//
//   ok[i]      = the window of the next `reach` signals of i's segment stays within eps
//   run        = maximal interval of consecutive ok
//   label[j]   = index of the latest run whose start is <= j, if an ok lies in [j-m+1, j]; else noise
//
//   pack_keys_*          key = segment << bits | pos, value = insertion index
//   radix sort           (segment, posA) stable                                       [tdt_sort.cuh]
//   window_runs<X_PAIRS> eps-range query on posA + run labelling + compaction of the labelled signals
//                        into the y-pass sort input  key2 = x-run << bits | posB
//   radix sort           (x-run, posB) stable
//   window_runs<Y_SUBS>  eps-range query on posB inside every x-run + sub-run numbering
//   group_extra_scan     exclusive scan of max(sub-runs - 1, 0) over the x-runs (DBSCAN.py:121-122)
//   group_bases          the two id bases of every x-run (DBSCAN.py:113-117)
//   final_labels         reference ids scattered back to insertion order
#include "tdt_common.cuh"
#include "tdt_sort.cuh"

namespace tdt {

constexpr int WR_THREADS = 256;
constexpr int WR_WARPS = WR_THREADS / 32;
constexpr int WR_TILE = 4096;            // signals per CTA
constexpr int WR_WORDS = WR_TILE / 32;   // ballot words per tile
constexpr int WR_MAX_M = 8192;           // shared-memory halo limit: (TILE + 2*m) keys per CTA
constexpr uint32_t WR_TMA_CHUNK = 32768; // bytes per bulk copy

enum { OUT_X_PAIRS = 0, OUT_X_LABELS = 1, OUT_Y_SUBS = 2 };
enum { ERR_NONE = 0, ERR_RANGE_A = 1, ERR_RANGE_B = 2, ERR_PAIR = 3 };

struct WRParams {
    const void *keys;   // sorted keys (K*), padded to a multiple of 16 bytes past n
    int64_t n;
    int m;              // min_pts
    u64 eps;            // 0: nothing is ever within eps
    int shift;          // key >> shift = segment (pair or x-run)
    u64 *status;        // look-back tile states, zeroed
    u32 *ticket;        // tile ticket, zeroed
    // X_PAIRS / X_LABELS
    const int32_t *vals;   // insertion index per sorted position (nullptr: identity)
    const int32_t *posB;
    int bwB;
    int32_t max_pos;
    u64 *key2;
    int32_t *val2;
    int32_t *grp_pair;
    int32_t *gfirst;
    u32 *totals;           // [0] = #x-runs, [1] = #labelled signals
    int *err;
    int32_t *labels_out;   // X_LABELS
    int32_t *last_id_out;  // X_LABELS
    // Y_SUBS
    int32_t *ys;
    int32_t *gcnt_start;
    int64_t n_groups;
};

static size_t wr_smem_bytes(int m, size_t key_bytes) {
    const size_t HL = ((size_t)(m - 1) + 31) & ~(size_t)31;
    const size_t HR = ((size_t)m + 3) & ~(size_t)3;
    return (HL + WR_TILE + HR) * key_bytes + ((HL + WR_TILE) / 32 + 4 * WR_WORDS) * sizeof(u32);
}

// ----------------------------------------------------------------------------------------------
// The eps-range-query + run-labelling kernel (DBSCAN.py:40-62 for posA, :90-110 for posB).
//
// One CTA = one tile of WR_TILE sorted signals.  Thread 0 takes the tile ticket and issues 1-D TMA
// bulk copies (UBLKCP) of the tile plus a left halo of m-1 and a right halo of m keys into shared
// memory; every warp then evaluates 32 windows at a time and packs the outcome with a ballot:
//   X (reach m, DBSCAN.py:44 slices data[i+1:i+m+1], cut at the segment end):
//        ok[i] = seg(i+m) == seg(i) ? key[i+m] - key[i] < eps
//                                   : seg(i+m-1) == seg(i) && key[i+m-1] - key[i] < eps
//   Y (reach m-1, DBSCAN.py:93 slices y[i+1:i+m], never cut inside range(0, k-m+1)):
//        ok[i] = seg(i+m-1) == seg(i) && key[i+m-1] - key[i] < eps
// (sorted keys of one segment differ by exactly the coordinate difference; GENERAL evaluates the
// reference's max(|x[j]-x[i]|) literally for input that is not sorted.)
// The last m-1 signals of a segment are never ok, so neither run starts nor the m-1 look-back of the
// labelling rule can leak across segments, and a halo of m-1 on the left makes both local to the
// tile: only the NUMBER of run starts (and of labelled signals) before the tile is carried, by the
// single-pass look-back scan of tdt_common.cuh.
// ----------------------------------------------------------------------------------------------
template <typename K, int OUT, bool GENERAL>
__global__ void __launch_bounds__(WR_THREADS) window_runs_kernel(const WRParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t mbar;
    __shared__ int s_tile;
    __shared__ u32 s_pref[4];  // exS, exC, aggS, aggC

    const int m = p.m;
    const int HL = ((m - 1) + 31) & ~31;
    const int HR = (m + 3) & ~3;
    const int EXT = HL + WR_TILE;  // positions whose ok bit is evaluated here
    K *keys_s = (K *)smem_raw;
    u32 *okw = (u32 *)(keys_s + (HL + WR_TILE + HR));
    u32 *stw = okw + EXT / 32;
    u32 *cvw = stw + WR_WORDS;
    u32 *pfxS = cvw + WR_WORDS;
    u32 *pfxC = pfxS + WR_WORDS;

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int64_t n = p.n;
    const K *keys = (const K *)p.keys;

    if (threadIdx.x == 0) {
        s_tile = (int)atomicAdd(p.ticket, 1u);
        mbar_init(&mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    const int tile = s_tile;
    const int64_t tile_base = (int64_t)tile * WR_TILE;
    const int64_t ext_base = tile_base - HL;  // global index of shared position 0

    if (threadIdx.x == 0) {
        constexpr int64_t PER16 = 16 / (int64_t)sizeof(K);
        const int64_t n_pad = (n + PER16 - 1) / PER16 * PER16;
        const int64_t jlo = ext_base < 0 ? 0 : ext_base;
        int64_t jhi = tile_base + WR_TILE + HR;
        if (jhi > n_pad) jhi = n_pad;
        const uint32_t bytes = (uint32_t)((jhi - jlo) * (int64_t)sizeof(K));
        mbar_expect_tx(&mbar, bytes);
        const unsigned char *src = (const unsigned char *)(keys + jlo);
        unsigned char *dst = (unsigned char *)(keys_s + (jlo - ext_base));
        for (uint32_t off = 0; off < bytes; off += WR_TMA_CHUNK) {
            const uint32_t len = bytes - off < WR_TMA_CHUNK ? bytes - off : WR_TMA_CHUNK;
            tma_load_1d(dst + off, src + off, len, &mbar);
        }
    }
    mbar_wait(&mbar, 0);

    // ---- phase 1: one ballot word per 32 windows -------------------------------------------
    for (int w = warp; w < EXT / 32; w += WR_WARPS) {
        const int e = w * 32 + lane;
        const int64_t j = ext_base + e;
        bool ok = false;
        if (j >= 0 && j + m - 1 < n) {
            const K kj = keys_s[e];
            if (GENERAL) {
                // DBSCAN.py:44-51 literally: max |x[i+d] - x[i]| over the next m signals (cut at n)
                const int64_t left = n - 1 - j;
                const int cnt = left < (int64_t)m ? (int)left : m;
                u64 dmax = 0;
                for (int d = 1; d <= cnt; d++) {
                    const K kt = keys_s[e + d];
                    const u64 dd = kt > kj ? (u64)(kt - kj) : (u64)(kj - kt);
                    dmax = dd > dmax ? dd : dmax;
                }
                ok = dmax < p.eps;
            } else if (OUT == OUT_Y_SUBS) {
                const K kt = keys_s[e + m - 1];
                ok = ((kt >> p.shift) == (kj >> p.shift)) && ((u64)(kt - kj) < p.eps);
            } else {
                K kt = keys_s[e + m];
                if (j + m < n && (kt >> p.shift) == (kj >> p.shift)) {
                    ok = (u64)(kt - kj) < p.eps;
                } else {
                    kt = keys_s[e + m - 1];
                    ok = ((kt >> p.shift) == (kj >> p.shift)) && ((u64)(kt - kj) < p.eps);
                }
            }
        }
        const u32 word = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) okw[w] = word;
    }
    __syncthreads();

    // ---- phase 2: run-start bits and "labelled" bits of the tile's own words ----------------
    for (int tw = warp; tw < WR_WORDS; tw += WR_WARPS) {
        const int w = HL / 32 + tw;
        const u32 cur = okw[w], prev = okw[w - 1];
        const u32 st = cur & ~((cur << 1) | (prev >> 31));
        bool c;
        if (m <= 32) {
            // an ok among the m positions ending at this lane: bits [33+lane-m, 32+lane] of prev:cur
            const u64 v = ((u64)cur << 32) | (u64)prev;
            const u64 mask = (m == 32) ? 0xffffffffull : ((1ull << m) - 1ull);
            c = ((v >> (33 + lane - m)) & mask) != 0ull;
        } else {
            c = (cur & lanemask_le()) != 0u;
            if (!c) {
                for (int d = 1; d <= HL / 32; d++) {
                    const u32 ww = okw[w - d];
                    if (ww) {
                        const int last = 32 * (w - d) + 31 - __clz((int)ww);
                        c = (32 * w + lane - last) <= m - 1;
                        break;
                    }
                }
            }
        }
        const u32 cv = __ballot_sync(0xffffffffu, c);
        if (lane == 0) {
            stw[tw] = st;
            cvw[tw] = cv;
        }
    }
    __syncthreads();

    // ---- phase 3: word prefixes inside the tile, then the carry from earlier tiles ----------
    if (warp == 0) {
        constexpr int PER = WR_WORDS / 32;
        u32 s[PER], c[PER], ts = 0, tc = 0;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            s[k] = __popc(stw[lane * PER + k]);
            c[k] = __popc(cvw[lane * PER + k]);
            ts += s[k];
            tc += c[k];
        }
        u32 is = ts, ic = tc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, is, o);
            const u32 b = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) {
                is += a;
                ic += b;
            }
        }
        u32 es = is - ts, ec = ic - tc;
#pragma unroll
        for (int k = 0; k < PER; k++) {
            pfxS[lane * PER + k] = es;
            pfxC[lane * PER + k] = ec;
            es += s[k];
            ec += c[k];
        }
        const u32 aggS = __shfl_sync(0xffffffffu, is, 31);
        const u32 aggC = __shfl_sync(0xffffffffu, ic, 31);
        u32 exS, exC;
        lookback(p.status, tile, aggS, aggC, exS, exC);
        if (lane == 0) {
            s_pref[0] = exS;
            s_pref[1] = exC;
            s_pref[2] = aggS;
            s_pref[3] = aggC;
        }
    }
    __syncthreads();
    const u32 exS = s_pref[0], exC = s_pref[1];

    // ---- phase 4: outputs ---------------------------------------------------------------------
    for (int tw = warp; tw < WR_WORDS; tw += WR_WARPS) {
        const int e = HL + tw * 32 + lane;
        const int64_t j = tile_base + tw * 32 + lane;
        if (j >= n) continue;
        const u32 st = stw[tw], cv = cvw[tw];
        const bool covered = (cv >> lane) & 1u;
        const bool is_start = (st >> lane) & 1u;
        const u32 starts_before = exS + pfxS[tw] + __popc(st & lanemask_lt());
        const u32 starts_incl = starts_before + (is_start ? 1u : 0u);
        const K kj = keys_s[e];
        const bool head = (j == 0) || ((keys_s[e - 1] >> p.shift) != (kj >> p.shift));
        if (OUT == OUT_Y_SUBS) {
            p.ys[j] = covered ? (int32_t)starts_incl : 0;
            if (head) p.gcnt_start[(int64_t)(kj >> p.shift)] = (int32_t)starts_before;
        } else {
            const int32_t idx = p.vals ? p.vals[j] : (int32_t)j;
            if (OUT == OUT_X_LABELS) {
                p.labels_out[idx] = covered ? (int32_t)(starts_incl - 1u) : -1;
            } else {
                if (head) p.gfirst[(int64_t)(kj >> p.shift)] = (int32_t)starts_before;
                if (is_start) p.grp_pair[starts_before] = (int32_t)(kj >> p.shift);
                if (covered) {
                    const u32 pos = exC + pfxC[tw] + __popc(cv & lanemask_lt());
                    const int32_t yb = p.posB[idx];
                    if (yb < 0 || yb > p.max_pos) atomicMax(p.err, ERR_RANGE_B);
                    p.key2[pos] = ((u64)(starts_incl - 1u) << p.bwB) | (u64)(u32)yb;
                    p.val2[pos] = idx;
                }
            }
        }
    }
    if (threadIdx.x == 0 && tile_base + WR_TILE >= n) {  // the last tile publishes the totals
        const u32 totS = exS + s_pref[2], totC = exC + s_pref[3];
        if (OUT == OUT_Y_SUBS) {
            p.gcnt_start[p.n_groups] = (int32_t)totS;
        } else {
            if (p.totals) {
                p.totals[0] = totS;
                p.totals[1] = totC;
            }
            if (OUT == OUT_X_LABELS && p.last_id_out) *p.last_id_out = (int32_t)totS - 1;
        }
    }
}

// ----------------------------------------------------------------------------------------------
// key packing (tiddit_cluster.pyx:152: the sort key is posA inside one (chrA,chrB) list)
// ----------------------------------------------------------------------------------------------
template <typename K>
__global__ void pack_keys_seg_kernel(const int32_t *__restrict__ posA, const int64_t *__restrict__ seg_off, int P,
                                     int64_t n, int bwA, int32_t max_pos, K *__restrict__ keys,
                                     int32_t *__restrict__ vals, int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int32_t a = posA[j];
        if (a < 0 || a > max_pos) atomicMax(err, ERR_RANGE_A);
        int lo = 0, hi = P;  // largest s in [0, P) with seg_off[s] <= j
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (seg_off[mid] <= j) lo = mid; else hi = mid;
        }
        keys[j] = ((K)lo << bwA) | (K)(u32)a;
        vals[j] = (int32_t)j;
    }
}

template <typename K>
__global__ void pack_keys_keyed_kernel(const int32_t *__restrict__ posA, const int32_t *__restrict__ pair_id, int P,
                                       int64_t n, int bwA, int32_t max_pos, K *__restrict__ keys,
                                       int32_t *__restrict__ vals, int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int32_t a = posA[j];
        int32_t s = pair_id[j];
        if (a < 0 || a > max_pos) atomicMax(err, ERR_RANGE_A);
        if (s < 0 || s >= P) {
            atomicMax(err, ERR_PAIR);
            s = 0;
        }
        keys[j] = ((K)s << bwA) | (K)(u32)a;
        vals[j] = (int32_t)j;
    }
}

// x as given (no sort): the stand-alone DBSCAN.py entry points cluster the caller's order
__global__ void pack_plain_kernel(const int32_t *__restrict__ x, int64_t n, int32_t max_pos, u32 *__restrict__ keys,
                                  int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int32_t a = x[j];
        if (a < 0 || a > max_pos) atomicMax(err, ERR_RANGE_A);
        keys[j] = (u32)a;
    }
}

// segments without signals take the run count of the next populated segment; gfirst[P] = #runs
__global__ void fill_gfirst_kernel(int32_t *gfirst, int P, const u32 *totals) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s > P) return;
    if (s == P) {
        gfirst[P] = (int32_t)totals[0];
        return;
    }
    if (gfirst[s] >= 0) return;
    int q = s + 1;
    while (q < P && gfirst[q] < 0) q++;
    // populated entries are only ever read here, unpopulated ones only written
    gfirst[s] = q < P ? gfirst[q] : (int32_t)totals[0];
}

// ----------------------------------------------------------------------------------------------
// DBSCAN.py:121-122: after x-run g the running id grows by max(sub-runs(g) - 1, 0).
// X[g] = sum over g' < g, single-pass look-back scan, 2048 x-runs per CTA.
// ----------------------------------------------------------------------------------------------
constexpr int GS_THREADS = 256;
constexpr int GS_ITEMS = 8;
constexpr int GS_TILE = GS_THREADS * GS_ITEMS;

__global__ void __launch_bounds__(GS_THREADS) group_extra_scan_kernel(const int32_t *__restrict__ gcnt_start,
                                                                      int64_t G, int32_t *__restrict__ X,
                                                                      u64 *status, u32 *ticket) {
    __shared__ int s_tile;
    __shared__ u32 s_warp[GS_THREADS / 32];
    __shared__ u32 s_ex;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_tile = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int tile = s_tile;
    const int64_t g0 = (int64_t)tile * GS_TILE + (int64_t)threadIdx.x * GS_ITEMS;
    u32 v[GS_ITEMS], tsum = 0;
    int32_t prev = g0 < G ? gcnt_start[g0] : 0;
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t g = g0 + k;
        u32 extra = 0;
        if (g < G) {
            const int32_t next = gcnt_start[g + 1];
            const int32_t subs = next - prev;
            extra = subs > 1 ? (u32)(subs - 1) : 0u;
            prev = next;
        }
        v[k] = extra;
        tsum += extra;
    }
    u32 inc = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 a = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += a;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        u32 w = lane < GS_THREADS / 32 ? s_warp[lane] : 0u;
        u32 wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 a = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += a;
        }
        const u32 agg = __shfl_sync(0xffffffffu, wi, 31);
        if (lane < GS_THREADS / 32) s_warp[lane] = wi - w;  // exclusive warp offsets
        u32 exA, exB;
        lookback(status, tile, agg, 0u, exA, exB);
        if (lane == 0) s_ex = exA;
    }
    __syncthreads();
    u32 run = s_ex + s_warp[warp] + (inc - tsum);
#pragma unroll
    for (int k = 0; k < GS_ITEMS; k++) {
        const int64_t g = g0 + k;
        if (g < G) X[g] = (int32_t)run;
        run += v[k];
        if (g == G - 1) X[G] = (int32_t)run;
    }
}

// DBSCAN.py:113-117 per x-run g of pair s (x ids of the pair are gfirst[s] .. gfirst[s+1]-1):
//   sub-run 1 keeps the pair-local x id            base1 = g - gfirst[s]
//   sub-run k >= 2 gets  k + cluster_id - 1        base2 + k, base2 = nx + (X[g] - X[gfirst[s]]) - 2
__global__ void group_bases_kernel(const int32_t *__restrict__ grp_pair, const int32_t *__restrict__ gfirst,
                                   const int32_t *__restrict__ X, int64_t G, int32_t *__restrict__ base1,
                                   int32_t *__restrict__ base2) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    const int32_t s = grp_pair[g];
    const int32_t gf = gfirst[s];
    const int32_t nx = gfirst[s + 1] - gf;
    base1[g] = (int32_t)g - gf;
    base2[g] = nx + (X[g] - X[gf]) - 2;
}

__global__ void final_labels_kernel(const u64 *__restrict__ key2, const int32_t *__restrict__ val2,
                                    const int32_t *__restrict__ ys, const int32_t *__restrict__ gcnt_start,
                                    const int32_t *__restrict__ base1, const int32_t *__restrict__ base2, int bwB,
                                    int64_t n2, int32_t *__restrict__ labels_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += stride) {
        const int32_t y = ys[j];
        int32_t label = -1;
        if (y > 0) {
            const int64_t g = (int64_t)(key2[j] >> bwB);
            const int32_t sub = y - gcnt_start[g];
            label = sub == 1 ? base1[g] : base2[g] + sub;
        }
        labels_out[val2[j]] = label;
    }
}

// stand-alone y-pass (DBSCAN.py:66-123 on caller-supplied x ids): key2 = id << bits | y, noise is
// given the id cluster_id + 1 so that it sorts behind every real cluster
__global__ void pack_y_kernel(const int32_t *__restrict__ y, const int32_t *__restrict__ labels, int64_t n,
                              int32_t cluster_id, int bwB, int32_t max_pos, u64 *__restrict__ key2,
                              int32_t *__restrict__ val2, int *err) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += stride) {
        const int32_t b = y[j];
        int32_t l = labels[j];
        if (b < 0 || b > max_pos) atomicMax(err, ERR_RANGE_B);
        if (l < -1 || l > cluster_id) {
            atomicMax(err, ERR_PAIR);
            l = -1;
        }
        if (l < 0) l = cluster_id + 1;
        key2[j] = ((u64)(u32)l << bwB) | (u64)(u32)b;
        val2[j] = (int32_t)j;
    }
}

// ids that no signal carries have no head to write their start count: take the next one's
__global__ void fill_gcnt_kernel(int32_t *gcnt_start, int64_t G) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G || gcnt_start[g] >= 0) return;
    int64_t q = g + 1;
    while (q < G && gcnt_start[q] < 0) q++;
    gcnt_start[g] = gcnt_start[q];  // q == G holds the total, written by the last tile
}

// DBSCAN.py:113-117 with caller ids: sub-run 1 keeps id g, sub-run k >= 2 gets k + (cluster_id + X[g]) - 1;
// the noise group (g == G - 1) stays noise whatever its windows say
__global__ void group_bases_plain_kernel(const int32_t *__restrict__ X, int64_t G, int32_t cluster_id,
                                         int32_t *__restrict__ base1, int32_t *__restrict__ base2,
                                         int32_t *cluster_id_out) {
    const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= G) return;
    base1[g] = (int32_t)g;
    base2[g] = cluster_id + X[g] - 1;
    if (g == G - 1) *cluster_id_out = cluster_id + X[g];
}

__global__ void final_labels_plain_kernel(const u64 *__restrict__ key2, const int32_t *__restrict__ val2,
                                          const int32_t *__restrict__ ys, const int32_t *__restrict__ gcnt_start,
                                          const int32_t *__restrict__ base1, const int32_t *__restrict__ base2,
                                          int bwB, int64_t n2, int64_t noise_group, int32_t *__restrict__ labels_out) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; j < n2; j += stride) {
        const int32_t y = ys[j];
        const int64_t g = (int64_t)(key2[j] >> bwB);
        int32_t label = -1;
        if (y > 0 && g != noise_group) {
            const int32_t sub = y - gcnt_start[g];
            label = sub == 1 ? base1[g] : base2[g] + sub;
        }
        labels_out[val2[j]] = label;
    }
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
struct ClusterPlan {
    int64_t n, n_pad;
    int P;
    size_t keys_bytes, vals_bytes, status_bytes, small_bytes, gfirst_bytes, grp_bytes, sort_bytes;
    size_t total;
};

static int64_t wr_tiles(int64_t n) { return (n + WR_TILE - 1) / WR_TILE; }

static ClusterPlan make_plan(int64_t n, int32_t P) {
    ClusterPlan pl;
    pl.n = n;
    pl.P = P;
    pl.n_pad = (n + 63) & ~(int64_t)63;
    auto al = [](size_t b) { return (b + 255) & ~(size_t)255; };
    pl.keys_bytes = al((size_t)pl.n_pad * 8 + 256);
    pl.vals_bytes = al((size_t)pl.n_pad * 4 + 256);
    pl.status_bytes = al(((size_t)wr_tiles(n) + 64) * 8);
    pl.small_bytes = 256;
    pl.gfirst_bytes = al(((size_t)P + 2) * 4);
    pl.grp_bytes = al(((size_t)n / 2 + 4) * 4);
    pl.sort_bytes = al(sort_temp_bytes(n));
    pl.total = 2 * pl.keys_bytes + 2 * pl.vals_bytes + 2 * pl.status_bytes + pl.small_bytes + pl.gfirst_bytes +
               5 * pl.grp_bytes + pl.sort_bytes + 4096;
    return pl;
}

struct Small {  // device scalars, one 256-byte block
    u32 ticket[4];
    u32 totals[2];
    int err;
    int32_t last_id;
};

static int grid_for(int64_t n, int threads) {
    int64_t b = (n + threads - 1) / threads;
    const int64_t cap = 148 * 16;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}

template <typename K, int OUT, bool GENERAL>
static int launch_window_runs(const WRParams &p, cudaStream_t st) {
    const size_t smem = wr_smem_bytes(p.m, sizeof(K));
    static thread_local size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        TDT_CUDA(cudaFuncSetAttribute(window_runs_kernel<K, OUT, GENERAL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)smem));
        configured = smem;
    }
    const int64_t tiles = wr_tiles(p.n);
    TDT_LAUNCH((window_runs_kernel<K, OUT, GENERAL>), (unsigned)tiles, WR_THREADS, smem, st, p);
    return TDT_OK;
}

static int check_common(int64_t n, int32_t min_pts, int32_t eps, void *ws, size_t ws_bytes, size_t need) {
    (void)eps;
    if (n < 0) return fail(TDT_E_ARG, "n = %lld is negative", (long long)n);
    if (n > 2000000000LL) return fail(TDT_E_ARG, "n = %lld exceeds the 2e9 signals one call supports", (long long)n);
    if (min_pts < 2)
        return fail(TDT_E_ARG, "min_pts = %d: the reference raises ValueError (max() of an empty window) for m < 2",
                    min_pts);
    if (min_pts > WR_MAX_M) return fail(TDT_E_ARG, "min_pts = %d exceeds the supported maximum %d", min_pts, WR_MAX_M);
    if (n > 0 && (ws == nullptr || ws_bytes < need))
        return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, need);
    return TDT_OK;
}

static int err_to_code(int e) {
    switch (e) {
        case ERR_NONE: return TDT_OK;
        case ERR_RANGE_A: return fail(TDT_E_RANGE, "a posA / x coordinate is negative or above max_pos");
        case ERR_RANGE_B: return fail(TDT_E_RANGE, "a posB / y coordinate is negative or above max_pos");
        default: return fail(TDT_E_RANGE, "a pair id is outside [0, P)");
    }
}

// y-pass on the compacted (x-run, posB) pairs + final ids.  key2/val2 in bufs `cur`; `alt` free.
static int run_ypass(const ClusterPlan &pl, char *keys_cur, char *keys_alt, int32_t *vals_cur, int32_t *vals_alt,
                     int64_t n2, int64_t G, int bwB, int32_t eps, int32_t m, int32_t *grp_pair, int32_t *gfirst,
                     int32_t *gcnt_start, int32_t *X, int32_t *base1, int32_t *base2, u64 *status, Small *small,
                     void *sort_temp, int32_t *labels_out, cudaStream_t st) {
    const int bits = bwB + bit_width_u32((uint32_t)(G - 1));
    int which = 0;
    int rc;
    {
        ProfScope ps("sort_y", st);
        rc = sort_pairs<u64>((u64 *)keys_cur, (u64 *)keys_alt, vals_cur, vals_alt, n2, bits, sort_temp,
                             pl.sort_bytes, st, &which);
    }
    if (rc) return rc;
    u64 *k2 = (u64 *)(which ? keys_alt : keys_cur);
    int32_t *v2 = which ? vals_alt : vals_cur;
    int32_t *ys = (int32_t *)(which ? keys_cur : keys_alt);  // the buffer the sort left free

    TDT_CUDA(cudaMemsetAsync(status, 0, pl.status_bytes, st));
    WRParams p = {};
    p.keys = k2;
    p.n = n2;
    p.m = m;
    p.eps = eps > 0 ? (u64)eps : 0;
    p.shift = bwB;
    p.status = status;
    p.ticket = &small->ticket[1];
    p.ys = ys;
    p.gcnt_start = gcnt_start;
    p.n_groups = G;
    {
        ProfScope ps("window_runs_y", st);
        rc = launch_window_runs<u64, OUT_Y_SUBS, false>(p, st);
    }
    if (rc) return rc;

    u64 *status2 = status + (pl.status_bytes / 8);
    TDT_CUDA(cudaMemsetAsync(status2, 0, pl.status_bytes, st));
    const unsigned gs_tiles = (unsigned)((G + GS_TILE - 1) / GS_TILE);
    {
        ProfScope ps("group_ids", st);
        TDT_LAUNCH(group_extra_scan_kernel, gs_tiles, GS_THREADS, 0, st, gcnt_start, G, X, status2,
                   &small->ticket[2]);
        TDT_LAUNCH(group_bases_kernel, (unsigned)((G + 255) / 256), 256, 0, st, grp_pair, gfirst, X, G, base1, base2);
    }
    ProfScope ps("final_labels", st);
    TDT_LAUNCH(final_labels_kernel, grid_for(n2, 256), 256, 0, st, k2, v2, ys, gcnt_start, base1, base2, bwB, n2,
               labels_out);
    return TDT_OK;
}

template <typename K>
static int cluster_impl(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, const int32_t *pair_id,
                        int64_t n, int32_t P, int32_t eps, int32_t m, int32_t max_pos, int bwA, int32_t *labels_out,
                        void *ws, size_t ws_bytes, cudaStream_t st, bool presorted_plain) {
    const ClusterPlan pl = make_plan(n, P);
    Arena ar(ws, ws_bytes);
    char *keysA = ar.take<char>(pl.keys_bytes);
    char *keysB = ar.take<char>(pl.keys_bytes);
    int32_t *valsA = (int32_t *)ar.take<char>(pl.vals_bytes);
    int32_t *valsB = (int32_t *)ar.take<char>(pl.vals_bytes);
    u64 *status = (u64 *)ar.take<char>(2 * pl.status_bytes);
    Small *small = (Small *)ar.take<char>(pl.small_bytes);
    int32_t *gfirst = (int32_t *)ar.take<char>(pl.gfirst_bytes);
    int32_t *grp_pair = (int32_t *)ar.take<char>(pl.grp_bytes);
    int32_t *gcnt_start = (int32_t *)ar.take<char>(pl.grp_bytes);
    int32_t *X = (int32_t *)ar.take<char>(pl.grp_bytes);
    int32_t *base1 = (int32_t *)ar.take<char>(pl.grp_bytes);
    int32_t *base2 = (int32_t *)ar.take<char>(pl.grp_bytes);
    void *sort_temp = ar.take<char>(pl.sort_bytes);
    if (!sort_temp) return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);

    const int bwP = bit_width_u32((uint32_t)(P - 1));
    TDT_CUDA(cudaMemsetAsync(labels_out, 0xff, (size_t)n * 4, st));
    TDT_CUDA(cudaMemsetAsync(status, 0, pl.status_bytes, st));
    TDT_CUDA(cudaMemsetAsync(small, 0, pl.small_bytes, st));
    TDT_CUDA(cudaMemsetAsync(gfirst, 0xff, pl.gfirst_bytes, st));

    K *kcur;
    int32_t *vcur;
    char *kalt;
    int32_t *valt;
    if (presorted_plain) {
        // DBSCAN.main on the caller's order: no sort, identity permutation
        TDT_LAUNCH(pack_plain_kernel, grid_for(n, 256), 256, 0, st, posA, n, max_pos, (u32 *)keysA, &small->err);
        kcur = (K *)keysA;
        vcur = nullptr;
        kalt = keysB;
        valt = valsB;
    } else {
        if (seg_off) {
            ProfScope ps("pack_keys", st);
            TDT_LAUNCH(pack_keys_seg_kernel<K>, grid_for(n, 256), 256, 0, st, posA, seg_off, P, n, bwA, max_pos,
                       (K *)keysA, valsA, &small->err);
        } else {
            ProfScope ps("pack_keys", st);
            TDT_LAUNCH(pack_keys_keyed_kernel<K>, grid_for(n, 256), 256, 0, st, posA, pair_id, P, n, bwA, max_pos,
                       (K *)keysA, valsA, &small->err);
        }
        int which = 0;
        int rc;
        {
            ProfScope ps("sort_x", st);
            rc = sort_pairs<K>((K *)keysA, (K *)keysB, valsA, valsB, n, bwA + bwP, sort_temp, pl.sort_bytes, st,
                               &which);
        }
        if (rc) return rc;
        kcur = (K *)(which ? keysB : keysA);
        vcur = which ? valsB : valsA;
        kalt = which ? keysA : keysB;
        valt = which ? valsA : valsB;
    }

    const int bwB = bwA;
    WRParams p = {};
    p.keys = kcur;
    p.n = n;
    p.m = m;
    p.eps = eps > 0 ? (u64)eps : 0;
    p.shift = bwA;
    p.status = status;
    p.ticket = &small->ticket[0];
    p.vals = vcur;
    p.posB = posB;
    p.bwB = bwB;
    p.max_pos = max_pos;
    p.key2 = (u64 *)kalt;
    p.val2 = valt;
    p.grp_pair = grp_pair;
    p.gfirst = gfirst;
    p.totals = small->totals;
    p.err = &small->err;
    int rc;
    {
        ProfScope ps("window_runs_x", st);
        rc = presorted_plain ? launch_window_runs<K, OUT_X_PAIRS, true>(p, st)
                             : launch_window_runs<K, OUT_X_PAIRS, false>(p, st);
    }
    if (rc) return rc;

    Small h;
    TDT_CUDA(cudaMemcpyAsync(&h, small, sizeof(Small), cudaMemcpyDeviceToHost, st));
    TDT_CUDA(cudaStreamSynchronize(st));
    if (h.err) return err_to_code(h.err);
    const int64_t G = h.totals[0], n2 = h.totals[1];
    if (G == 0) return TDT_OK;  // everything is noise
    TDT_LAUNCH(fill_gfirst_kernel, (unsigned)((P + 1 + 255) / 256), 256, 0, st, gfirst, P, small->totals);
    // the sorted x keys / insertion indices are dead now: their buffers become the sort's alternates
    char *dead_keys = (char *)kcur == keysA ? keysA : keysB;
    int32_t *dead_vals = (valt == valsA) ? valsB : valsA;
    return run_ypass(pl, kalt, dead_keys, valt, dead_vals, n2, G, bwB, eps, m, grp_pair, gfirst, gcnt_start, X, base1,
                     base2, status, small, sort_temp, labels_out, st);
}

static int pos_bits(int32_t max_pos) { return bit_width_u32(max_pos > 0 ? (uint32_t)max_pos : 0x7fffffffu); }

}  // namespace tdt

using namespace tdt;

extern "C" {

size_t tdt_cluster_workspace_bytes(int64_t n, int32_t P) {
    if (n < 0) n = 0;
    if (P < 1) P = 1;
    return make_plan(n, P).total;
}

int tdt_cluster_labels(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, int64_t n, int32_t P,
                       int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_out, void *ws, size_t ws_bytes,
                       void *stream) {
    if (P < 1 && n > 0) return fail(TDT_E_ARG, "P = %d pairs for %lld signals", P, (long long)n);
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, eps, ws, ws_bytes, make_plan(n, P < 1 ? 1 : P).total);
    if (rc || n == 0) return rc;
    if (!posA || !posB || !seg_off || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    const int bwA = pos_bits(max_pos);
    if (max_pos == 0) max_pos = 0x7fffffff;
    const int bits = bwA + bit_width_u32((uint32_t)(P - 1));
    if (bits <= 32)
        return cluster_impl<u32>(posA, posB, seg_off, nullptr, n, P, eps, min_pts, max_pos, bwA, labels_out, ws,
                                 ws_bytes, (cudaStream_t)stream, false);
    return cluster_impl<u64>(posA, posB, seg_off, nullptr, n, P, eps, min_pts, max_pos, bwA, labels_out, ws, ws_bytes,
                             (cudaStream_t)stream, false);
}

int tdt_cluster_labels_keyed(const int32_t *posA, const int32_t *posB, const int32_t *pair_id, int64_t n, int32_t P,
                             int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_out, void *ws,
                             size_t ws_bytes, void *stream) {
    if (P < 1 && n > 0) return fail(TDT_E_ARG, "P = %d pairs for %lld signals", P, (long long)n);
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, eps, ws, ws_bytes, make_plan(n, P < 1 ? 1 : P).total);
    if (rc || n == 0) return rc;
    if (!posA || !posB || !pair_id || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    const int bwA = pos_bits(max_pos);
    if (max_pos == 0) max_pos = 0x7fffffff;
    const int bits = bwA + bit_width_u32((uint32_t)(P - 1));
    if (bits <= 32)
        return cluster_impl<u32>(posA, posB, nullptr, pair_id, n, P, eps, min_pts, max_pos, bwA, labels_out, ws,
                                 ws_bytes, (cudaStream_t)stream, false);
    return cluster_impl<u64>(posA, posB, nullptr, pair_id, n, P, eps, min_pts, max_pos, bwA, labels_out, ws, ws_bytes,
                             (cudaStream_t)stream, false);
}

int tdt_dbscan_main(const int32_t *x, const int32_t *y, int64_t n, int32_t eps, int32_t min_pts, int32_t max_pos,
                    int32_t *labels_out, void *ws, size_t ws_bytes, void *stream) {
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, eps, ws, ws_bytes, make_plan(n, 1).total);
    if (rc || n == 0) return rc;
    if (!x || !y || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    const int bw = pos_bits(max_pos);
    if (max_pos == 0) max_pos = 0x7fffffff;
    // keys are the plain 31-bit x values: a shift of 31 puts every signal in segment 0
    return cluster_impl<u32>(x, y, nullptr, nullptr, n, 1, eps, min_pts, max_pos, bw, labels_out, ws, ws_bytes,
                             (cudaStream_t)stream, true);
}

int tdt_xpass_labels(const int32_t *x, int64_t n, int32_t eps, int32_t min_pts, int32_t *labels_out,
                     int32_t *last_id_out, void *ws, size_t ws_bytes, void *stream) {
    int rc = check_common(n, min_pts, eps, ws, ws_bytes, make_plan(n, 1).total);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        if (last_id_out) TDT_CUDA(cudaMemsetAsync(last_id_out, 0xff, 4, st));
        return TDT_OK;
    }
    if (!x || !labels_out) return fail(TDT_E_ARG, "null pointer argument");
    const ClusterPlan pl = make_plan(n, 1);
    Arena ar(ws, ws_bytes);
    char *keysA = ar.take<char>(pl.keys_bytes);
    ar.take<char>(pl.keys_bytes);
    ar.take<char>(2 * pl.vals_bytes);
    u64 *status = (u64 *)ar.take<char>(2 * pl.status_bytes);
    Small *small = (Small *)ar.take<char>(pl.small_bytes);
    TDT_CUDA(cudaMemsetAsync(status, 0, pl.status_bytes, st));
    TDT_CUDA(cudaMemsetAsync(small, 0, pl.small_bytes, st));
    TDT_LAUNCH(pack_plain_kernel, grid_for(n, 256), 256, 0, st, x, n, 0x7fffffff, (u32 *)keysA, &small->err);
    WRParams p = {};
    p.keys = keysA;
    p.n = n;
    p.m = min_pts;
    p.eps = eps > 0 ? (u64)eps : 0;
    p.shift = 31;
    p.status = status;
    p.ticket = &small->ticket[0];
    p.labels_out = labels_out;
    p.last_id_out = last_id_out;
    return launch_window_runs<u32, OUT_X_LABELS, true>(p, st);
}

int tdt_ypass_labels(const int32_t *y, int64_t n, int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_io,
                     int32_t *cluster_id_io, void *ws, size_t ws_bytes, void *stream) {
    if (max_pos < 0) return fail(TDT_E_ARG, "max_pos = %d is negative", max_pos);
    int rc = check_common(n, min_pts, eps, ws, ws_bytes, make_plan(n, 1).total);
    if (rc || n == 0) return rc;
    if (!y || !labels_io || !cluster_id_io) return fail(TDT_E_ARG, "null pointer argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int bwB = pos_bits(max_pos);
    if (max_pos == 0) max_pos = 0x7fffffff;
    int32_t cluster_id = -1;
    TDT_CUDA(cudaMemcpyAsync(&cluster_id, cluster_id_io, 4, cudaMemcpyDeviceToHost, st));
    TDT_CUDA(cudaStreamSynchronize(st));
    if (cluster_id < -1 || (int64_t)cluster_id >= n)
        return fail(TDT_E_ARG, "cluster_id = %d is not the id count of %lld signals", cluster_id, (long long)n);

    const ClusterPlan pl = make_plan(n, 1);
    Arena ar(ws, ws_bytes);
    char *keysA = ar.take<char>(pl.keys_bytes);
    char *keysB = ar.take<char>(pl.keys_bytes);
    int32_t *valsA = (int32_t *)ar.take<char>(pl.vals_bytes);
    int32_t *valsB = (int32_t *)ar.take<char>(pl.vals_bytes);
    u64 *status = (u64 *)ar.take<char>(2 * pl.status_bytes);
    Small *small = (Small *)ar.take<char>(pl.small_bytes);
    ar.take<char>(pl.gfirst_bytes);
    // G = cluster_id + 2 ids (noise last) can exceed n/2 + 4 only for ids without signals; size check
    const int64_t G = (int64_t)cluster_id + 2;
    int32_t *grp = (int32_t *)ar.take<char>(5 * pl.grp_bytes);
    void *sort_temp = ar.take<char>(pl.sort_bytes);
    if (!sort_temp) return fail(TDT_E_WORKSPACE, "workspace of %zu bytes given, %zu needed", ws_bytes, pl.total);
    const size_t per = (size_t)(G + 2);
    if (4 * per * 4 > 5 * pl.grp_bytes)
        return fail(TDT_E_ARG, "cluster_id = %d: more ids than the workspace of %lld signals holds", cluster_id,
                    (long long)n);
    int32_t *gcnt_start = grp, *X = grp + per, *base1 = grp + 2 * per, *base2 = grp + 3 * per;

    TDT_CUDA(cudaMemsetAsync(status, 0, 2 * pl.status_bytes, st));
    TDT_CUDA(cudaMemsetAsync(small, 0, pl.small_bytes, st));
    TDT_CUDA(cudaMemsetAsync(gcnt_start, 0xff, per * 4, st));
    TDT_LAUNCH(pack_y_kernel, grid_for(n, 256), 256, 0, st, y, labels_io, n, cluster_id, bwB, max_pos, (u64 *)keysA,
               valsA, &small->err);
    const int bits = bwB + bit_width_u32((uint32_t)(G - 1));
    int which = 0;
    rc = sort_pairs<u64>((u64 *)keysA, (u64 *)keysB, valsA, valsB, n, bits, sort_temp, pl.sort_bytes, st, &which);
    if (rc) return rc;
    u64 *k2 = (u64 *)(which ? keysB : keysA);
    int32_t *v2 = which ? valsB : valsA;
    int32_t *ys = (int32_t *)(which ? keysA : keysB);
    WRParams p = {};
    p.keys = k2;
    p.n = n;
    p.m = min_pts;
    p.eps = eps > 0 ? (u64)eps : 0;
    p.shift = bwB;
    p.status = status;
    p.ticket = &small->ticket[1];
    p.ys = ys;
    p.gcnt_start = gcnt_start;
    p.n_groups = G;
    rc = launch_window_runs<u64, OUT_Y_SUBS, false>(p, st);
    if (rc) return rc;
    TDT_LAUNCH(fill_gcnt_kernel, (unsigned)((G + 255) / 256), 256, 0, st, gcnt_start, G);
    u64 *status2 = status + (pl.status_bytes / 8);
    TDT_LAUNCH(group_extra_scan_kernel, (unsigned)((G + GS_TILE - 1) / GS_TILE), GS_THREADS, 0, st, gcnt_start, G, X,
               status2, &small->ticket[2]);
    TDT_LAUNCH(group_bases_plain_kernel, (unsigned)((G + 255) / 256), 256, 0, st, X, G, cluster_id, base1, base2,
               cluster_id_io);
    TDT_LAUNCH(final_labels_plain_kernel, grid_for(n, 256), 256, 0, st, k2, v2, ys, gcnt_start, base1, base2, bwB, n,
               G - 1, labels_io);
    int herr = 0;
    TDT_CUDA(cudaMemcpyAsync(&herr, &small->err, 4, cudaMemcpyDeviceToHost, st));
    TDT_CUDA(cudaStreamSynchronize(st));
    return err_to_code(herr);
}

}  // extern "C"

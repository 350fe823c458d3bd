// tdt_segsort3.cuh -- third generation of the LARGE-segment chain of the segmented stable sort (sm_100a):
// most-significant-digit partition rounds + ONE shared-memory finish per element, instead of a histogram read and
// four stable LSD passes.  (Included by tdt_segsort.cuh inside TDT_SEGSORT_IMPL; the tiny / small-segment kernels of
// tdt_segsort.cuh are unchanged.)
//
// Why: the LSD passes are bound by their ranking work (~111 thread-instructions per element and pass, DRAM at 26 % of
// peak), so the way to go faster is fewer ranking rounds per element, not fewer bytes.
//
//   round r (r = 0..ceil(key_bits/8)-1)   the queued RANGES (round 0: segments of > M3_CAP elements; later: buckets of
//       the previous round that are still > M3_CAP) are cut by the 8-bit digit at bit `shift` (round 0: the TOP eight
//       bits of the key range): m3_hist_kernel (digit counts per range), m3_plan_kernel (exclusive scan; consecutive
//       buckets are grouped greedily into BATCHES of <= M3_CAP elements for the finish kernel, larger buckets are queued
//       for round r + 1), m3_pass_kernel (the stable onesweep pass of tdt_segsort.cuh with this digit).  A range whose
//       digit was the lowest bits of the key is sorted by the pass alone (stable: equal keys keep their order).
//   finish   m3_finish_kernel: a CTA (512 threads, 2 per SM) takes a batch -- a contiguous element range that holds
//       every element of a contiguous KEY range [klo, klo + nslots << sh) -- and sorts it in shared memory with ONE
//       ranking round whatever the key width: slot = (key - klo) >> sh (up to 8192 slots, about one element each for
//       spread-out keys), a shared-memory atomic per element counts the slots, a block scan turns counts into slot
//       starts, a second atomic places (key, tie-break) at an arbitrary position inside its slot, and the output loop
//       ranks every element among the (mostly 1-3) members of its slot by direct 64-bit comparison.  The tie-break
//       is the value when values grow with the position (vals_in == nullptr: the value IS the element index), else
//       the position in the batch -- so the result is the STABLE order although the atomics are not ordered.
//       A batch whose slots are crowded (sum of squared slot counts > 64 x elements: pile-ups, heavy duplicates) is
//       sorted by a bitonic network on the same 64-bit words instead: bounded work for any input.
//
// Data moves in -> tmp (round 0) -> out (round 1) -> tmp -> out; a batch is read from wherever its range currently
// lies ("level") and written to out -- in place when it already lies there (all loads of a batch precede its stores).
// Rounds >= 1 (only pile-ups and pairs beyond 2 M signals need them) run on a forked side stream next to the finish
// kernel of the round-0 batches; their own batches form a second list finished on that stream.
#pragma once

namespace tdt {

constexpr int M3_FLAG_COPY = 1;
constexpr int M3_FLAG_EQUAL = 2;   // every element of input segment `sh` whose key is `klo`, in input order
constexpr size_t M3_UPASS_SMEM = (size_t)SS_TILE * 8 + 256 * 4 * 2 + 256 * 8 + 64;
constexpr int M3_PAD = 64;   // all-ones words behind the staged elements (see m3_rank_in_slot)
constexpr size_t M3_FIN_SMEM = (size_t)(M3_CAP + M3_PAD) * 8 + (size_t)M3_NSLOT * 4 + 64 * 4;

__device__ __forceinline__ void m3_level_ptrs(const SSArgs &a, int level, const uint32_t *&k, const int32_t *&v) {
    if (level == 0) {
        k = a.keys_in;
        v = a.vals_in;
    } else if (level & 1) {
        k = a.keys_tmp;
        v = a.vals_tmp;
    } else {
        k = a.keys_out;
        v = a.vals_out;
    }
}

// ---- digit counts of every queued range ------------------------------------------------------------------------
// A CTA owns a contiguous run of tiles and flushes its shared histogram only when the range changes.
__global__ void __launch_bounds__(SS_THREADS) m3_hist_kernel(SSArgs a, int round) {
    __shared__ uint32_t h[256];
    const int n_tiles = a.m3.cnt->n_tiles[round];
    if (n_tiles <= 0) return;
    const int per = (n_tiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_lo = (int)blockIdx.x * per;
    const int t_hi = t_lo + per < n_tiles ? t_lo + per : n_tiles;
    if (t_lo >= t_hi) return;
    const uint32_t *src;
    const int32_t *src_v;
    m3_level_ptrs(a, round, src, src_v);
    (void)src_v;
    if (threadIdx.x < 256) h[threadIdx.x] = 0u;
    __syncthreads();
    int cur = a.m3.tile_rng[round][t_lo].rng;
    for (int tile = t_lo; tile < t_hi; tile++) {
        const M3Tile T = a.m3.tile_rng[round][tile];
        const int r = T.rng;
        if (r != cur) {   // uniform over the CTA
            __syncthreads();
            if (threadIdx.x < 256) {
                const uint32_t v = h[threadIdx.x];
                if (v) atomicAdd(a.m3.hist[round] + (size_t)cur * 256 + threadIdx.x, v);
                h[threadIdx.x] = 0u;
            }
            __syncthreads();
            cur = r;
        }
        const M3Range R = a.m3.rng[round][r];   // (klo, shift): needed after the keys are on their way
        const int64_t t0 = T.t0;
        const int cnt = T.cnt;
        if (cnt == SS_TILE) {   // all loads of the tile in flight before the first atomic
            uint32_t k[SS_CHUNKS];
#pragma unroll
            for (int c = 0; c < SS_CHUNKS; c++) k[c] = src[t0 + c * SS_THREADS + threadIdx.x];
            bool bad = false;
#pragma unroll
            for (int c = 0; c < SS_CHUNKS; c++) {
                bad = bad || (a.key_bits < 32 && (k[c] >> a.key_bits));
                uint32_t d = (k[c] - R.klo) >> R.shift;
                d = d > 255u ? 255u : d;
                atomicAdd(&h[d], 1u);
            }
            if (round == 0 && bad) atomicMax(a.err, SS_ERR_KEY_RANGE);
        } else {
            for (int e = threadIdx.x; e < cnt; e += SS_THREADS) {
                const uint32_t key = src[t0 + e];
                if (round == 0 && a.key_bits < 32 && (key >> a.key_bits)) atomicMax(a.err, SS_ERR_KEY_RANGE);
                uint32_t d = (key - R.klo) >> R.shift;
                d = d > 255u ? 255u : d;
                atomicAdd(&h[d], 1u);
            }
        }
    }
    __syncthreads();
    if (threadIdx.x < 256) {
        const uint32_t v = h[threadIdx.x];
        if (v) atomicAdd(a.m3.hist[round] + (size_t)cur * 256 + threadIdx.x, v);
    }
}

// ---- plan: bucket starts, finish batches, the next round's ranges -------------------------------------------------
// writes batch `idx` of `list` (the caller reserved the index with one atomicAdd on n_batches for all of its batches:
// an atomic per batch cost ~0.5 us each in the single lane that walks a range's buckets -- 140 us per round)
__device__ __forceinline__ void m3_emit_batch(const SSArgs &a, int list, int32_t idx, int64_t start, int32_t count,
                                              uint32_t klo, uint64_t span, int level, int flags) {
    if (idx >= a.m3.batch_max) {
        atomicMax(a.err, SS_ERR_INTERNAL);
        return;
    }
    int bits = span > 1 ? 64 - __clzll((long long)(span - 1)) : 0;
    M3Batch B;
    B.start = start;
    B.count = count;
    B.klo = klo;
    B.sh = bits > M3_SLOT_BITS ? bits - M3_SLOT_BITS : 0;
    B.nslots = (int32_t)(((span ? span : 1) - 1) >> B.sh) + 1;
    B.level = level;
    B.flags = flags;
    a.m3.batch[list][idx] = B;
}

// one warp per range
__global__ void __launch_bounds__(256) m3_plan_kernel(SSArgs a, int round, int dst_level) {
    __shared__ __align__(16) uint32_t s_cnt[8][256], s_ex[8][256];
    __shared__ uint8_t s_push[8][256];
    __shared__ int32_t s_ptile[8][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int ri = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
    if (ri >= a.m3.cnt->n_rng[round]) return;   // uniform over the warp
    const M3Range R = a.m3.rng[round][ri];
    uint32_t *H = a.m3.hist[round] + (size_t)ri * 256 + lane * 8;
    uint32_t v[8], t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        v[i] = H[i];
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
        H[i] = ex;   // the pass kernel reads exclusive bucket starts
        s_cnt[w][lane * 8 + i] = v[i];
        s_ex[w][lane * 8 + i] = ex;
        ex += v[i];
    }
    __syncwarp();
    const int list = round == 0 ? 0 : 1;
    const bool byval = a.m3_byval != 0;
    if (R.shift == 0 && !byval) {
        // the digit was the lowest bits of the key: the stable pass leaves the range sorted.  Where it leaves it in
        // the scratch buffers, copy batches bring it home.
        if (dst_level & 1) {
            const int nch = (R.size + M3_CAP - 1) / M3_CAP;
            int32_t base = 0;
            if (lane == 0) base = atomicAdd(&a.m3.cnt->n_batches[list], nch);
            base = __shfl_sync(0xffffffffu, base, 0);
            for (int c = lane; c < nch; c += 32) {
                const int32_t left = R.size - c * M3_CAP;
                m3_emit_batch(a, list, base + c, R.start + (int64_t)c * M3_CAP, left < M3_CAP ? left : M3_CAP, 0u, 1,
                              dst_level, M3_FLAG_COPY);
            }
        }
        return;
    }
    // ---- grouping, all 32 lanes (a single lane walking the 256 buckets was a 12 us dependent chain per round) --------
    // Buckets are ORDINARY (1 .. M3_CAP / 4 elements), DENSE (more, up to M3_CAP: a batch of their own -- grouped with
    // sparse neighbours they would share a wide key span, i.e. coarse slots, and crowd a few of them) or QUEUED for the
    // next round (beyond M3_CAP -- or a pile-up: dense AND > 8x the range's average bucket, a narrow hot region that one
    // more cut turns into batches with a slot per position instead of a ~100 us bitonic batch).  Dense and queued
    // buckets end a RUN of ordinary buckets; inside a run, bucket b joins group (Pi(b) - 1) / (3/4 M3_CAP), Pi = inclusive
    // prefix of the run's counts: a group holds < 3/4 M3_CAP + M3_CAP / 4 elements and, except the last of a run, more
    // than M3_CAP / 2.
    constexpr uint32_t GRP = (uint32_t)(M3_CAP - M3_CAP / 4);
    const uint32_t pile = (uint32_t)R.size / 32u;
    uint32_t kind[8];   // 0 empty, 1 ordinary, 2 dense, 3 queued
    uint32_t nord = 0, nbrk = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t c = v[i];
        const bool over = c > (uint32_t)M3_CAP || (R.shift > 0 && c > (uint32_t)(M3_CAP / 4) && c > pile);
        kind[i] = c == 0u ? 0u : (over ? 3u : (c > (uint32_t)(M3_CAP / 4) ? 2u : 1u));
        nord += kind[i] == 1u ? c : 0u;
        nbrk += kind[i] >= 2u ? 1u : 0u;
    }
    // exclusive prefixes over the lanes: ordinary elements, breaks
    uint32_t pord = nord, pbrk = nbrk;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, pord, o), y = __shfl_up_sync(0xffffffffu, pbrk, o);
        if (lane >= o) {
            pord += x;
            pbrk += y;
        }
    }
    pord -= nord;
    pbrk -= nbrk;
    // ordinary elements before the current run starts = the ordinary prefix at the latest break: per lane the value at
    // its last break (if any), then a "latest defined" scan over the lanes
    uint32_t run0_lane = 0xffffffffu;   // ordinary prefix at this lane's last break (none: all-ones)
    {
        uint32_t po = pord;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (kind[i] == 1u) po += v[i];
            if (kind[i] >= 2u) run0_lane = po;
        }
    }
    uint32_t run0_in = run0_lane;   // inclusive "latest defined"
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, run0_in, o);
        if (lane >= o && run0_in == 0xffffffffu) run0_in = x;
    }
    uint32_t run0 = __shfl_up_sync(0xffffffffu, run0_in, 1);   // value entering this lane
    if (lane == 0 || run0 == 0xffffffffu) run0 = 0u;
    // group key of every ordinary bucket: (breaks before it, (Pi - 1) / GRP)
    uint32_t key[8];
    uint32_t last_key = 0xffffffffu;   // key of this lane's last ordinary bucket
    {
        uint32_t po = pord, pb = pbrk, r0 = run0;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            key[i] = 0xffffffffu;
            if (kind[i] == 1u) {
                po += v[i];
                key[i] = (pb << 20) | ((po - r0 - 1u) / GRP);   // <= 256 breaks, < 2^20 groups (n < 2^31)
                last_key = key[i];
            }
            if (kind[i] >= 2u) {
                pb++;
                r0 = po;
            }
        }
    }
    // key of the ordinary bucket before / after this lane's buckets
    uint32_t prev_in = last_key;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, prev_in, o);
        if (lane >= o && prev_in == 0xffffffffu) prev_in = x;
    }
    uint32_t prev_key = __shfl_up_sync(0xffffffffu, prev_in, 1);
    if (lane == 0) prev_key = 0xffffffffu;
    uint32_t first_key = 0xffffffffu;   // key of this lane's first ordinary bucket
#pragma unroll
    for (int i = 7; i >= 0; i--)
        if (kind[i] == 1u) first_key = key[i];
    uint32_t next_in = first_key;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_down_sync(0xffffffffu, next_in, o);
        if (lane + o < 32 && next_in == 0xffffffffu) next_in = x;
    }
    uint32_t next_key = __shfl_down_sync(0xffffffffu, next_in, 1);
    if (lane == 31) next_key = 0xffffffffu;
    // heads (first bucket of a group) and tails (last bucket); a tail emits the batch, so it needs its group's head
    uint32_t head_mask = 0, tail_mask = 0;
    {
        uint32_t pk = prev_key;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (kind[i] == 1u) {
                if (key[i] != pk) head_mask |= 1u << i;
                pk = key[i];
            }
        }
        uint32_t nk = next_key;
#pragma unroll
        for (int i = 7; i >= 0; i--) {
            if (kind[i] == 1u) {
                if (key[i] != nk) tail_mask |= 1u << i;
                nk = key[i];
            }
        }
    }
    // the head of the group that is open when this lane starts: bucket index of the latest head in the lanes before
    uint32_t lane_head = 0xffffffffu;
#pragma unroll
    for (int i = 0; i < 8; i++)
        if (head_mask & (1u << i)) lane_head = (uint32_t)(lane * 8 + i);
    uint32_t head_in = lane_head;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, head_in, o);
        if (lane >= o && head_in == 0xffffffffu) head_in = x;
    }
    uint32_t open_head = __shfl_up_sync(0xffffffffu, head_in, 1);
    if (lane == 0) open_head = 0xffffffffu;
    // batches this lane emits (tails + dense buckets), ranges it queues
    uint32_t n_emit = 0, n_push = 0, nt_lane = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        n_emit += ((tail_mask >> i) & 1u) + (kind[i] == 2u ? 1u : 0u);
        if (kind[i] == 3u) {
            n_push++;
            nt_lane += (v[i] + SS_TILE - 1) / SS_TILE;
        }
    }
    uint32_t e_inc = n_emit, p_inc = n_push, t_inc = nt_lane;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t x = __shfl_up_sync(0xffffffffu, e_inc, o), y = __shfl_up_sync(0xffffffffu, p_inc, o);
        const uint32_t z = __shfl_up_sync(0xffffffffu, t_inc, o);
        if (lane >= o) {
            e_inc += x;
            p_inc += y;
            t_inc += z;
        }
    }
    const uint32_t e_tot = __shfl_sync(0xffffffffu, e_inc, 31), p_tot = __shfl_sync(0xffffffffu, p_inc, 31);
    const uint32_t t_tot = __shfl_sync(0xffffffffu, t_inc, 31);
    const bool final_digit = R.shift == 0;   // by-value mode only (the stable mode returned above)
    int32_t e_base = 0, p_base = 0, t_base = 0;
    if (lane == 0) {
        // by-value mode, final digit: queued buckets are > M3_CAP EQUAL keys, regenerated by the finish kernel
        const uint32_t nb = e_tot + (final_digit ? p_tot : 0u);
        if (nb) e_base = atomicAdd(&a.m3.cnt->n_batches[list], (int32_t)nb);
        if (p_tot && !final_digit) {
            p_base = atomicAdd(&a.m3.cnt->n_rng[round + 1], (int32_t)p_tot);
            t_base = atomicAdd(&a.m3.cnt->n_tiles[round + 1], (int32_t)t_tot);
        }
    }
    e_base = __shfl_sync(0xffffffffu, e_base, 0);
    p_base = __shfl_sync(0xffffffffu, p_base, 0);
    t_base = __shfl_sync(0xffffffffu, t_base, 0);
    {
        int32_t ei = e_base + (int32_t)(e_inc - n_emit);
        uint32_t cur_head = open_head;
        uint32_t exi = ex - t;   // exclusive prefix of this lane's first bucket (ex was advanced past the lane above)
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const uint32_t b = (uint32_t)(lane * 8 + i);
            if (head_mask & (1u << i)) cur_head = b;
            if (tail_mask & (1u << i)) {
                const uint32_t gstart = s_ex[w][cur_head];
                m3_emit_batch(a, list, ei++, R.start + gstart, (int32_t)(exi + v[i] - gstart),
                              R.klo + (cur_head << R.shift), (uint64_t)(b - cur_head + 1u) << R.shift, dst_level, 0);
            }
            if (kind[i] == 2u)
                m3_emit_batch(a, list, ei++, R.start + exi, (int32_t)v[i], R.klo + (b << R.shift), (uint64_t)1 << R.shift,
                              dst_level, 0);
            exi += v[i];
        }
    }
    if (p_tot == 0u) return;
    if (final_digit) {
        int32_t bi = e_base + (int32_t)e_tot + (int32_t)(p_inc - n_push);
        uint32_t exi = ex - t;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (kind[i] == 3u) {
                if (bi >= a.m3.batch_max) {
                    atomicMax(a.err, SS_ERR_INTERNAL);
                } else {
                    M3Batch B;
                    B.start = R.start + exi;
                    B.count = (int32_t)v[i];
                    B.klo = R.klo + (uint32_t)(lane * 8 + i);
                    B.sh = R.seg;
                    B.nslots = 1;
                    B.level = dst_level;
                    B.flags = M3_FLAG_EQUAL;
                    a.m3.batch[list][bi] = B;
                }
                bi++;
            }
            exi += v[i];
        }
        return;
    }
    if ((int64_t)p_base + p_tot > a.m3.rng_max || (int64_t)t_base + t_tot > a.m3.tiles_max) {
        atomicMax(a.err, SS_ERR_INTERNAL);
        return;
    }
    {
        int32_t pi = p_base + (int32_t)(p_inc - n_push), ti = t_base + (int32_t)(t_inc - nt_lane);
        uint32_t exi = ex - t;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (kind[i] == 3u) {
                M3Range N;
                N.start = R.start + exi;
                N.size = (int32_t)v[i];
                N.tile_base = ti;
                N.klo = R.klo + ((uint32_t)(lane * 8 + i) << R.shift);
                N.shift = R.shift > 8 ? R.shift - 8 : 0;
                N.seg = R.seg;
                N.pad = 0;
                a.m3.rng[round + 1][pi] = N;
                s_push[w][pi - p_base] = (uint8_t)(lane * 8 + i);   // <= 256 queued buckets per range
                s_ptile[w][pi - p_base] = ti;
                pi++;
                ti += (int32_t)((v[i] + SS_TILE - 1) / SS_TILE);
            }
            exi += v[i];
        }
    }
    __syncwarp();
    // tile -> range entries and zeroed digit counts of the queued ranges, by the whole warp
    for (uint32_t p = 0; p < p_tot; p++) {
        const int b = s_push[w][p];
        const int32_t idx = p_base + (int32_t)p, tb = s_ptile[w][p];
        const int32_t nt = (int32_t)((s_cnt[w][b] + SS_TILE - 1) / SS_TILE);
        const int64_t rs = R.start + s_ex[w][b];
        const int32_t rz = (int32_t)s_cnt[w][b];
        for (int32_t i = lane; i < nt; i += 32) {
            M3Tile T;
            T.t0 = rs + (int64_t)i * SS_TILE;
            T.cnt = rz - i * SS_TILE < SS_TILE ? rz - i * SS_TILE : SS_TILE;
            T.rng = idx;
            a.m3.tile_rng[round + 1][tb + i] = T;
        }
        uint32_t *h2 = a.m3.hist[round + 1] + (size_t)idx * 256;
        for (int i = lane; i < 256; i += 32) h2[i] = 0u;
    }
}

// ---- partition pass: the onesweep pass of tdt_segsort.cuh, digit = ((key - klo) >> shift) of the tile's range -----
// STABLE: the warp-private stable ranking of tdt_segsort.cuh (sorts with explicit values: equal keys must keep their
// order).  !STABLE (by-value mode): the order inside a bucket is irrelevant -- the finish kernel orders ties by value --
// so an element's place among the tile's equal digits is simply what ONE shared-memory atomicAdd on the digit's counter
// returns: ~6 instead of ~33 instructions per element for the ranking, no per-warp counter arrays (16 KB less shared
// memory per CTA).
// (Measured and dropped: a second instantiation of the tile body for FULL tiles, without the sixteen bound checks of the
//  load / count / scatter phases -- the unrolled code spills at the 64-register cap that four CTAs per SM need: 95.6 vs
//  89 us; three CTAs per SM at 85 registers: no better.)
template <bool STABLE>
__global__ void __launch_bounds__(SS_THREADS, STABLE ? TDT_SS_PASS_MINBLOCKS : 4) m3_pass_kernel(SSArgs a, int round, int dst_level) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    unsigned char *p = ss_smem;
    uint2 *KV = (uint2 *)p; p += SS_TILE * 8;
    uint32_t(*wh)[256] = (uint32_t(*)[256])p; p += STABLE ? SS_WARPS * 256 * 4 : 256 * 4;   // !STABLE: one counter row
    uint32_t(*mm)[256] = (uint32_t(*)[256])p; p += STABLE ? SS_MM * SS_WARPS * 256 * 4 : 0;
    uint32_t *bin = (uint32_t *)p; p += 256 * 4;
    int32_t *gbase = (int32_t *)p;   // destination of a digit's first element, relative to the range (32-bit offsets:
                                     // the 64-bit form cost 8 more instructions per element in the write loop)

    const int n_tiles = a.m3.cnt->n_tiles[round];
    const uint32_t *src_k, *dk_c;
    const int32_t *src_v, *dv_c;
    m3_level_ptrs(a, round, src_k, src_v);
    m3_level_ptrs(a, dst_level, dk_c, dv_c);
    uint32_t *dst_k = (uint32_t *)dk_c;
    int32_t *dst_v = (int32_t *)dv_c;
    const uint32_t epoch = (uint32_t)round + 1u;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        if (STABLE) {
            for (int i = threadIdx.x; i < SS_WARPS * 256; i += SS_THREADS) {
                (&wh[0][0])[i] = 0u;
                if (SS_MM) (&mm[0][0])[i] = 0u;
            }
        } else {
            if (threadIdx.x < 256) wh[0][threadIdx.x] = 0u;
        }
        const M3Tile T = a.m3.tile_rng[round][tile];   // the key loads below depend on this record alone
        const int r = T.rng;
        const int64_t t0 = T.t0;
        const int cnt = T.cnt;
        const M3Range R = a.m3.rng[round][r];
        const int lt = tile - R.tile_base;
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        constexpr int epw = SS_TILE / SS_WARPS;
        const uint32_t klo = R.klo;
        const int shift = R.shift;
        auto dig = [&](uint32_t key) -> uint32_t {
            const uint32_t d = (key - klo) >> shift;
            return d > 255u ? 255u : d;
        };

        uint32_t key[SS_CHUNKS];
        int32_t val[SS_CHUNKS];
#pragma unroll
        for (int c = 0; c < SS_CHUNKS; c++) {
            const int e = warp * epw + c * 32 + lane;
            key[c] = 0u;
            val[c] = 0;
            if (e < cnt) {
                key[c] = src_k[t0 + e];
                val[c] = src_v ? src_v[t0 + e] : (int32_t)(t0 + e);
            }
        }
        __syncthreads();
        uint32_t info[SS_CHUNKS];
        uint32_t total, excl;
        if (STABLE) {
            ss_count<SS_CHUNKS>(cnt, epw, wh[warp], mm[warp], info, 8, [&](int, int c) -> uint32_t { return dig(key[c]); });
            __syncthreads();
            ss_digit_bases<SS_WARPS>(wh, bin, total, excl);
        } else {
#pragma unroll
            for (int c = 0; c < SS_CHUNKS; c++) {
                const int e = warp * epw + c * 32 + lane;
                info[c] = 0u;
                if (e < cnt) {
                    const uint32_t d = dig(key[c]);
                    info[c] = (d << 16) | atomicAdd(&wh[0][d], 1u);   // digit : place among the tile's equal digits
                }
            }
            __syncthreads();
            total = threadIdx.x < 256 ? wh[0][threadIdx.x] : 0u;
            excl = ss_block_excl_scan_256<SS_WARPS>(total, bin);
            if (threadIdx.x < 256) wh[0][threadIdx.x] = excl;
        }
        uint32_t *row = a.L.status + (size_t)tile * 256 + threadIdx.x;
        const bool live = threadIdx.x < 256;
        uint32_t gh = 0;
        if (live) {
            st_volatile_u32(row, ss_pack(lt == 0 ? 2u : 1u, epoch, total));
            gh = a.m3.hist[round][(size_t)r * 256 + threadIdx.x];
        }
        __syncthreads();
        if (STABLE) {
            ss_scatter<SS_CHUNKS>(epw, wh[warp], info, [&](int, int c) -> uint32_t { return dig(key[c]); },
                                  [&](int, int c, uint32_t pos) { KV[pos] = make_uint2(key[c], (uint32_t)val[c]); });
        } else {
#pragma unroll
            for (int c = 0; c < SS_CHUNKS; c++) {
                const int e = warp * epw + c * 32 + lane;
                if (e < cnt) KV[wh[0][info[c] >> 16] + (info[c] & 0xffffu)] = make_uint2(key[c], (uint32_t)val[c]);
            }
        }
        {
            uint32_t before = 0;
            if (lt != 0 && live) {
                constexpr int LB = TDT_SS_LB;
                const uint32_t *prow = row - 256;
                int left = lt;
                bool done = false;
                while (!done) {
                    uint32_t sv[LB];
#pragma unroll
                    for (int i = 0; i < LB; i++) sv[i] = i < left ? ld_volatile_u32(prow - (size_t)i * 256) : 0u;
#pragma unroll
                    for (int i = 0; i < LB; i++) {
                        if (!done && i < left) {
                            uint32_t s = sv[i];
                            while ((s >> 30) == 0u || ((s >> 26) & 15u) != epoch) s = ld_volatile_u32(prow - (size_t)i * 256);
                            before += s & 0x3ffffffu;
                            done = (s >> 30) == 2u;
                        }
                    }
                    prow -= (size_t)LB * 256;
                    left -= LB;
                }
                st_volatile_u32(row, ss_pack(2u, epoch, before + total));
            }
            if (threadIdx.x < 256) gbase[threadIdx.x] = (int32_t)(gh + before - excl);
        }
        __syncthreads();
        uint32_t *dk = dst_k + R.start;
        int32_t *dv = dst_v + R.start;
        for (int i = threadIdx.x; i < cnt; i += SS_THREADS) {
            const uint2 kv = KV[i];
            const int32_t g = gbase[dig(kv.x)] + i;
            dk[g] = kv.x;
            dv[g] = (int32_t)kv.y;
        }
        __syncthreads();
    }
}

// Members of slot [lo, hi) that sort before `me`; called by all 32 lanes (lanes without an element pass lo == hi == 0).
// The staged words are in slot order, so whatever follows a slot is LARGER than its members, and the M3_PAD words behind
// the last element are all-ones: the loop can run past hi without a bound check, to the largest slot among the warp's
// lanes (one REDUX, no divergent exits) -- four instructions per step.  (The generic per-lane loop the compiler built for
// `for j in [lo, hi)`, unrolled 16 / 8 / 4 / 2 / 1 with a reconvergence point each, was 68 % of the kernel's instructions.)
__device__ __forceinline__ uint32_t m3_rank_in_slot(const u64 *KV, uint32_t lo, uint32_t hi, u64 me) {
    uint32_t cmax = __reduce_max_sync(0xffffffffu, hi - lo);
    uint32_t rank = 0;
    if (cmax > 1u) {
        if (cmax > (uint32_t)M3_PAD) {   // a crowded slot in a batch that is not crowded overall: bounded loop
            for (uint32_t j = lo; j < hi; j++) rank += KV[j] < me ? 1u : 0u;
        } else {
            const u64 *p = KV + lo;
#pragma unroll 4
            for (uint32_t q = 0; q < cmax; q++) rank += p[q] < me ? 1u : 0u;
        }
    }
    return rank;
}

#ifndef TDT_M3_HOT
#define TDT_M3_HOT 32   // a batch goes through the bitonic network when sum(slot count ^ 2) > TDT_M3_HOT x elements
#endif

// ---- finish: one ranking round in shared memory ---------------------------------------------------------------
// BYVAL: values grow with the position inside every segment (the value is the tie-break); otherwise the tie-break is
// the position in the batch and every thread writes its own elements to their final places.
#ifdef TDT_M3_DEBUG
__device__ unsigned long long g_m3_dbg[16];
#define M3_DBG(stmt) do { if (threadIdx.x == 0) { stmt; } } while (0)
#else
#define M3_DBG(stmt) do { } while (0)
#endif

template <bool BYVAL>
__global__ void __launch_bounds__(M3_THREADS, 2) m3_finish_kernel(SSArgs a, int list) {
    extern __shared__ __align__(16) unsigned char ss_smem[];
    u64 *KV = (u64 *)ss_smem;
    uint32_t *cnt = (uint32_t *)(ss_smem + (size_t)(M3_CAP + M3_PAD) * 8);
    uint32_t *ws = cnt + M3_NSLOT;   // [0..15] warp sums, [32] sum of squared slot counts
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    int nb = a.m3.cnt->n_batches[list];
    if ((int64_t)nb > a.m3.batch_max) nb = (int)a.m3.batch_max;
    for (int b = blockIdx.x; b < nb; b += gridDim.x) {
        const M3Batch B = a.m3.batch[list][b];
#ifdef TDT_M3_DEBUG
        const long long dbg_t0 = clock64();
#endif
        const uint32_t *sk;
        const int32_t *sv;
        m3_level_ptrs(a, B.level, sk, sv);
        const int64_t g0 = B.start;
        const int count = B.count;
        if (B.flags & M3_FLAG_COPY) {   // uniform over the CTA; a copy never has src == dst (odd levels only)
            for (int i = t; i < count; i += M3_THREADS) {
                a.keys_out[g0 + i] = sk[g0 + i];
                a.vals_out[g0 + i] = sv[g0 + i];
            }
            continue;
        }
        if (B.flags & M3_FLAG_EQUAL) {
            // > M3_CAP equal keys (by-value mode): their values in order are the positions of the key in the input
            // segment, regenerated by a stable compaction of the segment (uniform over the CTA, rare)
            const int64_t s0 = a.off[B.sh], s1 = a.off[B.sh + 1];
            uint32_t run = 0;
            for (int64_t base = s0; base < s1; base += M3_THREADS) {
                const int64_t j = base + t;
                const bool hit = j < s1 && a.keys_in[j] == B.klo;
                const uint32_t bal = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) ws[warp] = __popc(bal);
                __syncthreads();
                uint32_t before = run, all = 0;
#pragma unroll
                for (int w2 = 0; w2 < M3_THREADS / 32; w2++) {
                    const uint32_t x = ws[w2];
                    if (w2 < warp) before += x;
                    all += x;
                }
                if (hit) {
                    const uint32_t pos = before + __popc(bal & lanemask_lt());
                    if (pos < (uint32_t)count) {
                        a.keys_out[g0 + pos] = B.klo;
                        a.vals_out[g0 + pos] = (int32_t)j;
                    }
                }
                run += all;
                __syncthreads();
            }
            continue;
        }
        const int nslots = B.nslots, sh = B.sh;
        const uint32_t klo = B.klo, smax = (uint32_t)nslots - 1u;
        const int nz = (nslots + 15) & ~15;
        for (int i = t * 4; i < nz; i += M3_THREADS * 4) *(uint4 *)(cnt + i) = make_uint4(0u, 0u, 0u, 0u);
        if (t == 0) ws[32] = 0u;
        uint32_t key[M3_EPT];
        int32_t val[M3_EPT];
#pragma unroll
        for (int c = 0; c < M3_EPT; c++) {
            key[c] = 0u;
            val[c] = 0;
        }
#pragma unroll
        for (int c = 0; c < M3_EPT; c++) {
            if (c * M3_THREADS >= count) break;   // uniform: a batch rarely fills all sixteen chunks
            const int e = c * M3_THREADS + t;
            if (e < count) {
                key[c] = sk[g0 + e];
                val[c] = sv ? sv[g0 + e] : (int32_t)(g0 + e);
            }
        }
        if (B.level == 0 && a.key_bits < 32) {
            bool bad = false;
#pragma unroll
            for (int c = 0; c < M3_EPT; c++) bad = bad || (key[c] >> a.key_bits);
            if (bad) atomicMax(a.err, SS_ERR_KEY_RANGE);
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < M3_EPT; c++) {
            if (c * M3_THREADS >= count) break;
            const int e = c * M3_THREADS + t;
            if (e < count) {
                uint32_t s = (key[c] - klo) >> sh;
                s = s > smax ? smax : s;
                atomicAdd(&cnt[s], 1u);
            }
        }
        __syncthreads();
        // block scan over the slot counts: thread t owns slots [16 t, 16 t + 16)
        uint32_t sum = 0, sq = 0;
        const bool own = t * 16 < nz;
        if (own) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const uint4 x = *(const uint4 *)(cnt + t * 16 + q * 4);
                sum += x.x + x.y + x.z + x.w;
                sq += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w;
            }
        }
        uint32_t inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        sq = warp_sum(sq);
        if (lane == 31) ws[warp] = inc;
        if (lane == 0 && sq) atomicAdd(&ws[32], sq);
        __syncthreads();
        uint32_t run = inc - sum;
#pragma unroll
        for (int w2 = 0; w2 < M3_THREADS / 32; w2++)
            if (w2 < warp) run += ws[w2];
        const bool hot = ws[32] > (uint32_t)TDT_M3_HOT * (uint32_t)count;
        if (own) {
#pragma unroll
            for (int q = 0; q < 4; q++) {
                uint4 x = *(const uint4 *)(cnt + t * 16 + q * 4);
                uint4 y;
                y.x = run; run += x.x;
                y.y = run; run += x.y;
                y.z = run; run += x.z;
                y.w = run; run += x.w;
                *(uint4 *)(cnt + t * 16 + q * 4) = y;
            }
        }
        __syncthreads();
        if (!hot) {
            // place: after this loop cnt[s] is the END of slot s (= the start of slot s + 1)
#pragma unroll
            for (int c = 0; c < M3_EPT; c++) {
                if (c * M3_THREADS >= count) break;
                const int e = c * M3_THREADS + t;
                if (e < count) {
                    uint32_t s = (key[c] - klo) >> sh;
                    s = s > smax ? smax : s;
                    const uint32_t pos = atomicAdd(&cnt[s], 1u);
                    KV[pos] = ((u64)key[c] << 32) | (u64)(uint32_t)(BYVAL ? val[c] : e);
                }
            }
            if (t < M3_PAD) KV[count + t] = ~0ull;
            __syncthreads();
            if (BYVAL) {
                uint32_t *ok = a.keys_out + g0;
                int32_t *ov = a.vals_out + g0;
                for (int base = 0; base < count; base += M3_THREADS) {   // uniform trip count: the ranking is warp-wide
                    const int i = base + t;
                    const bool live = i < count;
                    const u64 me = live ? KV[i] : 0ull;
                    const uint32_t k = (uint32_t)(me >> 32);
                    uint32_t s = (k - klo) >> sh;
                    s = s > smax ? smax : s;
                    uint32_t hi = 0, lo = 0;
                    if (live) {
                        hi = cnt[s];
                        lo = s ? cnt[s - 1] : 0u;
                    }
                    const uint32_t rank = m3_rank_in_slot(KV, lo, hi, me);
                    if (live) {
                        const uint32_t g = lo + rank;
                        ok[g] = k;
                        ov[g] = (int32_t)(uint32_t)me;
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < M3_EPT; c++) {
                    if (c * M3_THREADS >= count) break;
                    const int e = c * M3_THREADS + t;
                    const bool live = e < count;
                    const u64 me = ((u64)key[c] << 32) | (u64)(uint32_t)e;
                    uint32_t s = (key[c] - klo) >> sh;
                    s = s > smax ? smax : s;
                    uint32_t hi = 0, lo = 0;
                    if (live) {
                        hi = cnt[s];
                        lo = s ? cnt[s - 1] : 0u;
                    }
                    const uint32_t rank = m3_rank_in_slot(KV, lo, hi, me);
                    if (live) {
                        const int64_t g = g0 + lo + rank;
                        a.keys_out[g] = key[c];
                        a.vals_out[g] = val[c];
                    }
                }
            }
        } else {
            // Crowded slots: equal-width slots do not fit this batch (clusters of a sparse pair: the batch spans
            // megabases, a slot kilobases, and every cluster sits in one slot).  EQUI-DEPTH slots do: every 8th element
            // (in source order) is a sample, the sorted samples are the splitters, an element's slot is the number of
            // splitters <= its (key, tie-break) word -- about eight elements per slot whatever the distribution -- and
            // the rest is the same count / scan / place / rank-in-slot sequence.  Shared memory: the 32 KB of the slot
            // counters hold [0, 4 KB) the 1024 new counters, [8, 16 KB) the splitters, [16, 32 KB) the slot of every
            // placed element (16 bits each).  Only heavy exact duplicates crowd equi-depth slots; they go to the
            // bitonic network below.
            u64 *SP = (u64 *)(cnt + 2048);
            uint16_t *SL = (uint16_t *)(cnt + 4096);
            auto word = [&](int c, int e) -> u64 { return ((u64)key[c] << 32) | (u64)(uint32_t)(BYVAL ? val[c] : e); };
            auto slot_of = [&](u64 x) -> uint32_t {   // number of splitters <= x, capped at 1023
                uint32_t pos = 0;
#pragma unroll
                for (uint32_t step = 512; step > 0; step >>= 1)
                    if (SP[pos + step - 1] <= x) pos += step;
                return pos;
            };
            for (int i = t; i < 1024; i += M3_THREADS) cnt[i] = 0u;
            if ((t & 7) == 3) {   // e = c * 512 + t: e mod 8 == 3 for every c
#pragma unroll
                for (int c = 0; c < M3_EPT; c++) {
                    const int e = c * M3_THREADS + t;
                    SP[c * (M3_THREADS / 8) + (t >> 3)] = e < count ? word(c, e) : ~0ull;
                }
            }
            if (t == 0) ws[33] = 0u;
            __syncthreads();
            for (int k = 2; k <= 1024; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1)), hi = lo | j;   // 512 threads, 512 pairs
                    const u64 x = SP[lo], y = SP[hi];
                    if ((x > y) == ((lo & k) == 0)) {
                        SP[lo] = y;
                        SP[hi] = x;
                    }
                    __syncthreads();
                }
            }
#pragma unroll
            for (int c = 0; c < M3_EPT; c++) {
                if (c * M3_THREADS >= count) break;
                const int e = c * M3_THREADS + t;
                if (e < count) atomicAdd(&cnt[slot_of(word(c, e))], 1u);
            }
            __syncthreads();
            {   // scan of the 1024 counters: two per thread
                const uint32_t c0 = cnt[2 * t], c1 = cnt[2 * t + 1];
                uint32_t inc2 = c0 + c1;
                const uint32_t mine = inc2;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const uint32_t u = __shfl_up_sync(0xffffffffu, inc2, o);
                    if (lane >= o) inc2 += u;
                }
                if (lane == 31) ws[warp] = inc2;
                const uint32_t big = __reduce_max_sync(0xffffffffu, c0 > c1 ? c0 : c1);
                if (lane == 0 && big > (uint32_t)M3_PAD) atomicMax(&ws[33], big);
                __syncthreads();
                uint32_t run2 = inc2 - mine;
#pragma unroll
                for (int w2 = 0; w2 < M3_THREADS / 32; w2++)
                    if (w2 < warp) run2 += ws[w2];
                cnt[2 * t] = run2;
                cnt[2 * t + 1] = run2 + c0;
            }
            __syncthreads();
            const bool dup = ws[33] != 0u;   // a slot beyond the padded reach of the ranking loop: heavy duplicates
            if (!dup) {
#pragma unroll
                for (int c = 0; c < M3_EPT; c++) {
                    if (c * M3_THREADS >= count) break;
                    const int e = c * M3_THREADS + t;
                    if (e < count) {
                        const u64 x = word(c, e);
                        const uint32_t sl = slot_of(x);
                        const uint32_t pos = atomicAdd(&cnt[sl], 1u);   // afterwards cnt[sl] = END of slot sl
                        KV[pos] = x;
                        SL[pos] = (uint16_t)sl;
                    }
                }
                if (t < M3_PAD) KV[count + t] = ~0ull;
                __syncthreads();
                if (BYVAL) {
                    uint32_t *ok = a.keys_out + g0;
                    int32_t *ov = a.vals_out + g0;
                    for (int base = 0; base < count; base += M3_THREADS) {
                        const int i = base + t;
                        const bool live = i < count;
                        const u64 me = live ? KV[i] : 0ull;
                        uint32_t hi = 0, lo = 0;
                        if (live) {
                            const uint32_t sl = SL[i];
                            hi = cnt[sl];
                            lo = sl ? cnt[sl - 1] : 0u;
                        }
                        const uint32_t rank = m3_rank_in_slot(KV, lo, hi, me);
                        if (live) {
                            ok[lo + rank] = (uint32_t)(me >> 32);
                            ov[lo + rank] = (int32_t)(uint32_t)me;
                        }
                    }
                } else {
#pragma unroll
                    for (int c = 0; c < M3_EPT; c++) {
                        if (c * M3_THREADS >= count) break;
                        const int e = c * M3_THREADS + t;
                        const bool live = e < count;
                        const u64 me = word(c, e);
                        uint32_t hi = 0, lo = 0;
                        if (live) {
                            const uint32_t sl = slot_of(me);
                            hi = cnt[sl];
                            lo = sl ? cnt[sl - 1] : 0u;
                        }
                        const uint32_t rank = m3_rank_in_slot(KV, lo, hi, me);
                        if (live) {
                            const int64_t g = g0 + lo + rank;
                            a.keys_out[g] = key[c];
                            a.vals_out[g] = val[c];
                        }
                    }
                }
            } else {
            // heavy duplicates: bitonic network over the padded batch, same 64-bit (key, tie-break) words
            int P = 2;
            while (P < count) P <<= 1;
#pragma unroll
            for (int c = 0; c < M3_EPT; c++) {
                const int e = c * M3_THREADS + t;
                if (e < P) KV[e] = e < count ? (((u64)key[c] << 32) | (u64)(uint32_t)(BYVAL ? val[c] : e)) : ~0ull;
            }
            __syncthreads();
            for (int k = 2; k <= P; k <<= 1) {
                for (int j = k >> 1; j > 0; j >>= 1) {
                    for (int i = t; i < (P >> 1); i += M3_THREADS) {
                        const int lo = ((i & ~(j - 1)) << 1) | (i & (j - 1)), hi = lo | j;
                        const u64 x = KV[lo], y = KV[hi];
                        const bool asc = (lo & k) == 0;
                        if ((x > y) == asc) {
                            KV[lo] = y;
                            KV[hi] = x;
                        }
                    }
                    __syncthreads();
                }
            }
            if (BYVAL) {
                for (int i = t; i < count; i += M3_THREADS) {
                    const u64 me = KV[i];
                    a.keys_out[g0 + i] = (uint32_t)(me >> 32);
                    a.vals_out[g0 + i] = (int32_t)(uint32_t)me;
                }
            } else {
                for (int i = t; i < count; i += M3_THREADS) cnt[(uint32_t)KV[i]] = (uint32_t)i;   // final place of element e
                __syncthreads();
#pragma unroll
                for (int c = 0; c < M3_EPT; c++) {
                    if (c * M3_THREADS >= count) break;
                    const int e = c * M3_THREADS + t;
                    if (e < count) {
                        const int64_t g = g0 + cnt[e];
                        a.keys_out[g] = key[c];
                        a.vals_out[g] = val[c];
                    }
                }
            }
            }
        }
        __syncthreads();   // KV / cnt / ws are rewritten by the next batch
        M3_DBG({
            const unsigned long long dt = (unsigned long long)(clock64() - dbg_t0);
            atomicAdd(&g_m3_dbg[0], 1ull);
            atomicAdd(&g_m3_dbg[hot ? 1 : 2], 1ull);
            atomicAdd(&g_m3_dbg[hot ? 3 : 4], dt);
            atomicMax(&g_m3_dbg[5], dt);
            atomicAdd(&g_m3_dbg[6], (unsigned long long)count);
            atomicAdd(&g_m3_dbg[7], (unsigned long long)ws[32]);
            atomicAdd(&g_m3_dbg[8 + (B.level < 5 ? B.level : 5)], 1ull);
        });
    }
}

}  // namespace tdt

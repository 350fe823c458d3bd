/*
 * tdt_b200.h -- C ABI of libtdt_b200.so, the B200 (sm_100a) implementation of TIDDIT's
 * signal-clustering / coverage / GC hot path.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless its name ends in _h (host);
 *   - the caller owns every buffer; the library allocates nothing persistent: scratch comes from
 *     the caller-provided workspace `ws` (size it with the matching *_workspace_bytes call);
 *   - work is enqueued on `stream` (a cudaStream_t passed as void*); functions documented as
 *     "synchronises" block on that stream once, everything else returns as soon as it is enqueued;
 *   - return value: 0 = ok, < 0 = error (TDT_E_*), text via tdt_last_error() (thread-local);
 *   - no exceptions cross the boundary; distinct streams may be driven from distinct threads.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the upstream
 * TIDDIT tree, v3.9.5).
 */
#ifndef TDT_B200_H
#define TDT_B200_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define TDT_API __attribute__((visibility("default")))
#else
#define TDT_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define TDT_OK 0
#define TDT_E_ARG (-1)       /* invalid argument (min_pts < 2, bin_size <= 0, negative sizes ...)   */
#define TDT_E_WORKSPACE (-2) /* workspace missing or too small                                      */
#define TDT_E_CUDA (-3)      /* a CUDA runtime call failed                                          */
#define TDT_E_RANGE (-4)     /* a coordinate is outside the range the call was told to expect       */

TDT_API int tdt_version(void);
TDT_API const char *tdt_last_error(void);

/* --------------------------------------------------------------------------------------------
 * Clustering.  Replaces tiddit/tiddit_cluster.pyx:140-160 (per (chrA,chrB): stable sort by posA,
 * DBSCAN.main, labels back in insertion order) for ALL pairs in one call, and therefore
 * tiddit/DBSCAN.py:125-129 (main), :33-64 (x_coordinate_clustering), :66-123
 * (y_coordinate_clustering).
 *
 * Signals: posA[i], posB[i] >= 0 (int32), i in [0, n).  Pairs are segments: seg_off[P+1] (device, int64),
 * the signals of pair p are [seg_off[p], seg_off[p+1]) in insertion order (= the order of the reference's
 * global signal index).
 * labels_out[i] (int32, input order): the reference's own ids -- x-pass ids 0..nx-1 in order of run
 * start, y-pass sub-clusters nx, nx+1, ... in visiting order, -1 = noise -- per pair, so they are
 * bit-identical to int(DBSCAN.main(...)[i]) (ids are < the pair's signal count).
 * max_pos: an upper bound on every posA/posB (e.g. the longest contig); 0 = unknown (31 bits assumed).
 *          It only sizes the radix-sort keys; a coordinate with bits above it returns TDT_E_RANGE.
 * Everything is enqueued without host round trips (data-dependent sizes stay on the device); the call
 * synchronises `stream` once at the end to read the status word.
 * ------------------------------------------------------------------------------------------ */
TDT_API size_t tdt_cluster_workspace_bytes(int64_t n, int32_t P);

TDT_API int tdt_cluster_labels(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, int64_t n, int32_t P,
                       int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_out, void *ws,
                       size_t ws_bytes, void *stream);

/* The same without the final host synchronisation, for callers that pipeline several calls (chunks of pairs)
 * behind host<->device copies: nothing blocks; a data error (TDT_E_RANGE conditions) is max-reduced as a
 * positive code into *status_accum (device int32, zeroed by the caller: 1 = posA out of range, 2 = posB out of
 * range), to be read after the caller's own synchronisation.  The workspace may be reused by the next call on the
 * same stream. */
TDT_API int tdt_cluster_labels_async(const int32_t *posA, const int32_t *posB, const int64_t *seg_off, int64_t n,
                                     int32_t P, int32_t eps, int32_t min_pts, int32_t max_pos, int32_t *labels_out,
                                     void *ws, size_t ws_bytes, int32_t *status_accum, void *stream);

/* DBSCAN.py:125-129 on ONE array in the caller's order (the reference does not sort inside
 * DBSCAN.main; its caller does): x-pass over x[] as given -- window max of |x[j]-x[i]|, so unsorted
 * input behaves like the reference -- then the y-pass.  Same workspace as tdt_cluster_labels(n, 1).
 * Synchronises once. */
TDT_API int tdt_dbscan_main(const int32_t *x, const int32_t *y, int64_t n, int32_t eps, int32_t min_pts, int32_t max_pos,
                    int32_t *labels_out, void *ws, size_t ws_bytes, void *stream);

/* DBSCAN.py:33-64 (x_coordinate_clustering; second caller tiddit_contig_analysis.pyx:176):
 * labels_out[i] = run id or -1; *last_id_out (device int32) = the returned cluster_id (-1 = none).
 * Does not synchronise. */
TDT_API int tdt_xpass_labels(const int32_t *x, int64_t n, int32_t eps, int32_t min_pts, int32_t *labels_out,
                     int32_t *last_id_out, void *ws, size_t ws_bytes, void *stream);

/* DBSCAN.py:66-123 (y_coordinate_clustering) as a stand-alone call: labels_io holds the x-pass ids
 * (ids in [0, cluster_id], -1 = noise) and is rewritten in place; *cluster_id_io (device int32) is
 * advanced by the number of extra sub-clusters like the reference's return value.
 * Synchronises once. */
TDT_API int tdt_ypass_labels(const int32_t *y, int64_t n, int32_t eps, int32_t min_pts, int32_t max_pos,
                     int32_t *labels_io, int32_t *cluster_id_io, void *ws, size_t ws_bytes, void *stream);

/* --------------------------------------------------------------------------------------------
 * Candidate aggregation.  Replaces tiddit/tiddit_cluster.pyx:156-254 (labels folded into candidates, signal by
 * signal) and :258-336 (per-candidate N_discordants / N_splits / N_contigs, posA / posB, startA / endA / startB /
 * endB) for ALL pairs in one call.  Per-signal inputs are the reference's records (:72,101,134) as
 * struct-of-arrays in insertion order, pairs back to back like tdt_cluster_labels:
 *   labels   the output of tdt_cluster_labels;  posA/posB  int(rec[3]) / int(rec[5]) (0 <= pos <= max_pos < 2^30)
 *   span     [n][4] = rec[8..11] (startA, endA, startB, endB), 16-byte aligned
 *   name_id  equal ids <=> equal read names (0 <= id <= n_names < 2^30; n_names = 0: unknown)
 *   flags    TDT_SIG_* bits;  same_chrom[P]  1 if chrA == chrB
 * Noise (-1) survives only as an intra-chromosomal assembly contig with posB - posA < 2*max_ins_len, as a
 * singleton candidate with id len(pair) + k (:162-168).
 * Outputs (caller-owned, sized for the worst case n):
 *   cand_out[c][TDT_CAND_COLS]  one row per candidate, rows in the reference's dict insertion order (pair by
 *                               pair, candidates by first appearance): TDT_CAND_* columns
 *   member_idx_out[n]           signal indices grouped by candidate: the members of row c are
 *                               member_idx_out[row[MEMBER_OFF] .. + row[SIZE]) in insertion order
 *   counts_out[4] (int64)       {candidates, signals kept, data error (0 = none), 0}
 * Does not synchronise.  Workspace: tdt_aggregate_workspace_bytes(n, P).
 * ------------------------------------------------------------------------------------------ */
#define TDT_SIG_KIND_MASK 0x03 /* 0 = discordant pair ("D"), 1 = split read ("S"), 2 = assembly contig ("A") */
#define TDT_SIG_A_TRUE 0x04    /* orientation string of side A (rec[4]) == "True"  */
#define TDT_SIG_A_FALSE 0x08   /*                                        == "False" */
#define TDT_SIG_B_TRUE 0x10    /* rec[6] == "True"  */
#define TDT_SIG_B_FALSE 0x20   /* rec[6] == "False" */
#define TDT_CAND_COLS 16
enum {
    TDT_CAND_PAIR = 0, TDT_CAND_ID = 1, TDT_CAND_FIRST = 2, TDT_CAND_MEMBER_OFF = 3, TDT_CAND_SIZE = 4,
    TDT_CAND_N_DISCORDANTS = 5, TDT_CAND_N_SPLITS = 6, TDT_CAND_N_CONTIGS = 7, TDT_CAND_POSA = 8, TDT_CAND_POSB = 9,
    TDT_CAND_STARTA = 10, TDT_CAND_ENDA = 11, TDT_CAND_STARTB = 12, TDT_CAND_ENDB = 13,
    TDT_CAND_RULE = 14 /* 0 splits >= min_reads, 1 contigs, 2 splits, 3 discordants by orientation, 4 by mode */
};

TDT_API size_t tdt_aggregate_workspace_bytes(int64_t n, int32_t P);

TDT_API int tdt_cluster_aggregate(const int32_t *labels, const int32_t *posA, const int32_t *posB, const int32_t *span,
                                  const int32_t *name_id, const uint8_t *flags, const int64_t *seg_off,
                                  const uint8_t *same_chrom, int64_t n, int32_t P, int32_t max_ins_len, int32_t is_mp,
                                  int32_t min_reads, int32_t max_pos, int32_t n_names, int32_t *cand_out,
                                  int32_t *member_idx_out, int64_t *counts_out, void *ws, size_t ws_bytes,
                                  void *stream);

/* --------------------------------------------------------------------------------------------
 * Coverage.  Replaces tiddit/tiddit_coverage.pyx:48-74 (update_coverage) applied to a batch of
 * reads: bins[] (float64) is accumulated IN PLACE (the reference's `+=`).  Every addend is the
 * float32 quotient the reference computes, all partial sums are exact in float64, so the result is
 * bit-identical for any summation order.
 *   single contig : reads (start[i], end[i]) against bins[0..n_bins) with the contig's end_bin_size;
 *   all contigs   : reads grouped by contig, read_off[C+1] / bin_off[C+1] (device int64) delimit the
 *                   reads / the bins of contig c inside start/end / bins, end_bin_size[C] (device).
 * A read the reference would reject (IndexError: a touched bin outside [0, n_bins)) is skipped and
 * its index is min-reduced into *first_bad (device int64, caller initialises to -1 semantics:
 * INT64_MAX = none); negative bin indices (Python wrap-around in the reference) are rejected too.
 * Does not synchronise.
 * ------------------------------------------------------------------------------------------ */
TDT_API int tdt_coverage_accumulate(const int32_t *start, const int32_t *end, int64_t n_reads, int32_t bin_size,
                            int32_t end_bin_size, double *bins, int64_t n_bins, int64_t *first_bad, void *stream);

TDT_API int tdt_coverage_accumulate_contigs(const int32_t *start, const int32_t *end, int64_t n_reads,
                                    const int64_t *read_off, const int64_t *bin_off, const int32_t *end_bin_size,
                                    int32_t C, int32_t bin_size, double *bins, int64_t n_bins_total,
                                    int64_t *first_bad, void *stream);

/* --------------------------------------------------------------------------------------------
 * Masked coverage medians.  Replaces the bin loops of tiddit/tiddit_coverage_analysis.pyx:14-29
 * (determine_ploidy): for every contig c the numpy.median of bins[i] over bin_off[c] <= i < bin_off[c+1] with
 * bins[i] > 0 and gc[i] != -1 (:17-22), and the same over all bins (:26-27).  bins / gc are the coverage and GC
 * arrays of all contigs back to back (same bin size), bin_off[C+1] (device int64) delimits them.
 * medians_out[C+1] (float64): contigs, then the genome-wide median; NaN where nothing qualifies (numpy.median([])).
 * counts_out[C+1] (int64): qualifying bins.  Exact (radix select on the bit patterns; even counts average the two
 * middle values like numpy).  Does not synchronise.
 * ------------------------------------------------------------------------------------------ */
TDT_API size_t tdt_coverage_medians_workspace_bytes(int32_t C);

TDT_API int tdt_coverage_medians(const double *bins, const int8_t *gc, const int64_t *bin_off, int32_t C, int64_t n_bins,
                                 double *medians_out, int64_t *counts_out, void *ws, size_t ws_bytes, void *stream);

/* --------------------------------------------------------------------------------------------
 * GC bins.  Replaces tiddit/tiddit_gc.pyx:6-33 (binned_gc) on a contig sequence resident in HBM
 * (one byte per base, as pysam.FastaFile.fetch returns it): out[b] (int8) = -1 if
 * #N/bin_size > n_cutoff else rint(100 * #GC / #chars), ceil(len / bin_size) bins.
 * Does not synchronise.
 * ------------------------------------------------------------------------------------------ */
TDT_API int tdt_gc_bins(const uint8_t *seq, int64_t len, int32_t bin_size, double n_cutoff, int8_t *out, void *stream);

/* Test hook: the segmented stable radix sort on its own.  keys/vals (vals may be NULL = element index) and
 * off[nseg+1] are device arrays; every segment [off[s], off[s+1]) is sorted by its low key_bits key bits into
 * keys_out/vals_out.  segid (may be NULL): segment of every element, enables the counting path for tiny segments.  ws: at least 8*n + 1024 + the sort's temporaries (tdt_cluster_workspace_bytes(n, nseg) is
 * enough).  Synchronises. */
TDT_API int tdt_debug_segsort(const uint32_t *keys, const int32_t *vals, const int64_t *off, const int32_t *segid,
                              int64_t nseg, int64_t n, int32_t key_bits, uint32_t *keys_out, int32_t *vals_out,
                              void *ws, size_t ws_bytes, void *stream);

/* Per-stage device timing for bench.py's roofline line: between tdt_profile_begin() and tdt_profile_end()
 * every stage of every call is bracketed by CUDA events on the caller's stream; tdt_profile_end synchronises
 * on them and writes "stage=milliseconds\n" lines into out (returns the byte count).  Not thread-safe. */
TDT_API void tdt_profile_begin(void);
TDT_API int tdt_profile_end(char *out, size_t cap);

/* Launch counter: number of kernels this library has launched in the calling process (all threads),
 * for bench.py's "gpu_launches" line. */
TDT_API int64_t tdt_launch_count(void);

/* --------------------------------------------------------------------------------------------
 * Label exchange of the sharded clustering path over NVLink peer memory (csrc/tdt_peer.cu).
 * The reference clusters the (chrA,chrB) pairs serially in one process (tiddit/tiddit_cluster.pyx:140-154); pairs
 * are independent, so N ranks label disjoint pairs and exchange the finished labels once.  This is the
 * collective half of the tdt_cluster_labels_sharded entry SURVEY.md section 8(b) proposes -- without a
 * communicator: every rank owns one buffer [nranks * slot_elems int32 | flags | control] allocated with
 * tdt_peer_alloc (cudaMalloc + cudaIpcGetMemHandle; the 64-byte handle travels to the other ranks by any host
 * channel) and mapped by the others with tdt_peer_open.  The clustering call writes its labels into slot `rank`
 * of the local buffer; tdt_peer_allgather enqueues ONE kernel that pushes that slot into every peer's buffer
 * with 16-byte stores, publishes an epoch flag (release, system scope) and waits for every peer's flag
 * (acquire).  No per-call arguments change, so the launch can be captured in a CUDA graph.  A peer that does
 * not arrive within 5 s sets *status_accum to 77 instead of hanging the device.  Between two exchanges every
 * rank must have consumed the previous result (any collective / barrier between steps).
 * bufs_h: HOST array of nranks device pointers (the local mapping of every rank's buffer).
 * ------------------------------------------------------------------------------------------ */
TDT_API size_t tdt_peer_buffer_bytes(int64_t slot_elems, int32_t nranks);
TDT_API int tdt_peer_alloc(size_t bytes, void **dev_ptr, unsigned char *handle64_h);
TDT_API int tdt_peer_open(const unsigned char *handle64_h, void **dev_ptr);
TDT_API int tdt_peer_close(void *dev_ptr);
TDT_API int tdt_peer_free(void *dev_ptr);
TDT_API int tdt_peer_allgather(void *const *bufs_h, int64_t slot_elems, int32_t rank, int32_t nranks,
                               int32_t *status_accum, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* TDT_B200_H */

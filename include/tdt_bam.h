/*
 * tdt_bam.h -- C ABI of libtdt_bam.so, the host-side BAM scanner that feeds the GPU coverage / signal path.
 *
 * The reference walks a BAM one pysam.AlignedSegment at a time (tiddit/__main__.py:229-242 for `--cov`,
 * tiddit/tiddit_signal.pyx:169-221 in the signal worker) and calls update_coverage once per read.  This
 * scanner replaces that iteration: BGZF blocks are inflated on `threads` host threads, the records of a
 * window are decoded into plain columns (what the two loops read: reference id, start, end, flag, mapq,
 * mate, template length, first / last CIGAR operation, "has an SA tag"), and the caller pushes the start /
 * end columns of a whole batch to tdt_coverage_accumulate_contigs (include/tdt_b200.h) in one call.  Only
 * the few reads that carry a signal are looked at record by record (tdt_bam_batch_data + rec_off).
 *
 * All pointers are HOST pointers.  Return value: >= 0 ok, < 0 error (text via tdt_bam_last_error(),
 * thread-local).  A reader is not thread-safe; distinct readers are independent.
 */
#ifndef TDT_BAM_H
#define TDT_BAM_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define TDT_BAM_API __attribute__((visibility("default")))
#else
#define TDT_BAM_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define TDT_BAM_OK 0
#define TDT_BAM_E_ARG (-1)    /* bad argument                                   */
#define TDT_BAM_E_IO (-2)     /* open / map failed                              */
#define TDT_BAM_E_FORMAT (-3) /* not BGZF / not BAM / truncated / CRC mismatch  */

typedef struct tdt_bam_reader tdt_bam_reader;

TDT_BAM_API const char *tdt_bam_last_error(void);

/* pysam.AlignmentFile(path, "r") -- tiddit/__main__.py:225, tiddit_signal.pyx:156,232.  threads <= 0: all cores. */
TDT_BAM_API int tdt_bam_open(const char *path, int threads, tdt_bam_reader **out);
TDT_BAM_API void tdt_bam_close(tdt_bam_reader *r);

/* samfile.header: the SAM text and the @SQ table (bam_header["SQ"][i]["SN"/"LN"], tiddit_coverage.pyx:13-16) */
TDT_BAM_API const char *tdt_bam_header_text(const tdt_bam_reader *r, int64_t *len);
TDT_BAM_API int32_t tdt_bam_n_ref(const tdt_bam_reader *r);
TDT_BAM_API const char *tdt_bam_ref_name(const tdt_bam_reader *r, int32_t i);
TDT_BAM_API int32_t tdt_bam_ref_len(const tdt_bam_reader *r, int32_t i);

/* One batch of `samfile.fetch(until_eof=True)`: up to max_reads records in file order, one entry per record in
 * every non-NULL column.  Returns the number of records (0 = end of file), < 0 on error.
 *   ref_id, pos       read.reference_id / read.reference_start (0-based)
 *   end               read.reference_end (pos + reference length of the CIGAR); -1 where pysam gives None
 *                     (unmapped flag or no CIGAR)
 *   mate_ref, mate_pos, tlen   read.next_reference_id / next_reference_start / template_length (= isize)
 *   flag, mapq        read.flag / read.mapq
 *   cig_first, cig_last        raw BAM CIGAR words (len << 4 | op) of the first / last operation, 0 if none
 *   has_sa            1 if the record has an "SA" aux tag (read.has_tag("SA"), tiddit_signal.pyx:199)
 *   rec_off           byte offset of the record (its block_size word) inside tdt_bam_batch_data()
 * The batch's raw records stay valid until the next call on this reader. */
TDT_BAM_API int64_t tdt_bam_read_columns(tdt_bam_reader *r, int64_t max_reads, int32_t *ref_id, int32_t *pos,
                                         int32_t *end, int32_t *mate_ref, int32_t *mate_pos, int32_t *tlen,
                                         uint16_t *flag, uint8_t *mapq, uint32_t *cig_first, uint32_t *cig_last,
                                         uint8_t *has_sa, int64_t *rec_off);
TDT_BAM_API const uint8_t *tdt_bam_batch_data(const tdt_bam_reader *r, int64_t *len);

/* The block decoder on its own (tests, diagnosis): in[0, in_len) = one complete raw DEFLATE stream (the payload of a BGZF
 * block, what pysam / htslib hand to zlib's inflate), out[0, out_len) = exactly its inflated bytes.  -> 1 decoded,
 * 0 refused (not a clean stream of exactly out_len bytes: the reader then lets zlib decide), < 0 bad argument.
 * The reader checks every block's CRC32 whichever decoder produced it. */
TDT_BAM_API int tdt_bam_inflate_raw(const uint8_t *in, int64_t in_len, uint8_t *out, int64_t out_len);
/* The CRC-32 the reader compares with a block's trailer (zlib's crc32(0, data, len); carry-less multiplication where the
 * host has it). */
TDT_BAM_API uint32_t tdt_bam_crc32(const uint8_t *data, int64_t len);

#ifdef __cplusplus
}
#endif
#endif

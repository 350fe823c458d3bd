/*
 * tdt_tab.h -- C ABI of libtdt_tab.so, the host-side reader of TIDDIT's signal tab files.
 *
 * The reference carries every discordant pair / split read / assembly contig from tiddit_signal to tiddit_cluster
 * through text: <prefix>_tiddit/{discordants,splits,contigs}_<sample>.tab (writer tiddit/tiddit_signal.pyx:298-326,
 * tiddit/tiddit_contig_analysis.pyx:78-91), and tiddit_cluster.main reads them back one `line.rstrip().split("\t")`
 * at a time (tiddit/tiddit_cluster.pyx:47-137) -- 20 M lines on a 30X genome, the largest part of the stage's time
 * once the clustering itself runs on the GPU.  This scanner maps a file, splits it at line ends over `threads` host
 * threads, parses the fields the reader uses into plain columns and interns the strings (read names, contig names,
 * orientation strings) in order of first appearance -- the packed layout tdt_cluster_labels / tdt_cluster_aggregate
 * take (include/tdt_b200.h).  The reference's per-record RULES (breakpoint choice by orientation, clamping, minimum
 * contig length; tiddit_cluster.pyx:52-72, 80-101) are applied to the columns by the caller (tiddit_b200/signals.py).
 *
 * Only a perfectly regular file is accepted: every used field non-empty and free of surrounding white space,
 * coordinates plain decimal integers, discordant lines of exactly 9 fields, split / contig lines of at least 11.
 * Anything else returns TDT_TAB_IRREGULAR and appends nothing; the caller then reads that file line by line, exactly
 * like the reference.
 *
 * All pointers are HOST pointers.  Return value: >= 0 ok, < 0 error (text via tdt_tab_last_error(), thread-local).
 */
#ifndef TDT_TAB_H
#define TDT_TAB_H

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define TDT_TAB_API __attribute__((visibility("default")))
#else
#define TDT_TAB_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define TDT_TAB_OK 0
#define TDT_TAB_E_ARG (-1)
#define TDT_TAB_E_IO (-2)
#define TDT_TAB_IRREGULAR (-4)

#define TDT_TAB_DISCORDANTS 0 /* name chrA chrB startA endA revA startB endB revB                      */
#define TDT_TAB_SPLITS 1      /* name chrA chrB posA revA posB revB startA endA startB endB [+ 8 per extra record] */
#define TDT_TAB_CONTIGS 2     /* same layout as splits                                                  */

typedef struct tdt_tab_set tdt_tab_set;

TDT_TAB_API const char *tdt_tab_last_error(void);

/* A record set: the records of every file parsed into it, with ONE interning context (a read name that occurs in the
 * discordant and in the split file gets one id, as in the reference's sets of names). */
TDT_TAB_API tdt_tab_set *tdt_tab_new(void);
TDT_TAB_API void tdt_tab_free(tdt_tab_set *set);

/* tiddit_cluster.pyx:47-49 / :75-77 / :106-108: one file.  -> records appended, or < 0.  threads <= 0: all cores. */
TDT_TAB_API int64_t tdt_tab_parse(tdt_tab_set *set, const char *path, int kind, int threads);

TDT_TAB_API int64_t tdt_tab_n(const tdt_tab_set *set);

/* int32 columns of all records so far: 0 name id, 1 chrA id, 2 chrB id, 3 orientation-A id, 4 orientation-B id */
TDT_TAB_API const int32_t *tdt_tab_col_i32(const tdt_tab_set *set, int which);
/* int64 columns: the numeric fields in file order -- discordants: 0 startA, 1 endA, 2 startB, 3 endB (4, 5 zero);
 * splits / contigs: 0 posA, 1 posB, 2 startA, 3 endA, 4 startB, 5 endB */
TDT_TAB_API const int64_t *tdt_tab_col_i64(const tdt_tab_set *set, int which);

/* string tables in order of first appearance: 0 read names, 1 contig names, 2 orientation strings.
 * -> number of strings; *blob = their bytes back to back, *offsets = size + 1 byte offsets into blob */
TDT_TAB_API int64_t tdt_tab_table(const tdt_tab_set *set, int table, const char **blob, const int64_t **offsets);

#ifdef __cplusplus
}
#endif
#endif

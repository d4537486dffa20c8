/* oracle/filtlong_oracle.h -- TEST INFRASTRUCTURE, not product code.
 *
 * Plain-C, single-threaded restatement of Filtlong's per-read scoring / trim-split /
 * sort-threshold path, written from the reference's algorithm (each function cites the
 * reference file:line it follows). Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this; the product path (filtlong_b200/) never does.
 *
 * Parity status: PINNED. tests/test_oracle_pinning.py checks this restatement against the real
 * reference objects (oracle/_ref/refdump, built from /root/reference/src by oracle/Makefile) on
 * the reference's own fixtures and on randomised inputs, bit-for-bit on every double.
 */
#ifndef FILTLONG_ORACLE_H
#define FILTLONG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- reference 16-mer set: kmers.cpp:28-239 + bloom_filter.h as configured ---- */
typedef struct orc_kmers orc_kmers;

orc_kmers *orc_kmers_new(void);
void orc_kmers_free(orc_kmers *k);
/* kmers.cpp:96-121 for one sequence; require_multiple_copies selects kmers.cpp:137-139 (0)
 * or kmers.cpp:142-166 (1). */
void orc_kmers_add_sequence(orc_kmers *k, const char *seq, size_t len, int require_multiple_copies);
int orc_kmers_contains(const orc_kmers *k, uint32_t kmer);      /* kmers.cpp:170-172 */
void orc_kmers_insert(orc_kmers *k, const uint32_t *kmers, size_t n);   /* test hook: m_kmers.insert of an explicit list */
uint64_t orc_kmers_size(const orc_kmers *k);                    /* m_kmers.size() */
size_t orc_kmers_dump(const orc_kmers *k, uint32_t *out, size_t cap); /* ascending order */
/* bloom_filter.h:569-583 specialised to a 4-byte key, salt j of 13; and the table size */
uint32_t orc_bloom_hash(uint32_t kmer, int j);
uint64_t orc_bloom_table_bits(void);
uint32_t orc_base_fwd(char c);                                  /* kmers.cpp:176-196 */
uint32_t orc_base_rev(char c);                                  /* kmers.cpp:199-219 */

/* ---- per-read scoring: read.cpp:25-144, 208-273 ---- */
typedef struct {
    int window_size;
    int trim, split_set, split;
    int min_length_set, min_length, max_length_set, max_length;
    int min_mean_q_set, min_window_q_set;
    double min_mean_q, min_window_q;
    double length_weight, mean_q_weight, window_q_weight;
    int target_bases_set, keep_percent_set;
    long long target_bases;
    double keep_percent;
} orc_params;

typedef struct {
    int parent;            /* index of the input read this row belongs to */
    int start, end;        /* range in the parent (whole read: 0, length) */
    int length;
    double mean_q, window_q, length_score;   /* RAW values (read.cpp:60-62) */
    int passed;            /* hard cut-offs only (read.cpp:65-73) */
    int first, last;       /* m_first/last_base_in_kmer */
    int n_bad, n_child;
    double norm_mean, norm_window, final_score;   /* after orc_finalize */
    int passed_final;
} orc_row;

double orc_qscore_to_quality(char c);                                   /* read.cpp:270-273 */
double orc_mean_quality(const double *q, size_t n);                     /* read.cpp:208-213 */
double orc_window_quality(const double *q, size_t n, size_t window);    /* read.cpp:216-236 */
double orc_length_score(int length);                                    /* read.cpp:241-244 */

/* Scores one read the way Read::Read does. kmers == NULL or empty -> Phred mode.
 * Writes the parent row, up to bad_cap (start,end) pairs and up to child_cap child rows
 * (children are scored by re-running the constructor logic on the substring, read.cpp:137).
 * Returns the number of children, or -1 if a capacity was exceeded. */
int orc_score_read(const orc_kmers *kmers, const char *seq, const char *qual, int length,
                   const orc_params *p, int read_index, orc_row *parent, int *bad_ranges, int bad_cap,
                   orc_row *children, int child_cap);

/* ---- global normalisation + selection: main.cpp:169-261 ---- */
typedef struct {
    double min_q, max_q, mean_q, stdev_q, min_z, max_z;
    int status;            /* 0 none, 1 not enough reads, 2 already below target, 3 thresholded */
    long long target, passed_bases, keeping;
} orc_summary;

/* rows = the "reads2" table in file order (children replace their parent, main.cpp:138-147).
 * Ties in the descending sort keep file order (the reference's std::sort is unstable, so only
 * the tie class is comparable). */
void orc_finalize(orc_row *rows, size_t n, long long total_bases, const orc_params *p, orc_summary *out);

#ifdef __cplusplus
}
#endif
#endif

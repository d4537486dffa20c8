/* oracle/filtlong_oracle.c -- TEST INFRASTRUCTURE, not product code. See filtlong_oracle.h.
 *
 * Plain-C restatement of the reference's hot path; every function names the reference
 * file:line whose operation order it follows. Compiled with -ffp-contract=off so that no
 * multiply-add is fused (the stock reference build is SSE2 without FMA, Makefile:11-20).
 */
#include "filtlong_oracle.h"

#include <limits.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Base encoders, kmers.cpp:176-219. Forward: A0 C1 G2 T3 (case-insensitive), anything else 0.
 * Reverse: the complement code in bits 31:30 (T0 G1 C2 A3), anything else 0 -- so a non-ACGT
 * base is 'A' on the forward strand but behaves like 'T' on the reverse strand.
 * ---------------------------------------------------------------------------------------- */
uint32_t orc_base_fwd(char c) {
    switch (c) {
        case 'C': case 'c': return 1u;
        case 'G': case 'g': return 2u;
        case 'T': case 't': return 3u;
        default: return 0u;
    }
}

uint32_t orc_base_rev(char c) {
    switch (c) {
        case 'G': case 'g': return 1u << 30;
        case 'C': case 'c': return 2u << 30;
        case 'A': case 'a': return 3u << 30;
        default: return 0u;
    }
}

/* ------------------------------------------------------------------------------------------
 * Bloom filter as configured by kmers.cpp:29-39: 1e8 projected elements, FPP 1e-4, seed
 * 0xA5A5A5A5. bloom_filter.h:108-160 then yields 13 hashes over 1,917,295,480 bits, and
 * bloom_filter.h:183-195,519-528 yields the 13 salts below (derived values; pinned against the
 * reference's own object by tests via `refdump` bloom mode). For a 4-byte key hash_ap runs only
 * its "remaining_length >= 4, loop == 0" branch (bloom_filter.h:569-583):
 *     h = s ^ ~((s << 11) + (key ^ (s >> 5)))          bit = h % table_bits (bloom_filter.h:463)
 * ---------------------------------------------------------------------------------------- */
#define ORC_BLOOM_BITS 1917295480ull
#define ORC_BLOOM_K 13
static const uint32_t orc_salts[ORC_BLOOM_K] = {
    0x1B5793D2u, 0x81BDFA38u, 0xEB8E30D5u, 0x45B52496u, 0x85C1FE3Cu, 0x3DACB627u, 0x78776869u,
    0x94A40D1Eu, 0x5F9BB638u, 0x40FB59D5u, 0x8174BDB2u, 0x0B466EAAu, 0x209D29A7u};

uint32_t orc_bloom_hash(uint32_t kmer, int j) {
    uint32_t s = orc_salts[j];
    return s ^ ~((s << 11) + (kmer ^ (s >> 5)));
}

uint64_t orc_bloom_table_bits(void) { return ORC_BLOOM_BITS; }

/* ------------------------------------------------------------------------------------------
 * Kmers: one open-addressing table stands in for both m_kmers (state 0xFF) and m_kmer_counts
 * (state 2 or 3). Entry = occupied(63) | state(39:32) | key(31:0).
 * ---------------------------------------------------------------------------------------- */
struct orc_kmers {
    uint64_t *tab;
    size_t cap, used, in_set;
    uint8_t *bloom;        /* lazily allocated, ORC_BLOOM_BITS/8 bytes */
};

#define ENT_OCC (1ull << 63)
#define ENT_KEY(e) ((uint32_t)(e))
#define ENT_STATE(e) ((unsigned)(((e) >> 32) & 0xFF))
#define STATE_SET 0xFFu

static size_t tab_slot(const orc_kmers *k, uint32_t key) {
    uint64_t h = (uint64_t)key * 0x9E3779B97F4A7C15ull;
    size_t i = (size_t)(h >> 20) & (k->cap - 1);
    while ((k->tab[i] & ENT_OCC) && ENT_KEY(k->tab[i]) != key) i = (i + 1) & (k->cap - 1);
    return i;
}

static void tab_grow(orc_kmers *k) {
    uint64_t *old = k->tab;
    size_t oldcap = k->cap;
    k->cap = oldcap ? oldcap * 2 : (1u << 16);
    k->tab = (uint64_t *)calloc(k->cap, sizeof(uint64_t));
    for (size_t i = 0; i < oldcap; ++i)
        if (old[i] & ENT_OCC) k->tab[tab_slot(k, ENT_KEY(old[i]))] = old[i];
    free(old);
}

orc_kmers *orc_kmers_new(void) {
    orc_kmers *k = (orc_kmers *)calloc(1, sizeof(orc_kmers));
    tab_grow(k);
    return k;
}

void orc_kmers_free(orc_kmers *k) {
    if (!k) return;
    free(k->tab);
    free(k->bloom);
    free(k);
}

static void tab_put(orc_kmers *k, uint32_t key, unsigned state) {
    if ((k->used + 1) * 10 > k->cap * 6) tab_grow(k);
    size_t i = tab_slot(k, key);
    if (!(k->tab[i] & ENT_OCC)) k->used++;
    else if (ENT_STATE(k->tab[i]) == STATE_SET) k->in_set--;
    if (state == STATE_SET) k->in_set++;
    k->tab[i] = ENT_OCC | ((uint64_t)state << 32) | key;
}

static unsigned tab_get(const orc_kmers *k, uint32_t key) { /* 0 = absent */
    size_t i = tab_slot(k, key);
    return (k->tab[i] & ENT_OCC) ? ENT_STATE(k->tab[i]) : 0u;
}

static int bloom_contains(const orc_kmers *k, uint32_t key) { /* bloom_filter.h:303-319 */
    for (int j = 0; j < ORC_BLOOM_K; ++j) {
        uint64_t b = orc_bloom_hash(key, j) % ORC_BLOOM_BITS;
        if (!(k->bloom[b >> 3] & (1u << (b & 7)))) return 0;
    }
    return 1;
}

static void bloom_insert(orc_kmers *k, uint32_t key) { /* bloom_filter.h:260-273 */
    for (int j = 0; j < ORC_BLOOM_K; ++j) {
        uint64_t b = orc_bloom_hash(key, j) % ORC_BLOOM_BITS;
        k->bloom[b >> 3] |= (uint8_t)(1u << (b & 7));
    }
}

/* kmers.cpp:137-139 */
static void add_one_copy(orc_kmers *k, uint32_t key) {
    if (tab_get(k, key) != STATE_SET) tab_put(k, key, STATE_SET);
}

/* kmers.cpp:142-166: in set -> skip; Bloom miss -> Bloom insert; Bloom hit and not counted ->
 * count 2; else ++count and promote at 4 (required_kmer_copies, kmers.cpp:41). */
static void add_multiple_copies(orc_kmers *k, uint32_t key) {
    unsigned st = tab_get(k, key);
    if (st == STATE_SET) return;
    if (!k->bloom) k->bloom = (uint8_t *)calloc(ORC_BLOOM_BITS / 8, 1);
    if (!bloom_contains(k, key)) bloom_insert(k, key);
    else if (st == 0) tab_put(k, key, 2);
    else {
        unsigned seen = st + 1;
        if (seen >= 4) tab_put(k, key, STATE_SET);
        else tab_put(k, key, seen);
    }
}

/* kmers.cpp:96-121: sequences shorter than 16 give nothing; forward k-mer has its first base in
 * bits 31:30; the reverse k-mer shifts right and takes the new complement on top; forward is
 * added before reverse at every position. */
void orc_kmers_add_sequence(orc_kmers *k, const char *seq, size_t len, int multi) {
    if (len < 16) return;
    uint32_t fwd = 0, rev = 0;
    for (int i = 0; i < 16; ++i) {
        fwd = (fwd << 2) | orc_base_fwd(seq[i]);
        rev = (rev >> 2) | orc_base_rev(seq[i]);
    }
    void (*add)(orc_kmers *, uint32_t) = multi ? add_multiple_copies : add_one_copy;
    add(k, fwd);
    add(k, rev);
    for (size_t i = 16; i < len; ++i) {
        fwd = (fwd << 2) | orc_base_fwd(seq[i]);
        rev = (rev >> 2) | orc_base_rev(seq[i]);
        add(k, fwd);
        add(k, rev);
    }
}

int orc_kmers_contains(const orc_kmers *k, uint32_t kmer) { return tab_get(k, kmer) == STATE_SET; }

/* Test hook: load an explicit set (e.g. one exported by the CUDA build at a size the CPU hashing of the
 * short reads cannot follow), so that SCORING parity can still be checked against this restatement. */
void orc_kmers_insert(orc_kmers *k, const uint32_t *kmers, size_t n) {
    for (size_t i = 0; i < n; ++i) add_one_copy(k, kmers[i]);
}

uint64_t orc_kmers_size(const orc_kmers *k) { return k ? k->in_set : 0; }

static int cmp_u32(const void *a, const void *b) {
    uint32_t x = *(const uint32_t *)a, y = *(const uint32_t *)b;
    return x < y ? -1 : (x > y);
}

size_t orc_kmers_dump(const orc_kmers *k, uint32_t *out, size_t cap) {
    size_t n = 0;
    for (size_t i = 0; i < k->cap; ++i)
        if ((k->tab[i] & ENT_OCC) && ENT_STATE(k->tab[i]) == STATE_SET) {
            if (n < cap) out[n] = ENT_KEY(k->tab[i]);
            n++;
        }
    if (n <= cap) qsort(out, n, sizeof(uint32_t), cmp_u32);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * Per-read scoring
 * ---------------------------------------------------------------------------------------- */
double orc_qscore_to_quality(char c) { /* read.cpp:270-273; char is signed on the reference's targets */
    int q = (signed char)c - 33;
    return 1.0 - pow(10.0, -q / 10.0);
}

double orc_mean_quality(const double *q, size_t n) { /* read.cpp:208-213 */
    double sum = 0.0;
    for (size_t i = 0; i < n; ++i) sum += q[i];
    return 100.0 * sum / n;
}

double orc_window_quality(const double *q, size_t n, size_t window) { /* read.cpp:216-236 */
    if (n <= window) return orc_mean_quality(q, n);
    double sum = 0.0;
    for (size_t i = 0; i < window; ++i) sum += q[i];
    double w = sum / window;
    double best = w;
    for (size_t j = window; j < n; ++j) {
        size_t i = j - window;
        w -= q[i] / window;
        w += q[j] / window;
        if (w < best) best = w;
    }
    if (best < 0.5 / window) best = 0.0;
    return 100.0 * best;
}

double orc_length_score(int length) { /* read.cpp:241-244 */
    double half = 5000.0;
    return 100.0 * (1.0 + (-half / (length + half)));
}

static int hard_cutoffs(const orc_params *p, int length, double mean_q, double window_q) { /* read.cpp:65-73 */
    if (p->min_length_set && length < p->min_length) return 0;
    else if (p->max_length_set && length > p->max_length) return 0;
    else if (p->min_mean_q_set && mean_q < p->min_mean_q) return 0;
    else if (p->min_window_q_set && window_q < p->min_window_q) return 0;
    return 1;
}

int orc_score_read(const orc_kmers *kmers, const char *seq, const char *qual, int length,
                   const orc_params *p, int read_index, orc_row *row, int *bad, int bad_cap,
                   orc_row *children, int child_cap) {
    int kmer_mode = kmers && orc_kmers_size(kmers) > 0;          /* read.cpp:35: kmers->empty() */
    double *q = (double *)calloc(length > 0 ? (size_t)length : 1, sizeof(double));
    if (!kmer_mode) {
        for (int i = 0; i < length; ++i) q[i] = orc_qscore_to_quality(qual[i]);   /* read.cpp:36-38 */
    } else if (length >= 16) {                                                    /* read.cpp:43-58 */
        uint32_t kmer = 0;
        for (int i = 0; i < 16; ++i) kmer = (kmer << 2) | orc_base_fwd(seq[i]);
        for (int i = 15; i < length; ++i) {
            if (i > 15) kmer = (kmer << 2) | orc_base_fwd(seq[i]);
            if (orc_kmers_contains(kmers, kmer))
                for (int j = i - 15; j <= i; ++j) q[j] = 1.0;
        }
    }
    memset(row, 0, sizeof(*row));
    row->parent = read_index;
    row->start = 0;
    row->end = length;
    row->length = length;
    row->mean_q = orc_mean_quality(q, (size_t)length);
    row->window_q = orc_window_quality(q, (size_t)length, (size_t)p->window_size);
    row->length_score = orc_length_score(length);
    row->passed = hard_cutoffs(p, length, row->mean_q, row->window_q);
    row->first = row->last = -1;

    int n_bad = 0, n_child = 0, overflow = 0;
    if (kmer_mode) {
        for (int i = 0; i < length; ++i)                                          /* read.cpp:75-84 */
            if (q[i] != 0) {
                if (row->first == -1) row->first = i;
                row->last = i + 1;
            }
        if (p->trim || p->split_set) {
            if (p->split_set) {                                                   /* read.cpp:89-103 */
                int i = 0;
                while (i < length) {
                    if (q[i] == 0.0) {
                        int s = i;
                        while (i < length && q[i] == 0.0) ++i;
                        if (i - s >= p->split) {
                            if (n_bad < bad_cap) { bad[2 * n_bad] = s; bad[2 * n_bad + 1] = i; }
                            else overflow = 1;
                            n_bad++;
                        }
                    } else ++i;
                }
            }
            if (overflow) { free(q); return -1; }
            if (p->trim) {                                                        /* read.cpp:106-117 */
                if (row->first > 0) {
                    if (n_bad == 0 || !(bad[0] == 0 && bad[1] == row->first)) {
                        if (n_bad >= bad_cap) { free(q); return -1; }
                        memmove(bad + 2, bad, sizeof(int) * 2 * (size_t)n_bad);
                        bad[0] = 0;
                        bad[1] = row->first;
                        n_bad++;
                    }
                }
                if (row->last != -1 && row->last < length) {
                    if (n_bad == 0 || !(bad[2 * (n_bad - 1)] == row->last && bad[2 * (n_bad - 1) + 1] == length)) {
                        if (n_bad >= bad_cap) { free(q); return -1; }
                        bad[2 * n_bad] = row->last;
                        bad[2 * n_bad + 1] = length;
                        n_bad++;
                    }
                }
            }
            if (n_bad > 0) {                                                      /* read.cpp:119-141 */
                int rs = 0;
                for (int b = 0; b <= n_bad; ++b) {
                    int re = (b < n_bad) ? bad[2 * b] : length;
                    if (re - rs > 0) {
                        if (n_child >= child_cap) { free(q); return -1; }
                        orc_row *c = &children[n_child];
                        int gb[2];
                        /* the child is a full Read on the substring (read.cpp:137); a child never
                         * yields bad ranges of its own, asserted by the pinning test */
                        int sub = orc_score_read(kmers, seq + rs, qual ? qual + rs : NULL, re - rs, p,
                                                 read_index, c, gb, 1, NULL, 0);
                        if (sub != 0) { free(q); return -1; }
                        c->start = rs;
                        c->end = re;
                        n_child++;
                    }
                    if (b < n_bad) rs = bad[2 * b + 1];
                }
            }
        }
    }
    row->n_bad = n_bad;
    row->n_child = n_child;
    free(q);
    return n_child;
}

/* ------------------------------------------------------------------------------------------
 * Global normalisation + selection
 * ---------------------------------------------------------------------------------------- */
static double final_score(double ls, double mq, double wq, double lw, double mw, double ww) { /* read.cpp:249-267 */
    double product = pow(ls, lw) * pow(mq, mw);
    double total = lw + mw;
    double fs = pow(product, 1.0 / total);
    double sf;
    if (mq > 0.0) {
        double r = wq / mq;
        sf = (1.0 < r) ? 1.0 : r;          /* std::min(r, 1.0): a NaN r stays NaN */
    } else sf = 1.0;
    total = lw + mw + ww;
    double wf = ww / total;
    double nwf = 1.0 - wf;
    sf = nwf + (sf * wf);
    return fs * sf;
}

static void merge_sort_desc(size_t *idx, size_t *tmp, size_t n, const orc_row *rows) {
    if (n < 2) return;
    size_t h = n / 2;
    merge_sort_desc(idx, tmp, h, rows);
    merge_sort_desc(idx + h, tmp, n - h, rows);
    size_t a = 0, b = h, o = 0;
    while (a < h && b < n) {
        /* take from the right run only if it is strictly better: stable, ties keep file order */
        if (rows[idx[b]].final_score > rows[idx[a]].final_score) tmp[o++] = idx[b++];
        else tmp[o++] = idx[a++];
    }
    while (a < h) tmp[o++] = idx[a++];
    while (b < n) tmp[o++] = idx[b++];
    memcpy(idx, tmp, n * sizeof(size_t));
}

void orc_finalize(orc_row *rows, size_t n, long long total_bases, const orc_params *p, orc_summary *out) {
    memset(out, 0, sizeof(*out));
    double min_q = 100.0, max_q = 0.0, sum = 0.0;                                 /* main.cpp:170-178 */
    for (size_t i = 0; i < n; ++i) {
        sum += rows[i].mean_q;
        if (rows[i].mean_q > max_q) max_q = rows[i].mean_q;
        if (rows[i].mean_q < min_q) min_q = rows[i].mean_q;
    }
    double mean = sum / n;
    double sd_sum = 0.0;                                                          /* main.cpp:180-186 */
    for (size_t i = 0; i < n; ++i) {
        double d = rows[i].mean_q - mean;
        sd_sum += d * d;
    }
    double sd = sqrt(sd_sum / n);
    double min_z, max_z;
    if (sd > 0.0) { min_z = (min_q - mean) / sd; max_z = (max_q - mean) / sd; }
    else { min_z = 1.0; max_z = 1.0; }
    double span = max_z - min_z;
    for (size_t i = 0; i < n; ++i) {                                              /* main.cpp:202-212 */
        double ratio = rows[i].window_q / rows[i].mean_q;
        if (ratio > 1.0) ratio = 1.0;
        double z = (rows[i].mean_q - mean) / sd;
        rows[i].norm_mean = 100.0 * (z - min_z) / span;
        rows[i].norm_window = rows[i].norm_mean * ratio;
        rows[i].final_score = final_score(rows[i].length_score, rows[i].norm_mean, rows[i].norm_window,
                                          p->length_weight, p->mean_q_weight, p->window_q_weight);
        rows[i].passed_final = rows[i].passed;
    }
    out->min_q = min_q; out->max_q = max_q; out->mean_q = mean; out->stdev_q = sd;
    out->min_z = min_z; out->max_z = max_z;
    if (!(p->target_bases_set || p->keep_percent_set)) return;
    long long passed_bases = 0;                                                   /* main.cpp:221-226 */
    for (size_t i = 0; i < n; ++i) if (rows[i].passed) passed_bases += rows[i].length;
    long long target = p->target_bases_set ? p->target_bases : LLONG_MAX;         /* main.cpp:229-237 */
    if (p->keep_percent_set) {
        long long keep_target = (long long)((p->keep_percent / 100.0) * total_bases);
        if (keep_target < target) target = keep_target;
    }
    out->target = target;
    out->passed_bases = passed_bases;
    if (target >= total_bases) { out->status = 1; return; }                       /* main.cpp:239-244 */
    if (target >= passed_bases) { out->status = 2; return; }
    out->status = 3;
    size_t *idx = (size_t *)malloc(n * sizeof(size_t)), *tmp = (size_t *)malloc(n * sizeof(size_t));
    for (size_t i = 0; i < n; ++i) idx[i] = i;
    merge_sort_desc(idx, tmp, n, rows);                                           /* main.cpp:247-248 */
    long long so_far = 0;                                                         /* main.cpp:251-257 */
    for (size_t r = 0; r < n; ++r) {
        orc_row *row = &rows[idx[r]];
        if (row->passed_final && so_far < target) so_far += row->length;
        else row->passed_final = 0;
    }
    out->keeping = so_far;
    free(idx);
    free(tmp);
}

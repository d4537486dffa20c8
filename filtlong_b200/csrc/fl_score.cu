// filtlong_b200/csrc/fl_score.cu -- per-read scoring, replaces Read::Read (reference src/read.cpp:25-144).
//
// Two modes, chosen by the state of the k-mer set exactly like read.cpp:35 (`kmers->empty()`):
//
//  Phred mode (read.cpp:35-39) lives in fl_phred.cu: one THREAD per work item walks the quality string in the reference's
//    own operation order -- sum += q[c] for the mean (read.cpp:208-213), and the incremental
//    window recurrence w -= a[c_out]; w += a[c_in] (read.cpp:216-236) -- so mean and window
//    quality come out bit-identical to the reference (their rounding depends on the order of
//    the additions). q[] and a[] = q[]/window_size are 256-entry tables evaluated with the host
//    libm (read.cpp:270-273) and replicated in shared memory so that every lane owns its banks.
//    Rows are issued longest-first (length buckets) so lanes of a warp carry similar work.
//
//  k-mer mode (read.cpp:43-58): kernel A is the HBM-bound hot loop -- a warp streams 1024 bases
//    per step (coalesced 8-byte loads of 2-bit codes), forms the 32 forward 16-mers of each lane
//    with funnel shifts, probes the 512 MiB direct-address bitmap (one 32-byte sector per base
//    when the set exceeds L2), and paints 16-base hits into a 1-bit-per-base mask with shuffles.
//    Kernels B then work on that mask only (L/8 bytes per read): popcount -> mean (exact: the
//    reference sums 1.0s), first/last base in a k-mer (read.cpp:75-84), bad ranges and child
//    ranges (read.cpp:89-130), and the serial window recurrence on {0, 1/ws} per row. Children
//    are NOT re-probed: a matching 16-mer never overlaps a bad range, so a child's mask is the
//    parent's mask restricted to the child range; the reference re-runs the whole
//    constructor instead (read.cpp:137).
#include "fl_device.cuh"

namespace {

// ---------------------------------------------------------------------------------------------
// shared scalar helpers (double arithmetic in exactly the reference's order; file is compiled
// with --fmad=false so nothing is contracted)
// ---------------------------------------------------------------------------------------------


// ---------------------------------------------------------------------------------------------
// k-mer mode, kernel A: probe + paint
// ---------------------------------------------------------------------------------------------
struct ProbeArgs {
    const uint32_t *seq2b;
    const uint64_t *off;
    const int32_t *len;
    const unsigned long long *tile_start;   // [n+1]
    uint32_t n;
    unsigned long long n_tiles;
    const uint32_t *bitmap;
    const uint32_t *anchor;                  // position-anchored table (fl_anchor_slot), used when ANCH
    const unsigned long long *filter;        // L2-resident pre-filter (fl_kmers.cu), used when FILT
    unsigned filter_log2_words;
    int filter_kind;
    uint32_t *mask;                          // 1 bit per padded base, same coordinates as the arena
};

// One 4-byte load from the 512 MiB membership bitmap per k-mer. The flavour of the load decides how
// much HBM traffic a random probe costs (measured with ncu, profiles/): MODE 0 = ld.global.nc (L1
// allocates and pulls whole 128-byte lines), 1 = ld.global.cg (L2 only, sector granular),
// 2 = ld.global.nc.L1::no_allocate.
// MODE 3 adds L2 eviction policies: the pre-filter is loaded evict_last, the bitmap probes and the
// mask stores evict_first, so that streaming traffic does not push the filter out of L2.
__device__ __forceinline__ unsigned long long l2_policy_evict_last() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ unsigned long long l2_policy_evict_first() {
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

// word holding the membership bit of `kmer`, and the bit's index, for the k-mer that starts at a read
// position whose low two bits are pos_lo2
template <bool ANCH>
__device__ __forceinline__ void probe_slot(uint32_t kmer, unsigned pos_lo2, uint32_t &word, uint32_t &bit) {
    if (ANCH) fl_anchor_slot(kmer, 3u - pos_lo2, word, bit);
    else { word = kmer >> 5; bit = kmer & 31u; }
}

template <int MODE>
__device__ __forceinline__ uint32_t probe(const uint32_t *__restrict__ bitmap, uint32_t word_index, unsigned long long pol_first) {
    const uint32_t *p = bitmap + word_index;
    if (MODE == 3) {
        uint32_t v;
        asm volatile("ld.global.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol_first));
        return v;
    }
    if (MODE == 1) return __ldcg(p);
    if (MODE == 2) {
        uint32_t v;
        asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
        return v;
    }
    return __ldg(p);
}

template <int MODE, bool FILT, bool ANCH>
__global__ void __launch_bounds__(256, 4) k_probe_paint(ProbeArgs a) {
    const unsigned lane = threadIdx.x & 31;
    const uint32_t *__restrict__ table = ANCH ? a.anchor : a.bitmap;
    const unsigned long long pol_first = MODE == 3 ? l2_policy_evict_first() : 0ull;
    const unsigned long long pol_last = MODE == 3 ? l2_policy_evict_last() : 0ull;
    (void)pol_last;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long n_warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long tile = warp; tile < a.n_tiles; tile += n_warps) {
        const uint32_t s = fl_find_seq(a.tile_start, a.n, tile);
        const int L = a.len[s];
        const unsigned long long off = a.off[s];
        const uint32_t *seqw = a.seq2b + (off >> 4);
        uint32_t *maskw = a.mask + (off >> 5);
        const unsigned long long padded = ((unsigned long long)L + FL_ALIGN_BASES - 1) & ~(unsigned long long)(FL_ALIGN_BASES - 1);
        const unsigned long long tile_base = (tile - a.tile_start[s]) * FL_TILE_BASES;

        // hits of the 16 k-mer starts just before the tile (they paint into the tile's first bases)
        uint32_t carry = 0;   // bit k = hit of the k-mer starting at base (run_start - 32 + k)
        if (tile_base > 0) {
            uint32_t wa = __ldg(seqw + (tile_base >> 4) - 1), wb = __ldg(seqw + (tile_base >> 4));
            uint32_t hit = 0;
            if (lane < 16) {
                unsigned long long b = tile_base - 16 + lane;
                if (b + (FL_K - 1) < (unsigned long long)L) {
                    uint32_t k = __funnelshift_l(wb, wa, 2 * lane);
                    uint32_t word, bit;
                    probe_slot<ANCH>(k, lane & 3u, word, bit);         // tile_base is a multiple of 4
                    hit = (probe<MODE>(table, word, pol_first) >> bit) & 1u;
                }
            }
            carry = __ballot_sync(0xffffffffu, hit) << 16;
        }
        for (int step = 0; step < FL_TILE_STEPS; ++step) {
            const unsigned long long sb = tile_base + (unsigned long long)step * FL_STEP_BASES;
            if (sb >= padded) break;
            const LaneWords w = fl_load_lane_words(seqw, sb, padded, lane);
            const unsigned long long lb = sb + 32ull * lane;
            // number of valid k-mer starts in this lane's run: starts b with b + 15 < L
            long long nv = (long long)L - (FL_K - 1) - (long long)lb;
            const int nvalid = nv <= 0 ? 0 : (nv >= 32 ? 32 : (int)nv);
            uint32_t h = 0;
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                uint32_t words[16];
                uint32_t go = 0xFFFFu;                // which of the 16 k-mers still need the exact bitmap
                if (FILT) {
                    // 16 independent loads from the 64 MiB pre-filter (kept in L2): most k-mers of a
                    // noisy read are absent and stop here, without touching HBM
                    unsigned long long f[16];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t word;
                        unsigned long long fb;
                        fl_filter_slot(fl_kmer_at(w, half * 16 + i), a.filter_log2_words, a.filter_kind, word, fb);
                        if (half * 16 + i < nvalid) {
                            if (MODE == 3) asm volatile("ld.global.L2::cache_hint.u64 %0, [%1], %2;" : "=l"(f[i]) : "l"(a.filter + word), "l"(pol_last));
                            else if (a.filter_kind & 2) f[i] = __ldcg(a.filter + word);   // L2 only (no L1 line fill)
                            else f[i] = __ldg(a.filter + word);
                        } else f[i] = 0ull;
                    }
                    go = 0;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        uint32_t word;
                        unsigned long long fb;
                        fl_filter_slot(fl_kmer_at(w, half * 16 + i), a.filter_log2_words, a.filter_kind, word, fb);
                        go |= ((f[i] & fb) == fb ? 1u : 0u) << i;
                    }
                }
                if (ANCH) {
                    // the four 16-mers starting at 4g .. 4g+3 of the lane's run (a multiple of 32, so of 4)
                    // share one 32-byte sector of the anchored table: ONE 256-bit load per group
                    uint32_t sec[4][8];
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int p0 = half * 16 + 4 * g;
                        const uint32_t key = (fl_kmer_at(w, p0 + 3) >> 6) & 0x3FFFFFFu;   // bases p0+3 .. p0+15
                        const bool need = p0 < nvalid && ((go >> (4 * g)) & 0xFu);
                        if (need) {
                            asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                         : "=r"(sec[g][0]), "=r"(sec[g][1]), "=r"(sec[g][2]), "=r"(sec[g][3]), "=r"(sec[g][4]),
                                           "=r"(sec[g][5]), "=r"(sec[g][6]), "=r"(sec[g][7])
                                         : "l"(table + (size_t)key * 8u));
                        } else {
#pragma unroll
                            for (int q = 0; q < 8; ++q) sec[g][q] = 0u;
                        }
                    }
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int p = half * 16 + 4 * g + j;
                            uint32_t word, bit;
                            fl_anchor_slot(fl_kmer_at(w, p), 3u - (unsigned)j, word, bit);
                            const uint32_t v = (word & 1u) ? sec[g][2 * (3 - j) + 1] : sec[g][2 * (3 - j)];
                            const uint32_t ok = (p < nvalid && ((go >> (4 * g + j)) & 1u)) ? 1u : 0u;
                            h |= (((v >> bit) & 1u) & ok) << p;
                        }
                    }
                } else {
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int p = half * 16 + i;
                        const uint32_t k = fl_kmer_at(w, p);
                        words[i] = (p < nvalid && ((go >> i) & 1u)) ? probe<MODE>(table, k >> 5, pol_first) : 0u;     // read.cpp:52
                    }
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int p = half * 16 + i;
                        const uint32_t k = fl_kmer_at(w, p);
                        h |= ((words[i] >> (k & 31)) & 1u) << p;
                    }
                }
            }
            // paint: base covered if any of the 16 k-mers ending at or after it hit (read.cpp:53-54)
            uint32_t prev = __shfl_up_sync(0xffffffffu, h, 1);
            if (lane == 0) prev = carry;
            unsigned long long y = ((unsigned long long)h << 32) | prev;
            y |= y << 1;
            y |= y << 2;
            y |= y << 4;
            y |= y << 8;
            if (lb < padded) {
                if (MODE == 3) asm volatile("st.global.L2::cache_hint.u32 [%0], %1, %2;" :: "l"(maskw + (lb >> 5)), "r"((uint32_t)(y >> 32)), "l"(pol_first));
                else maskw[lb >> 5] = (uint32_t)(y >> 32);
            }
            carry = __shfl_sync(0xffffffffu, h, 31);
        }
    }
}

__global__ void k_tiles_of(const int32_t *__restrict__ len, uint32_t n, unsigned long long *__restrict__ tiles) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) tiles[i] = fl_tiles_of(len[i] > 0 ? (int)(((unsigned)len[i] + FL_ALIGN_BASES - 1) & ~(FL_ALIGN_BASES - 1)) : 0);
}

// ---------------------------------------------------------------------------------------------
// k-mer mode, kernels B: everything else from the 1-bit mask
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int find_next(const uint32_t *__restrict__ m, int p, int end, bool want_one) {
    while (p < end) {
        uint32_t w = m[p >> 5];
        if (!want_one) w = ~w;
        w &= 0xFFFFFFFFu << (p & 31);
        if (w) {
            int q = (p & ~31) + __ffs(w) - 1;
            return q < end ? q : end;
        }
        p = (p & ~31) + 32;
    }
    return end;
}

// Walks the bad ranges of read.cpp:89-117 in order and reports child ranges (read.cpp:119-130).
// F(start, end, child_index). Returns n_child; *n_bad_out = m_bad_ranges.size().
template <typename F>
__device__ __forceinline__ int enumerate_children(const uint32_t *__restrict__ m, int L, int first, int last,
                                                  const fl_params &p, int *n_bad_out, F emit) {
    int n_bad = 0, n_child = 0, rs = 0;
    auto bad = [&](int s, int e) {
        ++n_bad;
        if (s - rs > 0) emit(rs, s, n_child++);
        rs = e;
    };
    if (first < 0) {
        // no base is in a k-mer: one zero run [0, L); trimming adds nothing (read.cpp:107,112)
        if (p.split_set && L > 0 && L >= p.split) bad(0, L);
    } else {
        const bool lead_is_split = p.split_set && first >= p.split && first > 0;
        const bool tail_is_split = p.split_set && (L - last) >= p.split && last < L;
        if (lead_is_split) bad(0, first);
        else if (p.trim && first > 0) bad(0, first);                     // read.cpp:107-111
        if (p.split_set) {                                               // read.cpp:89-103, runs inside [first, last)
            int i = first;
            while (i < last) {
                int z = find_next(m, i, last, false);
                if (z >= last) break;
                int o = find_next(m, z, last, true);
                if (o - z >= p.split) bad(z, o);
                i = o;
            }
        }
        if (tail_is_split) bad(last, L);
        else if (p.trim && last < L) bad(last, L);                       // read.cpp:112-116
    }
    if (n_bad > 0 && L - rs > 0) emit(rs, L, n_child++);                 // read.cpp:127-129
    *n_bad_out = n_bad;
    return n_child;
}

struct KmerArgs {
    const uint32_t *mask;
    const uint64_t *off;
    const int32_t *len;
    uint32_t n;
    fl_params p;
    int32_t *r_len, *r_first, *r_last, *r_nbad, *r_nchild;
    double *r_mean, *r_window;
    uint8_t *r_passed;
    unsigned long long *r_rowstart;
    unsigned long long *rows_per_read;   // scratch [n]
    unsigned long long read_base, row_base;
};

// B1: per read -- first/last base in a k-mer, number of bad ranges and children
__global__ void __launch_bounds__(256) k_kmer_ranges(KmerArgs a) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const int L = a.len[r];
    const uint32_t *m = a.mask + (a.off[r] >> 5);
    int first = -1, last = -1;
    const int nw = (L + 31) >> 5;
    for (int w = 0; w < nw; ++w) {                       // read.cpp:77-84
        uint32_t x = m[w];
        if (x) {
            if (first < 0) first = (w << 5) + __ffs(x) - 1;
            last = (w << 5) + 32 - __clz(x);
        }
    }
    int n_bad = 0, n_child = 0;
    if (a.p.trim || a.p.split_set)
        n_child = enumerate_children(m, L, first, last, a.p, &n_bad, [](int, int, int) {});
    a.r_len[r] = L;
    a.r_first[r] = first;
    a.r_last[r] = last;
    a.r_nbad[r] = n_bad;
    a.r_nchild[r] = n_child;
    a.rows_per_read[r] = n_child > 0 ? (unsigned long long)n_child : 1ull;
    a.r_rowstart[r] = a.row_base + r;   // final when nothing can have children; else replaced by the scan
}

// B2: per read -- row descriptors (children, or the read itself)
struct RowArgs {
    const uint32_t *mask;
    const uint64_t *off;
    const int32_t *len;
    uint32_t n;
    fl_params p;
    const int32_t *r_first, *r_last, *r_nchild;
    unsigned long long *r_rowstart;             // in: exclusive scan of rows per read (batch-local); out: + row_base
    uint32_t *w_parent;                         // offset to this batch's first row
    int32_t *w_start, *w_end, *w_len;
    unsigned long long read_base, row_base;
};

__global__ void __launch_bounds__(256) k_kmer_rows(RowArgs a) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    const int L = a.len[r];
    const unsigned long long rs = a.r_rowstart[r];
    a.r_rowstart[r] = rs + a.row_base;
    if (a.r_nchild[r] == 0) {
        a.w_parent[rs] = (uint32_t)(a.read_base + r);
        a.w_start[rs] = 0;
        a.w_end[rs] = L;
        a.w_len[rs] = L;
        return;
    }
    const uint32_t *m = a.mask + (a.off[r] >> 5);
    int nb;
    enumerate_children(m, L, a.r_first[r], a.r_last[r], a.p, &nb, [&](int s, int e, int c) {
        a.w_parent[rs + c] = (uint32_t)(a.read_base + r);
        a.w_start[rs + c] = s;
        a.w_end[rs + c] = e;
        a.w_len[rs + c] = e - s;
    });
}

// B3: per row -- mean and window quality from the mask range [S, E) of the parent read.
// The k-mer-mode quality of a base is exactly 0.0 or 1.0 (read.cpp:42,55), so the mean's sequential
// sum is an exact integer, and the window recurrence adds / subtracts r = 1.0 / ws or 0.0 in the
// reference's order (read.cpp:226-232) -- bit-identical, including its rounding drift.
template <bool RUNS>
__device__ __forceinline__ void kmer_row_stats(const uint32_t *__restrict__ m, int S, int E, int ws, double *mean_out,
                                               double *window_out) {
    const int len = E - S;
    // popcount of [S, E)
    long long hits = 0;
    {
        int p = S;
        while (p < E) {
            uint32_t w = m[p >> 5] & (0xFFFFFFFFu << (p & 31));
            int wend = (p & ~31) + 32;
            if (wend > E) w &= 0xFFFFFFFFu >> (wend - E);
            hits += __popc(w);
            p = wend;
        }
    }
    const double mean = 100.0 * (double)hits / (double)len;             // read.cpp:208-213
    if (len <= ws) {                                                     // read.cpp:217-218
        *mean_out = mean;
        *window_out = mean;
        return;
    }
    const double wsd = (double)ws;
    const double rq = 1.0 / wsd;                                         // qualities[i] / window_size with q == 1.0
    long long c0 = 0;
    {
        int p = S;
        const int e0 = S + ws;
        while (p < e0) {
            uint32_t w = m[p >> 5] & (0xFFFFFFFFu << (p & 31));
            int wend = (p & ~31) + 32;
            if (wend > e0) w &= 0xFFFFFFFFu >> (wend - e0);
            c0 += __popc(w);
            p = wend;
        }
    }
    double w = (double)c0 / wsd;                                         // read.cpp:220-223
    double best = w;
    int pin = S + ws, pout = S;
    // scalar steps until the incoming position is word aligned
    auto step = [&](unsigned bin, unsigned bout) {
        w -= bout ? rq : 0.0;                                            // read.cpp:229
        w += bin ? rq : 0.0;                                             // read.cpp:230
        if (w < best) best = w;
    };
    for (; (pin & 31) && pin < E; ++pin, ++pout)
        step((m[pin >> 5] >> (pin & 31)) & 1u, (m[pout >> 5] >> (pout & 31)) & 1u);
    if (pin + 32 <= E) {
        const unsigned sh = (unsigned)(pout & 31);       // constant from here on
        int ow_idx = pout >> 5;
        uint32_t olo = m[ow_idx];
        for (; pin + 32 <= E; pin += 32, pout += 32) {
            const uint32_t inw = m[pin >> 5];
            const uint32_t ohi = m[ow_idx + 1];          // within the parent's padded mask: pout + 32 + 31 < pin + 32 <= E
            const uint32_t outw = __funnelshift_r(olo, ohi, sh);
            if (!RUNS) {
#pragma unroll
                for (int t = 0; t < 32; ++t) step((inw >> t) & 1u, (outw >> t) & 1u);
            } else {
                // Long rows are the serial tail; they run in k_kmer_stats_long with ONE active lane per warp
                // (no divergence) and walk the word by RUNS: a step with
                // both bits 0 adds and subtracts 0.0 (no change); a step with both bits 1 maps w to
                // g(w) = fl(fl(w - r) + r), and once g(w) == w every further (1,1) step is a no-op
                // too, so g is applied until its fixed point (almost always at once); only steps
                // whose bits differ really move w. Same values as the bit loop, step for step. (With 32
                // rows per warp the data-dependent trip counts diverge and this loses -- measured.)
                const uint32_t diff = inw ^ outw, both = inw & outw;
                int t = 0;
                while (t < 32) {
                    const uint32_t rem = diff >> t;
                    const int nd = rem ? t + (__ffs(rem) - 1) : 32;      // next step whose bits differ
                    if (nd > t) {
                        const uint32_t stretch = (nd - t == 32) ? 0xFFFFFFFFu : (((1u << (nd - t)) - 1u) << t);
                        int c11 = __popc(both & stretch);
                        while (c11-- > 0) {
                            const double w2 = (w - rq) + rq;                 // read.cpp:229-230, both qualities 1
                            if (w2 == w) break;
                            w = w2;
                            if (w < best) best = w;
                        }
                    }
                    if (nd >= 32) break;
                    if ((inw >> nd) & 1u) { w -= 0.0; w += rq; }             // base enters a k-mer match
                    else { w -= rq; w += 0.0; }                              // base leaves one
                    if (w < best) best = w;
                    t = nd + 1;
                }
            }
            olo = ohi;
            ++ow_idx;
        }
    }
    for (; pin < E; ++pin, ++pout)
        step((m[pin >> 5] >> (pin & 31)) & 1u, (m[pout >> 5] >> (pout & 31)) & 1u);
    if (best < 0.5 / wsd) best = 0.0;                                    // read.cpp:233-234
    *mean_out = mean;
    *window_out = 100.0 * best;
}

struct StatArgs {
    const uint32_t *mask;
    const uint64_t *off;          // per batch read
    const uint32_t *order;        // rows in descending length order
    uint32_t n_rows;
    fl_params p;
    // rows (batch-local indexing): parent is a GLOBAL read index; read_base converts to batch-local
    const uint32_t *w_parent;
    const int32_t *w_start, *w_end;
    double *w_mean, *w_window;
    uint8_t *w_passed;
    unsigned long long read_base;
    // when non-null: rows that are whole reads copy the parent's statistics instead of recomputing
    const int32_t *r_nchild;
    const double *r_mean, *r_window;
    const uint8_t *r_passed;
};

// Measured (profiles/r01 launch lists): walking long rows by runs with one lane per warp
// (k_kmer_stats_long) is 3x SLOWER on noisy hit masks than the unrolled bit loop, so every row takes
// the bit loop; the threshold is kept only so the experiment can be repeated.
#define FL_LONG_ROW 0x7FFFFFFF

__device__ __forceinline__ bool stat_row(const StatArgs &a, uint32_t row, bool long_pass) {
    const uint32_t r = (uint32_t)(a.w_parent[row] - a.read_base);
    const int S = a.w_start[row], E = a.w_end[row];
    const bool is_long = (E - S) >= FL_LONG_ROW;
    if (is_long != long_pass) return is_long;
    if (a.r_nchild && a.r_nchild[r] == 0) {
        a.w_mean[row] = a.r_mean[r];
        a.w_window[row] = a.r_window[r];
        a.w_passed[row] = a.r_passed[r];
        return is_long;
    }
    const uint32_t *m = a.mask + (a.off[r] >> 5);
    double mean, window;
    if (long_pass) kmer_row_stats<true>(m, S, E, a.p.window_size, &mean, &window);
    else kmer_row_stats<false>(m, S, E, a.p.window_size, &mean, &window);
    a.w_mean[row] = mean;
    a.w_window[row] = window;
    a.w_passed[row] = fl_hard_cutoffs(a.p, E - S, mean, window);
    return is_long;
}

__global__ void __launch_bounds__(256) k_kmer_stats(StatArgs a) {
    const size_t T = (size_t)gridDim.x * blockDim.x;
    for (size_t it = (size_t)blockIdx.x * blockDim.x + threadIdx.x; it < a.n_rows; it += T) stat_row(a, a.order[it], false);
}

// one warp per long row, lane 0 only: rows are ordered longest first, so the long ones are a prefix
__global__ void __launch_bounds__(32) k_kmer_stats_long(StatArgs a) {
    if (threadIdx.x != 0) return;
    for (size_t it = blockIdx.x; it < a.n_rows; it += gridDim.x)
        if (!stat_row(a, a.order[it], true)) break;
}

__global__ void k_iota(uint32_t *p, uint32_t n, uint32_t base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = base + i;
}

__global__ void k_identity_rows(uint32_t n, const int32_t *__restrict__ len, uint32_t *w_parent, int32_t *w_start,
                                int32_t *w_end, unsigned long long read_base) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    w_parent[i] = (uint32_t)(read_base + i);
    w_start[i] = 0;
    w_end[i] = len[i];
}

template <typename T>
cudaError_t grow(DevVec<T> &v, size_t n, size_t keep, cudaStream_t s) { return v.reserve(n, keep, s); }

}  // namespace

int fl_reserve_reads(fl_ctx *c, size_t n_total) {
    size_t k = c->n_reads;
    cudaStream_t s = c->stream;
    FL_CUDA(c, grow(c->r_len, n_total, k, s));
    FL_CUDA(c, grow(c->r_first, n_total, k, s));
    FL_CUDA(c, grow(c->r_last, n_total, k, s));
    FL_CUDA(c, grow(c->r_nbad, n_total, k, s));
    FL_CUDA(c, grow(c->r_nchild, n_total, k, s));
    FL_CUDA(c, grow(c->r_mean, n_total, k, s));
    FL_CUDA(c, grow(c->r_window, n_total, k, s));
    FL_CUDA(c, grow(c->r_passed, n_total, k, s));
    FL_CUDA(c, grow(c->r_rowstart, n_total, k, s));
    return FL_OK;
}

int fl_reserve_rows(fl_ctx *c, size_t n_total) {
    size_t k = c->n_rows;
    cudaStream_t s = c->stream;
    FL_CUDA(c, grow(c->w_parent, n_total, k, s));
    FL_CUDA(c, grow(c->w_start, n_total, k, s));
    FL_CUDA(c, grow(c->w_end, n_total, k, s));
    FL_CUDA(c, grow(c->w_mean, n_total, k, s));
    FL_CUDA(c, grow(c->w_window, n_total, k, s));
    FL_CUDA(c, grow(c->w_passed, n_total, k, s));
    return FL_OK;
}

static int score_kmer(fl_ctx *ctx, const BatchView &b) {
    if (!b.seq2b) { ctx->set_error("k-mer scoring needs seq2b"); return FL_EINVAL; }
    const size_t n = b.n;
    cudaStream_t st = ctx->stream;
    FL_TRY(fl_reserve_reads(ctx, ctx->n_reads + n));
    // ---- kernel A: probe + paint ----
    FL_CUDA(ctx, ctx->sc_mask.reserve((size_t)(b.padded_bases >> 5) + 1, 0, st));
    FL_CUDA(ctx, ctx->sc_u64a.reserve(n + 1, 0, st));
    k_tiles_of<<<fl_blocks(n, 256), 256, 0, st>>>(b.len, b.n, ctx->sc_u64a.p);
    ctx->launches++;
    FL_TRY(fl_exclusive_scan_u64(ctx, ctx->sc_u64a.p, ctx->sc_u64a.p, n, ctx->d_scalars));
    FL_CUDA(ctx, cudaMemcpyAsync(ctx->sc_u64a.p + n, ctx->d_scalars, sizeof(unsigned long long), cudaMemcpyDeviceToDevice, st));
    FL_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
    FL_CUDA(ctx, cudaStreamSynchronize(st));
    const unsigned long long n_tiles = ctx->h_scalars[0];
    if (n_tiles) {
        ProbeArgs pa{};
        pa.seq2b = b.seq2b; pa.off = b.off; pa.len = b.len; pa.tile_start = ctx->sc_u64a.p;
        pa.n = b.n; pa.n_tiles = n_tiles; pa.bitmap = ctx->d_bitmap; pa.anchor = ctx->d_anchor; pa.mask = ctx->sc_mask.p;
        pa.filter = ctx->d_filter; pa.filter_log2_words = ctx->filter_log2_words; pa.filter_kind = ctx->filter_kind;
        unsigned blocks = (unsigned)((n_tiles + 7) / 8);
        unsigned max_blocks = (unsigned)ctx->sm_count * 4;
        if (blocks > max_blocks) blocks = max_blocks;
        {
            KernelTimer kt(ctx, FL_KERNEL_PROBE_PAINT);
            // keep the pre-filter resident in the L2 set-aside while the probe kernel streams reads past it
            const bool persist = ctx->use_filter && ctx->l2_persist_bytes > 0;
            if (persist) {
                cudaStreamAttrValue attr{};
                attr.accessPolicyWindow.base_ptr = ctx->d_filter;
                size_t bytes = ((size_t)1 << ctx->filter_log2_words) * sizeof(unsigned long long);
                attr.accessPolicyWindow.num_bytes = bytes < ctx->l2_window_max ? bytes : ctx->l2_window_max;
                attr.accessPolicyWindow.hitRatio = bytes <= ctx->l2_persist_bytes ? 1.0f : (float)ctx->l2_persist_bytes / (float)bytes;
                attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
                attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
                FL_CUDA(ctx, cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
            }
            if (ctx->use_anchor) {
                if (ctx->use_filter) k_probe_paint<2, true, true><<<blocks, 256, 0, st>>>(pa);
                else k_probe_paint<2, false, true><<<blocks, 256, 0, st>>>(pa);
            } else if (ctx->use_filter) {
                switch (ctx->probe_mode) {
                    case 0: k_probe_paint<0, true, false><<<blocks, 256, 0, st>>>(pa); break;
                    case 2: k_probe_paint<2, true, false><<<blocks, 256, 0, st>>>(pa); break;
                    case 3: k_probe_paint<3, true, false><<<blocks, 256, 0, st>>>(pa); break;
                    default: k_probe_paint<1, true, false><<<blocks, 256, 0, st>>>(pa); break;
                }
            } else {
                switch (ctx->probe_mode) {
                    case 0: k_probe_paint<0, false, false><<<blocks, 256, 0, st>>>(pa); break;
                    case 2: k_probe_paint<2, false, false><<<blocks, 256, 0, st>>>(pa); break;
                    case 3: k_probe_paint<3, false, false><<<blocks, 256, 0, st>>>(pa); break;
                    default: k_probe_paint<1, false, false><<<blocks, 256, 0, st>>>(pa); break;
                }
            }
            if (persist) {
                cudaStreamAttrValue attr{};
                attr.accessPolicyWindow.num_bytes = 0;
                FL_CUDA(ctx, cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr));
            }
        }
        ctx->launches++;
        FL_CUDA(ctx, cudaGetLastError());
    }
    // ---- B1: ranges ----
    const size_t rb = ctx->n_reads, wb = ctx->n_rows;
    FL_CUDA(ctx, ctx->sc_u64b.reserve(n + 1, 0, st));
    KmerArgs ka{};
    ka.mask = ctx->sc_mask.p; ka.off = b.off; ka.len = b.len; ka.n = b.n; ka.p = ctx->p;
    ka.r_len = ctx->r_len.p + rb; ka.r_first = ctx->r_first.p + rb; ka.r_last = ctx->r_last.p + rb;
    ka.r_nbad = ctx->r_nbad.p + rb; ka.r_nchild = ctx->r_nchild.p + rb;
    ka.r_mean = ctx->r_mean.p + rb; ka.r_window = ctx->r_window.p + rb; ka.r_passed = ctx->r_passed.p + rb;
    ka.r_rowstart = ctx->r_rowstart.p + rb; ka.rows_per_read = ctx->sc_u64b.p;
    ka.read_base = rb; ka.row_base = wb;
    k_kmer_ranges<<<fl_blocks(n, 256), 256, 0, st>>>(ka);
    ctx->launches++;
    // ---- B3 on the parents (their own raw mean / window / passed: read.cpp:60-73) ----
    FL_CUDA(ctx, ctx->sc_order.reserve(n, 0, st));
    FL_TRY(fl_order_by_length(ctx, b.len, n, ctx->sc_order.p));
    FL_CUDA(ctx, ctx->sc_u32a.reserve(n, 0, st));
    k_iota<<<fl_blocks(n, 256), 256, 0, st>>>(ctx->sc_u32a.p, b.n, (uint32_t)rb);   // parent of "row" i is read rb + i
    ctx->launches++;
    const unsigned stat_blocks_max = (unsigned)ctx->sm_count * 8;
    {
        StatArgs sa{};
        sa.mask = ctx->sc_mask.p; sa.off = b.off; sa.order = ctx->sc_order.p; sa.n_rows = b.n; sa.p = ctx->p;
        sa.w_parent = ctx->sc_u32a.p;
        // whole-read ranges: start 0, end len -> reuse r_len as "end" and a zero array for start
        FL_CUDA(ctx, ctx->sc_u64c.reserve((n + 1) / 2 + 1, 0, st));
        int32_t *zeros = reinterpret_cast<int32_t *>(ctx->sc_u64c.p);
        FL_CUDA(ctx, cudaMemsetAsync(zeros, 0, n * sizeof(int32_t), st));
        sa.w_start = zeros; sa.w_end = b.len;
        sa.w_mean = ctx->r_mean.p + rb; sa.w_window = ctx->r_window.p + rb; sa.w_passed = ctx->r_passed.p + rb;
        sa.read_base = rb;
        unsigned blocks = fl_blocks(n, 256);
        if (blocks > stat_blocks_max) blocks = stat_blocks_max;
        {
            KernelTimer kt(ctx, FL_KERNEL_KMER_STATS);
            k_kmer_stats<<<blocks, 256, 0, st>>>(sa);
        }
        ctx->launches++;
    }
    // ---- rows ----
    const bool may_have_children = ctx->p.trim || ctx->p.split_set;
    size_t n_rows_batch = n;
    if (may_have_children) {
        FL_TRY(fl_exclusive_scan_u64(ctx, ctx->sc_u64b.p, ctx->r_rowstart.p + rb, n, ctx->d_scalars));
        FL_CUDA(ctx, cudaMemcpyAsync(ctx->h_scalars, ctx->d_scalars, sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
        FL_CUDA(ctx, cudaStreamSynchronize(st));
        n_rows_batch = (size_t)ctx->h_scalars[0];
    }
    FL_TRY(fl_reserve_rows(ctx, ctx->n_rows + n_rows_batch));
    if (!may_have_children) {
        k_identity_rows<<<fl_blocks(n, 256), 256, 0, st>>>(b.n, b.len, ctx->w_parent.p + wb, ctx->w_start.p + wb,
                                                          ctx->w_end.p + wb, rb);
        ctx->launches++;
        FL_CUDA(ctx, cudaMemcpyAsync(ctx->w_mean.p + wb, ctx->r_mean.p + rb, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        FL_CUDA(ctx, cudaMemcpyAsync(ctx->w_window.p + wb, ctx->r_window.p + rb, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
        FL_CUDA(ctx, cudaMemcpyAsync(ctx->w_passed.p + wb, ctx->r_passed.p + rb, n, cudaMemcpyDeviceToDevice, st));
    } else {
        RowArgs ra{};
        ra.mask = ctx->sc_mask.p; ra.off = b.off; ra.len = b.len; ra.n = b.n; ra.p = ctx->p;
        ra.r_first = ctx->r_first.p + rb; ra.r_last = ctx->r_last.p + rb; ra.r_nchild = ctx->r_nchild.p + rb;
        ra.r_rowstart = ctx->r_rowstart.p + rb;
        ra.w_parent = ctx->w_parent.p + wb; ra.w_start = ctx->w_start.p + wb; ra.w_end = ctx->w_end.p + wb;
        FL_CUDA(ctx, ctx->sc_u64c.reserve((n_rows_batch + 1) / 2 + 1, 0, st));
        ra.w_len = reinterpret_cast<int32_t *>(ctx->sc_u64c.p);
        ra.read_base = rb; ra.row_base = wb;
        k_kmer_rows<<<fl_blocks(n, 256), 256, 0, st>>>(ra);
        ctx->launches++;
        FL_CUDA(ctx, ctx->sc_order.reserve(n_rows_batch, 0, st));
        FL_TRY(fl_order_by_length(ctx, ra.w_len, n_rows_batch, ctx->sc_order.p));
        StatArgs sa{};
        sa.mask = ctx->sc_mask.p; sa.off = b.off; sa.order = ctx->sc_order.p; sa.n_rows = (uint32_t)n_rows_batch; sa.p = ctx->p;
        sa.w_parent = ctx->w_parent.p + wb; sa.w_start = ctx->w_start.p + wb; sa.w_end = ctx->w_end.p + wb;
        sa.w_mean = ctx->w_mean.p + wb; sa.w_window = ctx->w_window.p + wb; sa.w_passed = ctx->w_passed.p + wb;
        sa.read_base = rb;
        sa.r_nchild = ctx->r_nchild.p + rb; sa.r_mean = ctx->r_mean.p + rb; sa.r_window = ctx->r_window.p + rb;
        sa.r_passed = ctx->r_passed.p + rb;
        unsigned blocks = fl_blocks(n_rows_batch, 256);
        if (blocks > stat_blocks_max) blocks = stat_blocks_max;
        {
            KernelTimer kt(ctx, FL_KERNEL_KMER_STATS);
            k_kmer_stats<<<blocks, 256, 0, st>>>(sa);
        }
        ctx->launches++;
    }
    FL_CUDA(ctx, cudaGetLastError());
    ctx->n_reads += n;
    ctx->n_rows += n_rows_batch;
    return FL_OK;
}

int fl_score_view(fl_ctx *ctx, const BatchView &b) {
    if (b.n == 0) return FL_OK;
    if (ctx->kmers_count_stale || ctx->multi_pending) FL_TRY(fl_kmers_recount(ctx));
    ctx->finalized = false;
    if (ctx->n_kmers == 0) return fl_score_phred(ctx, b);     // read.cpp:35: kmers->empty()
    return score_kmer(ctx, b);
}

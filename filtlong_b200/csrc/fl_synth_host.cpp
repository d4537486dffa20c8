// filtlong_b200/csrc/fl_synth_host.cpp -- host-side synthetic workload generators (bench / tests only):
// the same integer-only model as the device generators (fl_synth.h), walked sequentially. Compiled into
// libfiltlong_b200.so (the C ABI declares them) AND, on its own, into libflsynth_host.so, which has no
// CUDA inside: bench.py's CPU legs (cpu_baseline, --impl reference) write their sample FASTQ / FASTA
// through that one, so the reference arm never maps the CUDA product.
#include "../../include/filtlong_b200.h"
#include "fl_synth.h"

static inline uint64_t padded_len(int64_t len) {
    return len <= 0 ? 0 : (((uint64_t)len + FL_ALIGN_BASES - 1) & ~(uint64_t)(FL_ALIGN_BASES - 1));
}

extern "C" void fl_synth_qual_host(uint64_t seed, uint32_t n, const uint64_t *off, const int32_t *len, const uint8_t *qbar,
                                   uint64_t read_index_base, uint8_t *qual) {
    for (uint32_t r = 0; r < n; ++r)
        for (int pos = 0; pos < len[r]; ++pos)
            qual[off[r] + pos] = fl_synth_qchar(seed, read_index_base + r, (unsigned long long)pos, qbar[r]);
}

extern "C" void fl_synth_genome_host(uint64_t seed, uint64_t n_bases, uint32_t *seq2b) {
    const uint64_t words = (n_bases + 15) >> 4;
    for (uint64_t w = 0; w < words; ++w) {
        uint32_t v = fl_synth_genome_word(seed, w);
        uint64_t rem = n_bases - (w << 4);
        if (rem < 16) v &= ~(0xFFFFFFFFu >> (2 * rem));
        seq2b[w] = v;
    }
}

extern "C" void fl_synth_assembly_host(uint64_t seed, uint32_t n_contigs, uint64_t contig_bases, uint32_t n_ppm, uint32_t *seq2b,
                                       uint32_t *nmask) {
    const uint64_t padded = padded_len((int64_t)contig_bases), wpc = padded >> 4;
    for (uint64_t w = 0; w < wpc * n_contigs; ++w) {
        const uint64_t c = w / wpc, base = (w - c * wpc) << 4;
        uint32_t v = 0, nm = 0;
        if (base < contig_bases) {
            const uint64_t rem = contig_bases - base;
            if (fl_synth_is_nrun(seed, (c * padded + base) / FL_SYNTH_NRUN_BASES, n_ppm)) nm = rem >= 16 ? 0xFFFFu : ((1u << rem) - 1u);
            else v = fl_synth_genome_word(seed, w) & (rem >= 16 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * rem)));
        }
        seq2b[w] = v;
        if (nmask) {
            if (w & 1) nmask[w >> 1] |= nm << 16;
            else nmask[w >> 1] = nm;
        }
    }
}

extern "C" void fl_synth_reads_host(uint64_t seed, const uint32_t *genome2b, const fl_synth_reads *d, uint64_t read_index_base,
                                    uint32_t *seq2b) {
    const int indels = (d->flags & FL_SYNTH_INDELS) ? 1 : 0;
    for (uint32_t r = 0; r < d->n; ++r) {
        const int L = d->len[r];
        uint32_t *o = seq2b + (d->off[r] >> 4);
        const int words = (int)(padded_len(L) >> 4);
        for (int w = 0; w < words; ++w) o[w] = 0;
        const unsigned long long span = fl_synth_span(L);
        const int jp = d->junk_pos[r], jl = d->junk_len[r];
        const int a5 = d->adap5 ? d->adap5[r] : 0, a3 = d->adap3 ? d->adap3[r] : 0;
        long long shift = 0;                                  // deletions at positions <= i minus insertions at positions < i
        for (int i = 0; i < L; ++i) {
            const unsigned long long h = fl_hash64(seed, read_index_base + r, (unsigned long long)i);
            const int kind = fl_synth_event(h, d->err_ppm[r], indels);
            if (kind == 3) ++shift;
            const long long t = (long long)i + shift;
            const int rnd = (jl > 0 && i >= jp && i < jp + jl) || i < a5 || i >= L - a3;
            const uint32_t code = fl_synth_read_base(h, kind, genome2b, d->start[r], span, d->strand[r],
                                                     t < 0 ? 0ull : (unsigned long long)t, rnd);
            o[i >> 4] |= code << (30 - 2 * (i & 15));
            if (kind == 2) --shift;
        }
    }
}

extern "C" void fl_synth_ascii_host(uint32_t n, const uint64_t *off, const int32_t *len, const uint32_t *seq2b, const uint32_t *nmask,
                                    uint8_t *ascii) {
    for (uint32_t r = 0; r < n; ++r)
        for (int i = 0; i < len[r]; ++i) {
            const uint64_t b = off[r] + (uint64_t)i;
            const int isn = nmask ? (int)((nmask[b >> 5] >> (b & 31)) & 1u) : 0;
            ascii[b] = isn ? 'N' : (uint8_t)"ACGT"[(seq2b[b >> 4] >> (30 - 2 * (b & 15))) & 3u];
        }
}

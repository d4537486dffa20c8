// filtlong_b200/csrc/fl_synth.h -- the synthetic workloads of SURVEY 8d as integer-only, counter-based
// functions, shared by the device generators (fl_synth.cu, part of libfiltlong_b200.so) and the host
// generators (fl_synth_host.cpp, also built into the tiny libflsynth_host.so that bench.py's CPU legs load,
// so that the reference arm maps nothing of the CUDA product). Identical bits on host and device.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define FL_HD __host__ __device__ static inline
#else
#define FL_HD static inline
#endif

#define FL_SYNTH_GENOME_STREAM 0x47454E4F4D45ull
#define FL_SYNTH_NRUN_STREAM 0x4E52554E53ull
#define FL_SYNTH_EDGE_STREAM 0x45444745ull
#define FL_SYNTH_NRUN_BASES 1024u          // N runs are whole 1024-base blocks of the assembly

FL_HD unsigned long long fl_hash64(unsigned long long seed, unsigned long long a, unsigned long long b) {
    unsigned long long x = seed + a * 0x9E3779B97F4A7C15ull + b * 0xD6E8FEB86659FD93ull;
    x ^= x >> 30; x *= 0xBF58476D1CE4E5B9ull;
    x ^= x >> 27; x *= 0x94D049BB133111EBull;
    x ^= x >> 31;
    return x;
}

// Phred+33 character: clip(qbar + z, 1, 50) + 33 with z ~ N(0, 4^2) from four summed bytes
FL_HD uint8_t fl_synth_qchar(unsigned long long seed, unsigned long long read, unsigned long long pos, int qbar) {
    unsigned long long h = fl_hash64(seed, read, pos);
    int s = (int)(h & 0xFF) + (int)((h >> 8) & 0xFF) + (int)((h >> 16) & 0xFF) + (int)((h >> 24) & 0xFF);   // ~N(510, 147.8^2)
    int z = ((s - 510) * 111 + 2048) >> 12;                                                                  // ~N(0, 4^2)
    int q = qbar + z;
    q = q < 1 ? 1 : (q > 50 ? 50 : q);
    return (uint8_t)(q + 33);
}

// 16 uniform random bases (one 2-bit word) of the genome / assembly
FL_HD uint32_t fl_synth_genome_word(unsigned long long seed, unsigned long long w) {
    return (uint32_t)(fl_hash64(seed, FL_SYNTH_GENOME_STREAM, w) >> 16);
}

// is the 1024-base block `blk` of the assembly a run of N? (n_ppm = fraction of blocks, parts per million)
FL_HD int fl_synth_is_nrun(unsigned long long seed, unsigned long long blk, uint32_t n_ppm) {
    return n_ppm && (fl_hash64(seed, FL_SYNTH_NRUN_STREAM, blk) % 1000000ull) < n_ppm;
}

FL_HD uint32_t fl_synth_genome_code(const uint32_t *g, unsigned long long pos) {
    return (g[pos >> 4] >> (30 - 2 * (pos & 15))) & 3u;
}

// Per-position error events of a long read (per-read error rate e split 50/25/25 into
// substitutions / insertions / deletions). kind: 0 none, 1 substitution, 2 insertion (this output base is
// random and consumes no template base), 3 deletion (one template base is skipped before this output base).
FL_HD int fl_synth_event(unsigned long long h, uint32_t err_ppm, int indels) {
    const unsigned long long thr = ((unsigned long long)err_ppm << 20) / 1000000ull;
    const unsigned long long u = (h >> 8) & 0xFFFFFull;
    if (u >= thr) return 0;
    if (!indels) return 1;
    if (u < (thr >> 1)) return 1;
    if (u < (thr >> 1) + (thr >> 2)) return 2;
    return 3;
}

// template bases a read of `len` output bases may consume (deletions advance the template faster)
FL_HD unsigned long long fl_synth_span(int len) { return (unsigned long long)len + (unsigned long long)(len >> 3) + 64ull; }

// The output base at position i of read `read`, given t = i + (deletions at positions <= i) - (insertions
// at positions < i), the index of its template base. Forward strand: template base t is genome[start + t];
// reverse strand: the read is the reverse complement of genome[start, start + span).
FL_HD uint32_t fl_synth_read_base(unsigned long long h, int kind, const uint32_t *genome, unsigned long long start,
                                  unsigned long long span, int strand, unsigned long long t, int in_random_block) {
    if (in_random_block || kind == 2) return (uint32_t)(h >> 40) & 3u;
    if (t >= span) t = span - 1;
    uint32_t code = strand ? 3u - fl_synth_genome_code(genome, start + (span - 1 - t)) : fl_synth_genome_code(genome, start + t);
    if (kind == 1) code = (code + 1u + (uint32_t)((h >> 32) % 3ull)) & 3u;
    return code;
}

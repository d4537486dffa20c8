// filtlong_b200/csrc/fl_device.cuh -- device helpers shared by the k-mer build and probe kernels.
#pragma once
#include "fl_internal.cuh"

// A warp walks a sequence in steps of 1024 bases: lane l owns bases [32 l, 32 l + 32) of the step.
// Work is cut into tiles of FL_TILE_BASES bases of ONE sequence so that long and short sequences
// load-balance; tile_start[] is the exclusive scan of tiles-per-sequence.
#define FL_STEP_BASES 1024
#define FL_TILE_STEPS 8
#define FL_TILE_BASES (FL_STEP_BASES * FL_TILE_STEPS)

__device__ __forceinline__ unsigned long long fl_tiles_of(int len) {
    return len <= 0 ? 0ull : ((unsigned long long)len + FL_TILE_BASES - 1) / FL_TILE_BASES;
}

// largest i in [0, n) with tile_start[i] <= t   (tile_start is non-decreasing, tile_start[0] == 0)
__device__ __forceinline__ uint32_t fl_find_seq(const unsigned long long *__restrict__ tile_start, uint32_t n,
                                                unsigned long long t) {
    uint32_t lo = 0, hi = n;   // invariant: tile_start[lo] <= t, (hi == n or tile_start[hi] > t)
    while (hi - lo > 1) {
        uint32_t mid = lo + ((hi - lo) >> 1);
        if (tile_start[mid] <= t) lo = mid;
        else hi = mid;
    }
    return lo;
}

// The three 32-bit words a lane needs for its 32 forward 16-mers: w0 = bases 0..15 of its run,
// w1 = bases 16..31, w2 = the first 16 bases after its run (owned by the next lane / next step).
struct LaneWords {
    uint32_t w0, w1, w2;
};

// seqw: the sequence's first word; step_base: first base of this step (multiple of 1024);
// padded_len: padded length of the sequence (words beyond it must not be touched).
__device__ __forceinline__ LaneWords fl_load_lane_words(const uint32_t *__restrict__ seqw, unsigned long long step_base,
                                                        unsigned long long padded_len, unsigned lane) {
    LaneWords r;
    unsigned long long b = step_base + 32ull * lane;
    uint2 v = make_uint2(0u, 0u);
    if (b < padded_len) v = __ldg(reinterpret_cast<const uint2 *>(seqw + (b >> 4)));
    r.w0 = v.x;
    r.w1 = v.y;
    uint32_t nxt = __shfl_down_sync(0xffffffffu, r.w0, 1);
    if (lane == 31) {
        unsigned long long nb = step_base + FL_STEP_BASES;
        nxt = (nb < padded_len) ? __ldg(seqw + (nb >> 4)) : 0u;
    }
    r.w2 = nxt;
    return r;
}

// forward 16-mer starting at position p (0..31) of the lane's run (kmers.cpp:222-229 encoding:
// first base in bits 31:30)
__device__ __forceinline__ uint32_t fl_kmer_at(const LaneWords &w, int p) {
    return p < 16 ? __funnelshift_l(w.w1, w.w0, 2 * p) : __funnelshift_l(w.w2, w.w1, 2 * (p - 16));
}

// reverse the order of the sixteen 2-bit fields of x
__device__ __forceinline__ uint32_t fl_reverse_pairs(uint32_t x) {
    uint32_t b = __brev(x);
    return ((b >> 1) & 0x55555555u) | ((b & 0x55555555u) << 1);
}

// spread the low 16 bits of x to the even bit positions
__device__ __forceinline__ uint32_t fl_spread16(uint32_t x) {
    x &= 0xFFFFu;
    x = (x | (x << 8)) & 0x00FF00FFu;
    x = (x | (x << 4)) & 0x0F0F0F0Fu;
    x = (x | (x << 2)) & 0x33333333u;
    x = (x | (x << 1)) & 0x55555555u;
    return x;
}

// Bloom hash j of a 4-byte key: bloom_filter.h:569-583 specialised (see oracle for derivation)
__device__ __forceinline__ uint32_t fl_bloom_hash(uint32_t kmer, uint32_t salt) {
    return salt ^ ~((salt << 11) + (kmer ^ (salt >> 5)));
}

// hard cut-offs on the RAW qualities, read.cpp:65-73 (else-if chain; each test only if its option is set)
__device__ __forceinline__ uint8_t fl_hard_cutoffs(const fl_params &p, int length, double mean_q, double window_q) {
    if (p.min_length_set && length < p.min_length) return 0;
    else if (p.max_length_set && length > p.max_length) return 0;
    else if (p.min_mean_q_set && mean_q < p.min_mean_q) return 0;
    else if (p.min_window_q_set && window_q < p.min_window_q) return 0;
    return 1;
}

// length buckets for load balancing: 8 per octave, 0..255, 255 = longest
__device__ __forceinline__ unsigned fl_length_bucket(int len) {
    if (len <= 0) return 0;
    unsigned l = (unsigned)len;
    unsigned msb = 31 - __clz(l);
    unsigned frac = msb >= 3 ? ((l >> (msb - 3)) & 7u) : ((l << (3 - msb)) & 7u);
    return msb * 8 + frac;   // <= 31*8+7 = 255
}

// Position-anchored membership table (2 GiB, built from the bitmap by k_anchor_build). A random probe
// of the 512 MiB bitmap costs a whole DRAM burst for one bit, and consecutive 16-mers of a read land in
// unrelated words. But the four 16-mers starting at read positions 4g, 4g+1, 4g+2, 4g+3 all contain the
// 13 bases at positions 4g+3 .. 4g+15: keyed by those 13 bases (26 bits), one 32-byte sector holds
// everything the four probes need -- for each alignment r = 3 - (position & 3) a 64-bit quarter indexed
// by the 3 remaining bases (the r bases before the key and the 3 - r after it). Every member of the set
// is entered once per alignment (4x the bits of the bitmap), and a read's probes touch ONE sector per
// four 16-mers instead of four.
__host__ __device__ __forceinline__ void fl_anchor_slot(uint32_t kmer, unsigned r, uint32_t &word, uint32_t &bit) {
    const unsigned sh = 6u - 2u * r;                                  // bits of the 3 - r bases after the key
    const uint32_t key = (kmer >> sh) & 0x3FFFFFFu;                   // 13 bases
    const uint32_t top = r ? (kmer >> (32u - 2u * r)) : 0u;           // the r bases before the key
    const uint32_t rest = (top << sh) | (kmer & ((1u << sh) - 1u));   // 6 bits
    word = key * 8u + r * 2u + (rest >> 5);
    bit = rest & 31u;
}

// L2-resident pre-filter in front of the membership tables (fl_kmers.cu / fl_score.cu): two bits (kind bit 4: four)
// in ONE 64-bit word of a 2^log2_words-word table. No false negatives, so "filter says absent" is final.
// This is the flavour with one word per 16-mer; the keyed flavours follow.
__device__ __forceinline__ void fl_filter_slot(uint32_t kmer, unsigned log2_words, int kind, uint32_t &word, unsigned long long &bits) {
    const uint32_t h1 = kmer * 0x9E3779B1u;
    const uint32_t h2 = (kmer ^ (kmer >> 15)) * 0x85EBCA6Bu;
    word = h1 >> (32 - log2_words);
    bits = (1ull << (h2 >> 26)) | (1ull << ((h2 >> 20) & 63u));
    if (kind & 16) bits |= (1ull << ((h2 >> 14) & 63u)) | (1ull << ((h2 >> 8) & 63u));   // four bits per member
}

// Group-keyed flavours of the pre-filter (FL_FILTER_KIND bit 2 / bit 3): the WORD is chosen by what a run of consecutive
// 16-mers of a read has in common, so ONE load answers for the whole run, and each 16-mer keeps its own two bits.
//   group of 4 (the anchored table's group): the 13-base key of fl_anchor_slot(kmer, r); a member is inserted four times (r = 0..3)
//   pair: the 15 bases two neighbours share; role 0 = the earlier one (its low 30 bits), role 1 = the later one (its high 30 bits)
__host__ __device__ __forceinline__ unsigned long long fl_filter_bits_role(uint32_t kmer, unsigned role, int kind) {
    const uint32_t h2 = ((kmer ^ (kmer >> 15)) + role * 0x632BE5ABu) * 0x85EBCA6Bu;
    unsigned long long bits = (1ull << (h2 >> 26)) | (1ull << ((h2 >> 20) & 63u));
    if (kind & 16) bits |= (1ull << ((h2 >> 14) & 63u)) | (1ull << ((h2 >> 8) & 63u));
    return bits;
}
__host__ __device__ __forceinline__ uint32_t fl_filter_word_group4(uint32_t kmer, unsigned r, unsigned log2_words) {
    const uint32_t key = (kmer >> (6u - 2u * r)) & 0x3FFFFFFu;
    return (key * 0x9E3779B1u) >> (32 - log2_words);
}
__host__ __device__ __forceinline__ uint32_t fl_filter_word_pair(uint32_t kmer, unsigned role, unsigned log2_words) {
    const uint32_t key = role ? (kmer >> 2) : (kmer & 0x3FFFFFFFu);
    return ((key ^ (key >> 13)) * 0x9E3779B1u) >> (32 - log2_words);
}

// Four characters -> four 2-bit codes (kmers.cpp:176-196: A/a 0, C/c 1, G/g 2, T/t 3, anything else 0), first
// character in bits 7:6 of code8; other4 has bit i set where character i is not one of ACGTacgt.
__device__ __forceinline__ void fl_pack4(uint32_t x, uint32_t &code8, uint32_t &other4) {
    x &= 0xDFDFDFDFu;                                     // fold lower case onto upper case
    const uint32_t mA = __vcmpeq4(x, 0x41414141u), mC = __vcmpeq4(x, 0x43434343u), mG = __vcmpeq4(x, 0x47474747u),
                   mT = __vcmpeq4(x, 0x54545454u);
    const uint32_t v = ((mC | mT) & 0x01010101u) | ((mG | mT) & 0x02020202u);      // 2-bit code in every byte
    code8 = (v * 0x40100401u) >> 24;                       // first character in bits 7:6 (no carries: fields never overlap)
    const uint32_t o = ~(mA | mC | mG | mT) & 0x01010101u;
    other4 = (o | (o >> 7) | (o >> 14) | (o >> 21)) & 0xFu; // first character in bit 0
}


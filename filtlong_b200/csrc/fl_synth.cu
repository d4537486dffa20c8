// filtlong_b200/csrc/fl_synth.cu -- device-side synthetic workload generators (bench / tests only).
// The model itself lives in fl_synth.h (shared with the host generators of fl_synth_host.cpp).
#include "fl_device.cuh"
#include "fl_synth.h"

namespace {

__global__ void k_synth_qual(unsigned long long seed, uint32_t n, const uint64_t *__restrict__ off, const int32_t *__restrict__ len,
                             const uint8_t *__restrict__ qbar, unsigned long long read_base, uint8_t *__restrict__ qual) {
    // one warp per read, lanes stride over 16-byte groups
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = warp; r < n; r += n_warps) {
        const int L = len[r];
        const int qb = qbar[r];
        uint8_t *q = qual + off[r];
        const int groups = (L + 15) >> 4;
        for (int g = lane; g < groups; g += 32) {
            uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                int pos = g * 16 + i;
                uint8_t ch = pos < L ? fl_synth_qchar(seed, read_base + r, (unsigned long long)pos, qb) : 0;
                w[i >> 2] |= (uint32_t)ch << (8 * (i & 3));
            }
            reinterpret_cast<uint4 *>(q)[g] = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
}

__global__ void k_synth_genome(unsigned long long seed, unsigned long long n_bases, uint32_t *__restrict__ out) {
    const unsigned long long words = (n_bases + 15) >> 4;
    for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < words;
         w += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t v = fl_synth_genome_word(seed, w);
        unsigned long long rem = n_bases - (w << 4);
        if (rem < 16) v &= ~(0xFFFFFFFFu >> (2 * rem));     // bases beyond the end stay 0
        out[w] = v;
    }
}

// contigs at padded offsets; word w of the arena belongs to contig w / words_per_contig. One thread per
// PAIR of sequence words (= one 32-bit word of the non-ACGT mask).
__device__ __forceinline__ void assembly_word(unsigned long long seed, unsigned long long w, unsigned long long wpc,
                                              unsigned long long contig_bases, unsigned long long padded, uint32_t n_ppm,
                                              uint32_t &v, uint32_t &nm) {
    const unsigned long long c = w / wpc, base = (w - c * wpc) << 4;
    v = 0;
    nm = 0;
    if (base >= contig_bases) return;
    const unsigned long long rem = contig_bases - base;
    if (fl_synth_is_nrun(seed, (c * padded + base) / FL_SYNTH_NRUN_BASES, n_ppm)) nm = rem >= 16 ? 0xFFFFu : ((1u << rem) - 1u);
    else v = fl_synth_genome_word(seed, w) & (rem >= 16 ? 0xFFFFFFFFu : ~(0xFFFFFFFFu >> (2 * rem)));
}

__global__ void k_synth_assembly(unsigned long long seed, uint32_t n_contigs, unsigned long long contig_bases,
                                 unsigned long long padded, uint32_t n_ppm, uint32_t *__restrict__ out, uint32_t *__restrict__ nmask) {
    const unsigned long long wpc = padded >> 4, pairs = (wpc * n_contigs) >> 1;     // wpc is a multiple of 4
    for (unsigned long long m = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; m < pairs;
         m += (unsigned long long)gridDim.x * blockDim.x) {
        uint32_t v0, v1, n0, n1;
        assembly_word(seed, 2 * m, wpc, contig_bases, padded, n_ppm, v0, n0);
        assembly_word(seed, 2 * m + 1, wpc, contig_bases, padded, n_ppm, v1, n1);
        reinterpret_cast<uint2 *>(out)[m] = make_uint2(v0, v1);
        if (nmask) nmask[m] = n0 | (n1 << 16);
    }
}

// One warp per read, 1024 output positions per step, lane l owns positions [32 l, 32 l + 32) of the step.
// The template index of an output base needs the running (deletions - insertions) count: a popcount
// prefix inside the lane, a warp scan over the lanes and a carry across steps.
__global__ void __launch_bounds__(256) k_synth_reads(unsigned long long seed, const uint32_t *__restrict__ genome, fl_synth_reads d,
                                                     unsigned long long read_base, uint32_t *__restrict__ out) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    const int indels = (d.flags & FL_SYNTH_INDELS) ? 1 : 0;
    for (size_t r = warp; r < d.n; r += n_warps) {
        const int L = d.len[r];
        if (L <= 0) continue;
        const unsigned long long st = d.start[r], span = fl_synth_span(L);
        const int strand = d.strand[r];
        const uint32_t err = d.err_ppm[r];
        const int jp = d.junk_pos[r], jl = d.junk_len[r];
        const int a5 = d.adap5 ? d.adap5[r] : 0, a3 = d.adap3 ? d.adap3[r] : 0;
        uint32_t *o = out + (d.off[r] >> 4);
        const int padded = (int)(((unsigned)L + 63u) & ~63u);
        long long carry = 0;                                   // deletions - insertions before this step
        for (int sb = 0; sb < padded; sb += 1024) {
            const int lb = sb + 32 * (int)lane;
            uint32_t insm = 0, delm = 0, subm = 0;
            // events of the lane's 32 positions
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const int i = lb + k;
                if (i < L) {
                    const int kind = fl_synth_event(fl_hash64(seed, read_base + r, (unsigned long long)i), err, indels);
                    insm |= (kind == 2 ? 1u : 0u) << k;
                    delm |= (kind == 3 ? 1u : 0u) << k;
                    subm |= (kind == 1 ? 1u : 0u) << k;
                }
            }
            int delta = __popc(delm) - __popc(insm), incl = delta;
#pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= (unsigned)s) incl += t;
            }
            const long long lane_base = carry + (incl - delta);
            uint32_t w0 = 0, w1 = 0;
#pragma unroll 4
            for (int k = 0; k < 32; ++k) {
                const int i = lb + k;
                if (i < L) {
                    const uint32_t below = k ? (0xFFFFFFFFu >> (32 - k)) : 0u;
                    // deletions at positions <= i, insertions at positions < i
                    const long long t = (long long)i + lane_base + __popc(delm & (below | (1u << k))) - __popc(insm & below);
                    const int kind = ((insm >> k) & 1u) ? 2 : (((delm >> k) & 1u) ? 3 : (((subm >> k) & 1u) ? 1 : 0));
                    const int rnd = (jl > 0 && i >= jp && i < jp + jl) || i < a5 || i >= L - a3;
                    const unsigned long long h = fl_hash64(seed, read_base + r, (unsigned long long)i);
                    const uint32_t code = fl_synth_read_base(h, kind, genome, st, span, strand, t < 0 ? 0ull : (unsigned long long)t, rnd);
                    if (k < 16) w0 |= code << (30 - 2 * k);
                    else w1 |= code << (30 - 2 * (k - 16));
                }
            }
            if (lb < padded) {
                o[lb >> 4] = w0;
                o[(lb >> 4) + 1] = w1;
            }
            carry += __shfl_sync(0xffffffffu, incl, 31);
        }
    }
}

__global__ void __launch_bounds__(256) k_synth_ascii(uint32_t n, const uint64_t *__restrict__ off, const int32_t *__restrict__ len,
                                                     const uint32_t *__restrict__ seq2b, const uint32_t *__restrict__ nmask,
                                                     uint8_t *__restrict__ ascii) {
    const unsigned lane = threadIdx.x & 31;
    const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((size_t)gridDim.x * blockDim.x) >> 5;
    for (size_t r = warp; r < n; r += n_warps) {
        const int L = len[r];
        const unsigned long long o = off[r];
        const int words = L > 0 ? (int)((((unsigned)L + 63u) & ~63u) >> 4) : 0;
        for (int w = lane; w < words; w += 32) {
            const uint32_t v = seq2b[(o >> 4) + w];
            const uint32_t nm = nmask ? (nmask[(o >> 5) + (w >> 1)] >> (16 * (w & 1))) & 0xFFFFu : 0u;
            uint32_t q[4] = {0, 0, 0, 0};
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int i = w * 16 + k;
                uint32_t ch = 0;
                if (i < L) ch = ((nm >> k) & 1u) ? 'N' : (0x54474341u >> (8 * ((v >> (30 - 2 * k)) & 3u))) & 0xFFu;   // "ACGT"
                q[k >> 2] |= ch << (8 * (k & 3));
            }
            reinterpret_cast<uint4 *>(ascii + o)[w] = make_uint4(q[0], q[1], q[2], q[3]);
        }
    }
}

}  // namespace

extern "C" int fl_synth_qual_device(fl_ctx *c, uint64_t seed, uint32_t n, const uint64_t *dev_off, const int32_t *dev_len,
                                    const uint8_t *dev_qbar, uint64_t read_index_base, uint8_t *dev_qual) {
    if (!c) return FL_EINVAL;
    if (n == 0) return FL_OK;
    FL_CUDA(c, cudaSetDevice(c->device));
    k_synth_qual<<<c->sm_count * 8, 256, 0, c->stream>>>(seed, n, dev_off, dev_len, dev_qbar, read_index_base, dev_qual);
    FL_CUDA(c, cudaGetLastError());
    return FL_OK;
}

extern "C" int fl_synth_genome_device(fl_ctx *c, uint64_t seed, uint64_t n_bases, uint32_t *dev_seq2b) {
    if (!c) return FL_EINVAL;
    FL_CUDA(c, cudaSetDevice(c->device));
    k_synth_genome<<<c->sm_count * 8, 256, 0, c->stream>>>(seed, n_bases, dev_seq2b);
    FL_CUDA(c, cudaGetLastError());
    return FL_OK;
}

extern "C" int fl_synth_assembly_device(fl_ctx *c, uint64_t seed, uint32_t n_contigs, uint64_t contig_bases, uint32_t n_ppm,
                                        uint32_t *dev_seq2b, uint32_t *dev_nmask) {
    if (!c || !dev_seq2b) return FL_EINVAL;
    if (n_contigs == 0 || contig_bases == 0) return FL_OK;
    FL_CUDA(c, cudaSetDevice(c->device));
    k_synth_assembly<<<c->sm_count * 8, 256, 0, c->stream>>>(seed, n_contigs, contig_bases, fl_padded_len((int64_t)contig_bases), n_ppm,
                                                             dev_seq2b, dev_nmask);
    FL_CUDA(c, cudaGetLastError());
    return FL_OK;
}

extern "C" int fl_synth_reads_device(fl_ctx *c, uint64_t seed, const uint32_t *dev_genome2b, const fl_synth_reads *dev_desc,
                                     uint64_t read_index_base, uint32_t *dev_seq2b) {
    if (!c || !dev_desc) return FL_EINVAL;
    if (dev_desc->n == 0) return FL_OK;
    FL_CUDA(c, cudaSetDevice(c->device));
    k_synth_reads<<<c->sm_count * 8, 256, 0, c->stream>>>(seed, dev_genome2b, *dev_desc, read_index_base, dev_seq2b);
    FL_CUDA(c, cudaGetLastError());
    return FL_OK;
}

extern "C" int fl_synth_ascii_device(fl_ctx *c, uint32_t n, const uint64_t *dev_off, const int32_t *dev_len, const uint32_t *dev_seq2b,
                                     const uint32_t *dev_nmask, uint8_t *dev_ascii) {
    if (!c || !dev_seq2b || !dev_ascii) return FL_EINVAL;
    if (n == 0) return FL_OK;
    FL_CUDA(c, cudaSetDevice(c->device));
    k_synth_ascii<<<c->sm_count * 8, 256, 0, c->stream>>>(n, dev_off, dev_len, dev_seq2b, dev_nmask, dev_ascii);
    FL_CUDA(c, cudaGetLastError());
    return FL_OK;
}

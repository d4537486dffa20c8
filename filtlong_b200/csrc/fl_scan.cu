// filtlong_b200/csrc/fl_scan.cu -- device-wide exclusive scan (uint64) and the length-bucket
// ordering used to balance thread-per-row kernels. Hand-written; n is at most ~10^8.
#include "fl_device.cuh"

namespace {

constexpr int SCAN_THREADS = 1024;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v) {
    const unsigned lane = threadIdx.x & 31;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= (unsigned)d) v += o;
    }
    return v;
}

// exclusive scan of one value per thread across the block; returns the exclusive prefix and the
// block total through *total
__device__ __forceinline__ unsigned long long block_excl_scan(unsigned long long v, unsigned long long *total) {
    __shared__ unsigned long long warp_sums[32];
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    unsigned long long incl = warp_incl_scan(v);
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        unsigned long long s = (lane < (blockDim.x >> 5)) ? warp_sums[lane] : 0ull;
        unsigned long long si = warp_incl_scan(s);
        warp_sums[lane] = si - s;
        if (lane == 31) *total = si;
    }
    __syncthreads();
    unsigned long long r = incl - v + warp_sums[wid];
    __syncthreads();
    return r;
}

// in == out is allowed (callers scan in place), so neither pointer is __restrict__
__global__ void __launch_bounds__(SCAN_THREADS) k_scan_tiles(const unsigned long long *in, unsigned long long *out, size_t n,
                                                            unsigned long long *__restrict__ tile_sums) {
    __shared__ unsigned long long total;
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS], s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        v[i] = (base + i < n) ? in[base + i] : 0ull;
        s += v[i];
    }
    unsigned long long ex = block_excl_scan(s, &total);
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_sums(unsigned long long *__restrict__ tile_sums, size_t n_tiles,
                                                           unsigned long long *__restrict__ total_out) {
    __shared__ unsigned long long total;
    unsigned long long carry = 0;
    for (size_t base = 0; base < n_tiles; base += SCAN_THREADS) {
        size_t i = base + threadIdx.x;
        unsigned long long v = (i < n_tiles) ? tile_sums[i] : 0ull;
        unsigned long long ex = block_excl_scan(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_add(unsigned long long *__restrict__ out, size_t n,
                                                          const unsigned long long *__restrict__ tile_sums) {
    size_t base = (size_t)blockIdx.x * SCAN_TILE + (size_t)threadIdx.x * SCAN_ITEMS;
    unsigned long long add = tile_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_ITEMS; ++i)
        if (base + i < n) out[base + i] += add;
}

// ---- bucket ordering: rows sorted by DESCENDING bucket key (counting sort, FL_ORDER_BUCKETS keys) ----
struct LenKey {
    const int32_t *len;
    __device__ __forceinline__ unsigned operator()(size_t i) const { return fl_length_bucket(len[i]); }
};
struct RawKey {
    const uint32_t *key;
    __device__ __forceinline__ unsigned operator()(size_t i) const { return key[i] < FL_ORDER_BUCKETS ? key[i] : FL_ORDER_BUCKETS - 1; }
};

template <typename K>
__global__ void k_bucket_hist(K keyfn, size_t n, uint32_t *__restrict__ buckets) {
    __shared__ uint32_t h[FL_ORDER_BUCKETS];
    for (int i = threadIdx.x; i < FL_ORDER_BUCKETS; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        atomicAdd(&h[keyfn(i)], 1u);
    __syncthreads();
    for (int i = threadIdx.x; i < FL_ORDER_BUCKETS; i += blockDim.x)
        if (h[i]) atomicAdd(&buckets[i], h[i]);
}

// cursors[b] = number of rows in higher buckets (descending order)
__global__ void k_bucket_scan(uint32_t *__restrict__ buckets) {
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int b = FL_ORDER_BUCKETS - 1; b >= 0; --b) {
            uint32_t c = buckets[b];
            buckets[FL_ORDER_BUCKETS + b] = run;
            run += c;
        }
    }
}

template <typename K>
__global__ void k_bucket_scatter(K keyfn, size_t n, uint32_t *__restrict__ buckets, uint32_t *__restrict__ order) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        unsigned b = keyfn(i);
        // warp-aggregated cursor bump per bucket
        unsigned peers = __match_any_sync(__activemask(), b);
        unsigned leader = __ffs(peers) - 1;
        unsigned lane = threadIdx.x & 31;
        uint32_t basepos = 0;
        if (lane == leader) basepos = atomicAdd(&buckets[FL_ORDER_BUCKETS + b], (uint32_t)__popc(peers));
        basepos = __shfl_sync(peers, basepos, leader);
        order[basepos + __popc(peers & ((1u << lane) - 1))] = (uint32_t)i;
    }
}

template <typename K>
int order_by(fl_ctx *ctx, K keyfn, size_t n, uint32_t *order) {
    if (n == 0) return FL_OK;
    if (!ctx->d_buckets) FL_CUDA(ctx, cudaMalloc(&ctx->d_buckets, 2 * FL_ORDER_BUCKETS * sizeof(uint32_t)));
    FL_CUDA(ctx, cudaMemsetAsync(ctx->d_buckets, 0, 2 * FL_ORDER_BUCKETS * sizeof(uint32_t), ctx->stream));
    unsigned blocks = fl_blocks(n, 256);
    if (blocks > (unsigned)ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_bucket_hist<<<blocks, 256, 0, ctx->stream>>>(keyfn, n, ctx->d_buckets);
    k_bucket_scan<<<1, 32, 0, ctx->stream>>>(ctx->d_buckets);
    k_bucket_scatter<<<blocks, 256, 0, ctx->stream>>>(keyfn, n, ctx->d_buckets, order);
    ctx->launches += 3;
    FL_CUDA(ctx, cudaGetLastError());
    return FL_OK;
}

}  // namespace

int fl_exclusive_scan_u64(fl_ctx *ctx, const unsigned long long *in, unsigned long long *out, size_t n,
                          unsigned long long *total_dev) {
    if (n == 0) {
        if (total_dev) FL_CUDA(ctx, cudaMemsetAsync(total_dev, 0, sizeof(unsigned long long), ctx->stream));
        return FL_OK;
    }
    size_t tiles = (n + SCAN_TILE - 1) / SCAN_TILE;
    FL_CUDA(ctx, ctx->sc_scan.reserve(tiles, 0, ctx->stream));
    k_scan_tiles<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(in, out, n, ctx->sc_scan.p);
    k_scan_sums<<<1, SCAN_THREADS, 0, ctx->stream>>>(ctx->sc_scan.p, tiles, total_dev);
    k_scan_add<<<(unsigned)tiles, SCAN_THREADS, 0, ctx->stream>>>(out, n, ctx->sc_scan.p);
    ctx->launches += 3;
    FL_CUDA(ctx, cudaGetLastError());
    return FL_OK;
}

int fl_order_by_length(fl_ctx *ctx, const int32_t *len, size_t n, uint32_t *order) { return order_by(ctx, LenKey{len}, n, order); }

int fl_order_by_key(fl_ctx *ctx, const uint32_t *key, size_t n, uint32_t *order) { return order_by(ctx, RawKey{key}, n, order); }

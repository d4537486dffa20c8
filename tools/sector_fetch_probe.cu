// How many DRAM bytes does one random 32-byte sector load cost on this GPU, by flavour of the load?
// (k_probe_paint's table loads: profiles/r02_probe_filter_experiments.md measured ~122 B per 32-B sector.)
// Each kernel reads N random 32-byte sectors of a 2 GiB table; run under
//   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum --clock-control none
// and divide. Measurement tool only; not part of the library.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o sector_fetch_probe sector_fetch_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t mix(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

template <int V>
__device__ __forceinline__ uint32_t load32B(const uint32_t *p) {
    uint32_t r[8];
    if (V == 0) asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 1) asm volatile("ld.global.nc.L1::no_allocate.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 2) asm volatile("ld.global.nc.L1::no_allocate.L2::128B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 3) asm volatile("ld.global.nc.L1::no_allocate.L2::256B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 4) asm volatile("ld.global.cg.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 5) asm volatile("ld.global.cg.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 6) asm volatile("ld.global.cv.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p));
    if (V == 7) {   // two 16-byte halves
        asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "l"(p));
        asm volatile("ld.global.nc.L1::no_allocate.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p + 4));
    }
    if (V == 8) {   // one 4-byte word only
        asm volatile("ld.global.nc.L1::no_allocate.b32 %0, [%1];" : "=r"(r[0]) : "l"(p));
        r[1] = r[2] = r[3] = r[4] = r[5] = r[6] = r[7] = 0;
    }
    if (V == 9) {   // L2 eviction policy: no allocate-ish (evict_first) + 64B
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.L2::64B.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "l"(p), "l"(pol));
    }
    return r[0] ^ r[1] ^ r[2] ^ r[3] ^ r[4] ^ r[5] ^ r[6] ^ r[7];
}

// V = load flavour; every thread does 8 independent loads per round like the probe kernel's groups
template <int V>
__global__ void __launch_bounds__(256, 4) k_fetch(const uint32_t *__restrict__ table, unsigned log2_sectors, unsigned rounds, uint32_t *out) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (unsigned r = 0; r < rounds; ++r) {
        uint32_t v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t s = mix(tid * 977u + r * 8u + i + 0x9e3779b9u * (r + 1)) >> (32 - log2_sectors);
            v[i] = load32B<V>(table + (size_t)s * 8u);
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) acc ^= v[i];
    }
    if (acc == 0x12345u) out[0] = acc;
}

template <int V>
void run(const uint32_t *table, unsigned log2_sectors, uint32_t *out, const char *what) {
    const int blocks = 148 * 4 * 4;
    const unsigned rounds = 64;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_fetch<V><<<blocks, 256>>>(table, log2_sectors, 4, out);
    cudaEventRecord(a);
    k_fetch<V><<<blocks, 256>>>(table, log2_sectors, rounds, out);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double sectors = (double)blocks * 256 * rounds * 8;
    printf("{\"variant\": %d, \"what\": \"%s\", \"sectors\": %.0f, \"ms\": %.3f, \"useful_GBps\": %.1f, \"err\": \"%s\"}\n", V, what, sectors, ms,
           sectors * 32 / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const unsigned log2_sectors = argc > 1 ? atoi(argv[1]) : 26;     // 2^26 x 32 B = 2 GiB
    uint32_t *table, *out;
    const size_t bytes = (size_t)32 << log2_sectors;
    if (cudaMalloc(&table, bytes) != cudaSuccess) { printf("alloc failed\n"); return 1; }
    cudaMalloc(&out, 64);
    cudaMemset(table, 0x5a, bytes);
    printf("{\"table_MiB\": %zu}\n", bytes >> 20);
    run<0>(table, log2_sectors, out, "nc.L1::no_allocate.v8 (shipped)");
    if (argc > 2) return 0;                                           // size sweep: the shipped flavour only
    run<1>(table, log2_sectors, out, "nc.L1::no_allocate.L2::64B.v8");
    run<2>(table, log2_sectors, out, "nc.L1::no_allocate.L2::128B.v8");
    run<3>(table, log2_sectors, out, "nc.L1::no_allocate.L2::256B.v8");
    run<4>(table, log2_sectors, out, "cg.v8");
    run<5>(table, log2_sectors, out, "cg.L2::64B.v8");
    run<6>(table, log2_sectors, out, "cv.v8");
    run<7>(table, log2_sectors, out, "2 x nc.L1::no_allocate.v4");
    run<8>(table, log2_sectors, out, "nc.L1::no_allocate.b32 (4 bytes of the sector)");
    run<9>(table, log2_sectors, out, "nc.L1::no_allocate.L2::evict_first.L2::64B.v8");
    size_t g = 32;
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, g);
    cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
    printf("{\"limit_max_l2_fetch_granularity\": %zu}\n", g);
    run<0>(table, log2_sectors, out, "shipped, cudaLimitMaxL2FetchGranularity=32");
    run<1>(table, log2_sectors, out, "L2::64B, cudaLimitMaxL2FetchGranularity=32");
    return 0;
}

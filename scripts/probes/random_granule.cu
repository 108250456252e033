// Probe: random aligned loads of 32 / 64 / 128 bytes from a footprint far beyond L2, under the three settings of
// cudaLimitMaxL2FetchGranularity.  Answers: what does a random DRAM access cost as a function of its size, i.e. are
// 64-byte buckets / 64-byte read slots as cheap as 32-byte ones?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_granule random_granule.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
template <int BYTES>
__global__ void rnd_kernel(const uint4 *buf, uint64_t n_units, uint64_t n_ops, uint32_t *sink) {
    uint32_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_ops; i += (uint64_t) gridDim.x * blockDim.x) {
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 31;
        h *= 0xD6E8FEB86659FD93ull;
        const uint64_t s = (h >> 20) % n_units;
        const uint4 *p = buf + s * (BYTES / 16);
#pragma unroll
        for (int k = 0; k < BYTES / 16; k++) {
            uint4 a;
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p + k));
            acc += a.x ^ a.w;
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int BYTES>
float run(const uint4 *buf, size_t bytes, uint64_t n_ops, uint32_t *sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const uint64_t n_units = bytes / BYTES;
    rnd_kernel<BYTES><<<148 * 16, 256>>>(buf, n_units, n_ops / 10, sink);
    cudaEventRecord(a);
    rnd_kernel<BYTES><<<148 * 16, 256>>>(buf, n_units, n_ops, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}
int main() {
    const size_t max_bytes = 4ull << 30;
    uint4 *buf;
    uint32_t *sink;
    cudaMalloc(&buf, max_bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, max_bytes);
    const uint64_t n_ops = 100000000ull;
    for (size_t gran : {0, 32, 64, 128}) {
        if (gran) {
            cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            printf("-- cudaLimitMaxL2FetchGranularity := %zu (%s), now %zu\n", gran, cudaGetErrorString(e), got);
        } else {
            size_t got = 0;
            cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
            printf("-- default cudaLimitMaxL2FetchGranularity = %zu\n", got);
        }
        for (size_t mb : {64, 1024, 4096}) {
            const size_t bytes = mb << 20;
            const float m32 = run<32>(buf, bytes, n_ops, sink), m64 = run<64>(buf, bytes, n_ops, sink), m128 = run<128>(buf, bytes, n_ops, sink);
            printf("%5zu MB: 32 B %6.1f G/s (%6.0f GB/s) | 64 B %6.1f G/s (%6.0f GB/s) | 128 B %6.1f G/s (%6.0f GB/s)\n", mb,
                   n_ops / m32 / 1e6, n_ops * 32.0 / m32 / 1e6, n_ops / m64 / 1e6, n_ops * 64.0 / m64 / 1e6, n_ops / m128 / 1e6,
                   n_ops * 128.0 / m128 / 1e6);
        }
    }
    return 0;
}

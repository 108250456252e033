// Probe: random 32-byte sector loads over buffers of growing size (L2 / TLB reach of the chip).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__global__ void rnd_kernel(const uint4 *buf, uint64_t n_sectors, uint64_t n_ops, uint32_t *sink) {
    uint32_t acc = 0;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_ops; i += (uint64_t) gridDim.x * blockDim.x) {
        uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
        h ^= h >> 31;
        h *= 0xD6E8FEB86659FD93ull;
        const uint64_t s = (h >> 20) % n_sectors;
        const uint4 a = __ldg(buf + 2 * s), b = __ldg(buf + 2 * s + 1);
        acc += a.x ^ b.w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
int main() {
    const size_t max_bytes = 8ull << 30;
    uint4 *buf;
    uint32_t *sink;
    cudaMalloc(&buf, max_bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, max_bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const uint64_t n_ops = 100000000ull;
    for (size_t mb : {32, 64, 128, 256, 512, 768, 1024, 1536, 2048, 3072, 4096, 8192}) {
        const uint64_t n_sectors = (uint64_t) mb * 1024 * 1024 / 32;
        rnd_kernel<<<148 * 16, 256>>>(buf, n_sectors, n_ops / 10, sink);
        cudaEventRecord(a);
        rnd_kernel<<<148 * 16, 256>>>(buf, n_sectors, n_ops, sink);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        printf("%5zu MB: %7.3f ms  %6.1f G sectors/s  %7.1f GB/s\n", mb, ms, n_ops / ms / 1e6, n_ops * 32.0 / ms / 1e6);
    }
    return 0;
}

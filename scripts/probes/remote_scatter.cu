// Probe: rate of "atomicAdd with return + 16-byte store" scattered over a local vs a peer buffer.
#include <cuda_runtime.h>
#include <stdint.h>
__global__ void scatter_kernel(uint32_t *indeg, uint4 *rows, uint32_t n_targets, uint64_t n_ops, uint32_t cap) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_ops; i += (uint64_t) gridDim.x * blockDim.x) {
        uint64_t h = i * 0x9E3779B97F4A7C15ull;
        h ^= h >> 29;
        const uint32_t c = (uint32_t) ((h * 0xD6E8FEB86659FD93ull) >> 32) % n_targets;
        const uint32_t pos = atomicAdd(indeg + c, 1u);
        rows[(uint64_t) c * cap + (pos % cap)] = make_uint4((uint32_t) i, pos, c, 7u);
    }
}
extern "C" float probe_scatter(void *indeg, void *rows, uint32_t n_targets, uint64_t n_ops, uint32_t cap, int blocks, int reps) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    scatter_kernel<<<blocks, 256>>>((uint32_t *) indeg, (uint4 *) rows, n_targets, n_ops, cap);
    cudaEventRecord(a);
    for (int r = 0; r < reps; r++) scatter_kernel<<<blocks, 256>>>((uint32_t *) indeg, (uint4 *) rows, n_targets, n_ops, cap);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return cudaGetLastError() == cudaSuccess ? ms / reps : -1.f;
}

// Probe: do several lanes of ONE load instruction that touch sectors of the same 128-byte line cost one request or one per
// sector?  Groups of G lanes fetch G consecutive 32-byte sectors of a random G*32-byte aligned block with one LDG.256.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_coop random_coop.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint64_t rnd(uint64_t i) {
    uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 31;
    h *= 0xD6E8FEB86659FD93ull;
    return h >> 20;
}
template <int G, int MLP>
__global__ void k(const uint32_t *buf, uint64_t n_blocks, uint64_t n_ops, uint32_t *sink) {
    uint32_t acc = 0;
    const uint64_t tid = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x, stride = (uint64_t) gridDim.x * blockDim.x;
    const uint64_t grp = tid / G, n_grp = stride / G;
    const int sub = (int) (tid % G);
    for (uint64_t i = grp; i < n_ops; i += n_grp * MLP) {
        uint32_t v[MLP][8];
#pragma unroll
        for (int m = 0; m < MLP; m++) {
            const uint64_t b = rnd(i + m * n_grp) % n_blocks;
            const uint32_t *p = buf + (b * G + sub) * 8;
            asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[m][0]), "=r"(v[m][1]), "=r"(v[m][2]), "=r"(v[m][3]), "=r"(v[m][4]), "=r"(v[m][5]), "=r"(v[m][6]), "=r"(v[m][7]) : "l"(p));
        }
#pragma unroll
        for (int m = 0; m < MLP; m++)
#pragma unroll
            for (int q = 0; q < 8; q++) acc += v[m][q];
    }
    if (acc == 0x12345678u) *sink = acc;
}
template <int G, int MLP>
void run(const uint32_t *buf, size_t bytes, int bps, uint32_t *sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const uint64_t n_ops = 40000000ull, n_blocks = bytes / (32 * G);
    k<G, MLP><<<148 * bps, 256>>>(buf, n_blocks, n_ops / 8, sink);
    cudaEventRecord(a);
    k<G, MLP><<<148 * bps, 256>>>(buf, n_blocks, n_ops, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("  %d lanes x LDG.256 = %3d contiguous bytes per block, MLP %d, %d blocks/SM: %6.1f G blocks/s = %6.1f G sectors/s = %6.0f GB/s\n", G,
           32 * G, MLP, bps, n_ops / ms / 1e6, n_ops * G / ms / 1e6, n_ops * 32.0 * G / ms / 1e6);
}
int main() {
    const size_t max_bytes = 4ull << 30;
    uint32_t *buf, *sink;
    cudaMalloc(&buf, max_bytes);
    cudaMalloc(&sink, 4);
    cudaMemset(buf, 1, max_bytes);
    for (size_t mb : {48, 2048}) {
        printf("== footprint %zu MB\n", mb);
        const size_t bytes = mb << 20;
        run<1, 1>(buf, bytes, 8, sink); run<1, 4>(buf, bytes, 4, sink);
        run<2, 1>(buf, bytes, 8, sink); run<2, 4>(buf, bytes, 4, sink);
        run<4, 1>(buf, bytes, 8, sink); run<4, 4>(buf, bytes, 4, sink);
        run<8, 1>(buf, bytes, 8, sink); run<8, 4>(buf, bytes, 4, sink);
    }
    return 0;
}

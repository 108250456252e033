// Probe: what limits random DRAM access on B200 -- sectors or REQUESTS?  One random 32-byte sector (or 64-byte pair) per
// operation from a footprint far beyond L2, fetched with different instruction mixes.  Independent operations per thread
// (MLP) and resident warps per SM are varied.   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o random_requests ...
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

__device__ __forceinline__ uint64_t rnd(uint64_t i) {
    uint64_t h = (i + 1) * 0x9E3779B97F4A7C15ull;
    h ^= h >> 31;
    h *= 0xD6E8FEB86659FD93ull;
    return h >> 20;
}
enum Mode { LDG128x2_NA = 0, LDG128x2_CA, LDG256_NA, LDG256_CA, LDG32x8_CA, LDG32x8_NA, CPASYNC16x2_CG, CPASYNC16x2_CA, LDG256x2_NA_64B, LDG128_HALF_NA,
            CPASYNC16_HALF, LDG32_ONE, N_MODES };
const char *names[] = {"2 x LDG.128 no_allocate", "2 x LDG.128 (L1 allocate)", "1 x LDG.256 no_allocate", "1 x LDG.256 (L1 allocate)",
                       "8 x LDG.32 (L1 allocate)", "8 x LDG.32 no_allocate", "2 x cp.async.cg 16", "2 x cp.async.ca 16", "2 x LDG.256 no_allocate (64 B)",
                       "1 x LDG.128 no_allocate (half sector)", "1 x cp.async.cg 16 (half sector)", "1 x LDG.32 (4 bytes of a sector)"};

template <int MODE, int MLP>
__global__ void k(const uint32_t *buf, uint64_t n_sectors, uint64_t n_ops, uint32_t *sink) {
    __shared__ __align__(16) uint32_t sm[256 * 8 * MLP];
    uint32_t acc = 0;
    const uint64_t stride = (uint64_t) gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_ops; i += stride * MLP) {
        uint32_t v[MLP][8];
#pragma unroll
        for (int m = 0; m < MLP; m++) {
            const uint64_t s = rnd(i + m * stride) % n_sectors;
            const uint32_t *p = buf + s * 8;
#pragma unroll
            for (int q = 0; q < 8; q++) v[m][q] = 0;
            if (MODE == LDG128x2_NA || MODE == LDG128_HALF_NA) {
#pragma unroll
                for (int h = 0; h < (MODE == LDG128_HALF_NA ? 1 : 2); h++)
                    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[m][4 * h]), "=r"(v[m][4 * h + 1]), "=r"(v[m][4 * h + 2]), "=r"(v[m][4 * h + 3]) : "l"(p + 4 * h));
            } else if (MODE == LDG128x2_CA) {
#pragma unroll
                for (int h = 0; h < 2; h++)
                    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[m][4 * h]), "=r"(v[m][4 * h + 1]), "=r"(v[m][4 * h + 2]), "=r"(v[m][4 * h + 3]) : "l"(p + 4 * h));
            } else if (MODE == LDG256_NA || MODE == LDG256x2_NA_64B) {
                const uint32_t *q = MODE == LDG256x2_NA_64B ? buf + (s & ~1ull) * 8 : p;
                asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[m][0]), "=r"(v[m][1]), "=r"(v[m][2]), "=r"(v[m][3]), "=r"(v[m][4]), "=r"(v[m][5]), "=r"(v[m][6]), "=r"(v[m][7]) : "l"(q));
                if (MODE == LDG256x2_NA_64B) {
                    uint32_t w[8];
                    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "l"(q + 8));
                    v[m][0] ^= w[0] ^ w[7];
                }
            } else if (MODE == LDG256_CA) {
                asm volatile("ld.global.nc.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=r"(v[m][0]), "=r"(v[m][1]), "=r"(v[m][2]), "=r"(v[m][3]), "=r"(v[m][4]), "=r"(v[m][5]), "=r"(v[m][6]), "=r"(v[m][7]) : "l"(p));
            } else if (MODE == LDG32x8_CA) {
#pragma unroll
                for (int q = 0; q < 8; q++) v[m][q] = __ldg(p + q);
            } else if (MODE == LDG32x8_NA) {
#pragma unroll
                for (int q = 0; q < 8; q++) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v[m][q]) : "l"(p + q));
            } else if (MODE == LDG32_ONE) {
                v[m][0] = __ldg(p);
            } else {  // cp.async
                uint32_t *d = sm + (threadIdx.x * MLP + m) * 8;
                const uint32_t da = (uint32_t) __cvta_generic_to_shared(d);
                if (MODE == CPASYNC16x2_CG || MODE == CPASYNC16_HALF) {
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da), "l"(p) : "memory");
                    if (MODE == CPASYNC16x2_CG) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(da + 16), "l"(p + 4) : "memory");
                } else {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(da), "l"(p) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(da + 16), "l"(p + 4) : "memory");
                }
            }
        }
        if (MODE == CPASYNC16x2_CG || MODE == CPASYNC16x2_CA || MODE == CPASYNC16_HALF) {
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int m = 0; m < MLP; m++) acc += sm[(threadIdx.x * MLP + m) * 8] ^ sm[(threadIdx.x * MLP + m) * 8 + 3];
        } else {
#pragma unroll
            for (int m = 0; m < MLP; m++)
#pragma unroll
                for (int q = 0; q < 8; q++) acc += v[m][q];
        }
    }
    if (acc == 0x12345678u) *sink = acc;
}

template <int MODE, int MLP>
void run(const uint32_t *buf, size_t bytes, int blocks_per_sm, uint32_t *sink) {
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const uint64_t n_ops = 60000000ull, n_sectors = bytes / 32;
    k<MODE, MLP><<<148 * blocks_per_sm, 256>>>(buf, n_sectors, n_ops / 8, sink);
    cudaEventRecord(a);
    k<MODE, MLP><<<148 * blocks_per_sm, 256>>>(buf, n_sectors, n_ops, sink);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms;
    cudaEventElapsedTime(&ms, a, b);
    printf("  %-42s MLP %d, %d blocks/SM (%4d thr/SM): %6.1f G ops/s\n", names[MODE], MLP, blocks_per_sm, blocks_per_sm * 256, n_ops / ms / 1e6);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) printf("  CUDA error: %s\n", cudaGetErrorString(e));
}
template <int MODE>
void sweep(const uint32_t *buf, size_t bytes, uint32_t *sink) {
    run<MODE, 1>(buf, bytes, 8, sink);
    run<MODE, 1>(buf, bytes, 2, sink);
    run<MODE, 4>(buf, bytes, 8, sink);
    run<MODE, 4>(buf, bytes, 2, sink);
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
        sweep<LDG128x2_NA>(buf, bytes, sink);
        sweep<LDG128x2_CA>(buf, bytes, sink);
        sweep<LDG256_NA>(buf, bytes, sink);
        sweep<LDG256_CA>(buf, bytes, sink);
        sweep<LDG32x8_CA>(buf, bytes, sink);
        sweep<LDG32x8_NA>(buf, bytes, sink);
        sweep<CPASYNC16x2_CG>(buf, bytes, sink);
        sweep<CPASYNC16x2_CA>(buf, bytes, sink);
        sweep<LDG256x2_NA_64B>(buf, bytes, sink);
        sweep<LDG128_HALF_NA>(buf, bytes, sink);
        sweep<CPASYNC16_HALF>(buf, bytes, sink);
        sweep<LDG32_ONE>(buf, bytes, sink);
    }
    return 0;
}

// The two stages of the single-GPU build that are a RANDOM SCATTER by nature, done by sorting instead:
//
//   seed index   every read drops two 6-byte entries into 128-byte buckets chosen by a hash (GraphCreatorPrefSuf.cpp:90-104
//                builds `prefixHashes` / `suffixHashes` here).  One atomic + two stores per entry at random addresses are three
//                requests to the memory system (a random access costs one request whatever it carries, 39 G requests/s:
//                DESIGN.md 4.1) -- 340 M requests, 8.6 ms, and 220 bytes of DRAM traffic per 6-byte entry (ncu r2f).
//   CSR rows     phase 2 emits the edges grouped by TARGET, the result (Graph::V, Graph.h:24) is grouped by SOURCE: one
//                atomic on the row cursor + one 8-byte store per edge, 170 bytes of DRAM traffic per 12-byte edge.
//
// Sorted by bucket / by source, the same data is written in address order: every line leaves L2 once.  The sort itself is
// cub::DeviceRadixSort (library code, four passes over 12-byte records, sequential traffic: 0.46 ms per pass for 57 M records); the kernels around it are here.
// Only the order INSIDE a bucket changes (it was the order of the atomics before and never mattered: probing collects the
// matches of a window and orders them by read id), and rows are sorted by (neighbour, offset) afterwards as before.
#include <cub/cub.cuh>

#include "common.cuh"
#include "launch.h"

namespace alga {
namespace {

constexpr size_t kAlign = 256;
inline size_t up(size_t b) { return (b + kAlign - 1) & ~(kAlign - 1); }
inline int bits_for(uint64_t max_value) {  // bits needed to tell 0 .. max_value apart
    int b = 1;
    while (b < 64 && (max_value >> b)) b++;
    return b;
}
inline void bump(const LaunchCfg &cfg, int n = 1) {
    if (cfg.launches) *cfg.launches += (uint64_t) n;
}
inline int grid_1d(uint64_t n, int block, const LaunchCfg &cfg, int per_sm = 8) {
    uint64_t need = (n + block - 1) / block, cap = (uint64_t) cfg.sm_count * per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

// ---- seed index ------------------------------------------------------------------------------------------------
// key = bucket, value = read id | tag << 32; reads that do not take part get the bucket number n_buckets (sorts behind all).
// Both sides in one pass over the reads (any layout: the caller's compact one will do, so the pass does not wait for the repack):
// prefix table = window at the start of the read, suffix table = window at its end.
__global__ void seed_records_kernel(ReadsDev R, PsDev P, SeedTable Tp, SeedTable Ts, uint32_t n, uint32_t *__restrict__ keys_p,
                                    uint64_t *__restrict__ vals_p, uint32_t *__restrict__ keys_s, uint64_t *__restrict__ vals_s) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t len = P.uniform_len ? P.uniform_len : R.len[i];
        const bool any = len != 0 && (int64_t) len >= P.lmin;
        const bool on_p = any && flag_to(R, (uint32_t) i);
        const bool on_s = any && flag_from(R, (uint32_t) i) && (int64_t) len - P.min_offset >= P.lmin;
        const uint32_t *p = read_ptr(R, (uint32_t) i);
        uint32_t key = Tp.n_buckets;
        uint64_t val = i;
        if (on_p) {
            const uint64_t win = bits64(p, 0u) & P.seed_mask, h = mix64(win);
            key = bucket_index_rt(Tp, win, h, (uint32_t) P.seed_nt);
            val |= (uint64_t) tag_of(h) << 32;
        }
        keys_p[i] = key;
        vals_p[i] = val;
        key = Ts.n_buckets;
        val = i;
        if (on_s) {
            const uint64_t win = bits64(p, 2u * (len - (uint32_t) P.seed_nt)) & P.seed_mask, h = mix64(win);
            key = bucket_index_rt(Ts, win, h, (uint32_t) P.seed_nt);
            val |= (uint64_t) tag_of(h) << 32;
        }
        keys_s[i] = key;
        vals_s[i] = val;
    }
}

// one past the last record of the bucket that starts at record i of the sorted keys: gallop, then bisect (a bucket holds three
// records on average, but a repeat can fill one with thousands: no linear scans)
__device__ __forceinline__ uint32_t bucket_end(const uint32_t *__restrict__ keys, uint32_t n, uint32_t i, uint32_t b) {
    uint32_t lo = i, step = 1;
    while (lo + step < n && keys[lo + step] == b) {
        lo += step;
        step <<= 1;
    }
    uint32_t hi = lo + step < n ? lo + step : n;  // keys[lo] is in the bucket, keys[hi] (if any) is not
    while (lo + 1 < hi) {
        const uint32_t mid = lo + ((hi - lo) >> 1);
        if (keys[mid] == b) lo = mid;
        else hi = mid;
    }
    return lo + 1;
}

// The thread of the FIRST record of a bucket writes the whole 128-byte line (count, tags, ids: eight 16-byte stores, no sector is
// read back), the others have nothing to do.  The table was cleared before (the empty buckets).  Records beyond the capacity of
// their bucket are listed as (first record, how many) for chain_overflow_kernel.
__global__ void fill_buckets_kernel(SeedTable T, const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals, uint32_t n,
                                    uint2 *__restrict__ over, uint32_t over_cap, uint32_t *n_over) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t b = keys[i];
        if (b >= T.n_buckets || (i && keys[i - 1] == b)) continue;
        const uint32_t cnt = bucket_end(keys, n, (uint32_t) i, b) - (uint32_t) i;
        const uint32_t m = cnt < (uint32_t) kBucketCap ? cnt : (uint32_t) kBucketCap;
        uint32_t w[kBucketWords];
#pragma unroll
        for (int k = 0; k < kBucketWords; k++) w[k] = 0u;
        w[0] = cnt;
#pragma unroll
        for (int k = 0; k < kBucketCap; k++) {
            if ((uint32_t) k < m) {
                const uint64_t v = vals[i + k];
                w[2 + (k >> 1)] |= ((uint32_t) (v >> 32) & 0xFFFFu) << (16 * (k & 1));
                w[12 + k] = (uint32_t) v;
            }
        }
        uint4 *line = reinterpret_cast<uint4 *>(T.slots + (uint64_t) b * kBucketWords);
#pragma unroll
        for (int q = 0; q < kBucketWords / 4; q++) line[q] = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        if (cnt > (uint32_t) kBucketCap) {
            const uint32_t at = atomicAdd(n_over, 1u);
            if (at < over_cap) over[at] = make_uint2((uint32_t) i + (uint32_t) kBucketCap, cnt - (uint32_t) kBucketCap);
        }
    }
}
// after fill_buckets_kernel has finished: the listed records chain on into the buckets that follow theirs, as insert_tag_at does
// (one warp per list entry)
__global__ void chain_overflow_kernel(SeedTable T, const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals,
                                      const uint2 *__restrict__ over, uint32_t over_cap, const uint32_t *n_over) {
    const uint32_t n_list = *n_over < over_cap ? *n_over : over_cap;
    const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5, lane = threadIdx.x & 31u;
    for (uint32_t e = warp; e < n_list; e += n_warps) {
        const uint2 o = over[e];
        const uint32_t b = keys[o.x];
        for (uint32_t k = lane; k < o.y; k += 32) {
            const uint64_t v = vals[o.x + k];
            insert_tag_at(T, (uint32_t) (v >> 32) & 0xFFFFu, next_bucket(T, b), (uint32_t) v);
        }
    }
}

// ---- CSR ---------------------------------------------------------------------------------------------------------
// key = row (source - lo; n_rows for an edge whose source lies outside [lo, hi)), value = neighbour | offset << 32
__global__ void csr_records_kernel(const int32_t *__restrict__ triples, uint64_t n, uint32_t lo, uint32_t hi, int swap,
                                   uint32_t *__restrict__ keys, uint64_t *__restrict__ vals) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t) triples[3 * i + (swap ? 1 : 0)];
        keys[i] = b >= lo && b < hi ? b - lo : hi - lo;
        vals[i] = (uint64_t) (uint32_t) triples[3 * i + (swap ? 0 : 1)] | ((uint64_t) (uint32_t) triples[3 * i + 2] << 32);
    }
}
// row_off[r] = first sorted edge whose row is >= r, for r = 0 .. n_rows; the pairs split into the two result arrays
__global__ void csr_rows_kernel(const uint32_t *__restrict__ keys, const uint64_t *__restrict__ vals, uint64_t n, uint32_t n_rows,
                                uint64_t *__restrict__ row_off, int32_t *__restrict__ nbr, int32_t *__restrict__ off) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i <= n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t k = i < n ? keys[i] : n_rows;
        const int64_t kp = i ? (int64_t) keys[i - 1] : -1;
        for (int64_t r = kp + 1; r <= (int64_t) k; r++) row_off[r] = i;
        if (i < n && k < n_rows) {
            const uint64_t v = vals[i];
            nbr[i] = (int32_t) (uint32_t) v;
            off[i] = (int32_t) (uint32_t) (v >> 32);
        }
    }
}

struct IndexWs {
    uint32_t *k0, *k1, *n_over;
    uint64_t *v0, *v1;
    uint2 *over;
    uint32_t over_cap;
    void *tmp;
    size_t tmp_bytes, total;
};
IndexWs carve_index(void *ws, uint32_t n) {
    IndexWs w{};
    size_t t = 0;
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr);
    cub::DoubleBuffer<uint64_t> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, t, dk, dv, (int) n, 0, 32);
    char *p = (char *) ws;
    size_t o = 0;
    w.over_cap = n / (uint32_t) (kBucketCap + 1) + 1;  // every listed bucket holds more than kBucketCap records
    w.v0 = (uint64_t *) (p + o), o += up((size_t) n * 8);
    w.v1 = (uint64_t *) (p + o), o += up((size_t) n * 8);
    w.k0 = (uint32_t *) (p + o), o += up((size_t) n * 4);
    w.k1 = (uint32_t *) (p + o), o += up((size_t) n * 4);
    w.over = (uint2 *) (p + o), o += up((size_t) w.over_cap * 8);
    w.n_over = (uint32_t *) (p + o), o += kAlign;
    w.tmp = p + o, o += up(t + 16);
    w.tmp_bytes = t;
    w.total = o;
    return w;
}
struct CsrWs {
    uint32_t *k0, *k1;
    uint64_t *v0, *v1;
    void *tmp;
    size_t tmp_bytes, total;
};
CsrWs carve_csr(void *ws, uint64_t n) {
    CsrWs w{};
    size_t t = 0;
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr);
    cub::DoubleBuffer<uint64_t> dv(nullptr, nullptr);
    cub::DeviceRadixSort::SortPairs(nullptr, t, dk, dv, (int) n, 0, 32);
    char *p = (char *) ws;
    size_t o = 0;
    w.v0 = (uint64_t *) (p + o), o += up((size_t) n * 8);
    w.v1 = (uint64_t *) (p + o), o += up((size_t) n * 8);
    w.k0 = (uint32_t *) (p + o), o += up((size_t) n * 4);
    w.k1 = (uint32_t *) (p + o), o += up((size_t) n * 4);
    w.tmp = p + o, o += up(t + 16);
    w.tmp_bytes = t;
    w.total = o;
    return w;
}

}  // namespace

size_t sorted_index_workspace_bytes(uint32_t n_reads) { return carve_index(nullptr, n_reads ? n_reads : 1).total; }
size_t sorted_csr_workspace_bytes(uint64_t n_edges) { return carve_csr(nullptr, n_edges ? n_edges : 1).total; }

// Seed records of both tables for the reads [0, n) (R in any layout) into the two workspaces.  n < 2^31.
void launch_seed_records(const ReadsDev &R, const PsDev &P, const SeedTable &Tp, const SeedTable &Ts, uint32_t n, void *ws_prefix,
                         void *ws_suffix, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    IndexWs wp = carve_index(ws_prefix, n), wq = carve_index(ws_suffix, n);
    seed_records_kernel<<<grid_1d(n, 256, cfg), 256, 0, s>>>(R, P, Tp, Ts, n, wp.k0, wp.v0, wq.k0, wq.v0);
    bump(cfg);
}
// One table out of its records (launch_seed_records, same workspace); the table must have been cleared.
int launch_sorted_index(const SeedTable &T, uint32_t n, void *ws, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return 0;
    IndexWs w = carve_index(ws, n);
    cub::DoubleBuffer<uint32_t> dk(w.k0, w.k1);
    cub::DoubleBuffer<uint64_t> dv(w.v0, w.v1);
    cudaError_t e = cub::DeviceRadixSort::SortPairs(w.tmp, w.tmp_bytes, dk, dv, (int) n, 0, bits_for(T.n_buckets), s);
    if (e != cudaSuccess) return (int) e;
    cudaMemsetAsync(w.n_over, 0, 4, s);
    fill_buckets_kernel<<<grid_1d(n, 256, cfg), 256, 0, s>>>(T, dk.Current(), dv.Current(), n, w.over, w.over_cap, w.n_over);
    chain_overflow_kernel<<<cfg.sm_count, 256, 0, s>>>(T, dk.Current(), dv.Current(), w.over, w.over_cap, w.n_over);
    bump(cfg, 2);
    return 0;
}

// CSR rows [lo, hi) out of `n` (source, target, offset) triples (swap: (target, source, offset)); row_off has hi - lo + 1
// entries, row_off[hi - lo] = edges kept.  Rows come out in emission order: launch_sort_rows afterwards.  n < 2^31.
int launch_sorted_csr(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, void *ws, uint64_t *row_off,
                      int32_t *nbr, int32_t *off, cudaStream_t s, const LaunchCfg &cfg) {
    const uint32_t n_rows = hi - lo;
    CsrWs w = carve_csr(ws, n ? n : 1);
    cub::DoubleBuffer<uint32_t> dk(w.k0, w.k1);
    cub::DoubleBuffer<uint64_t> dv(w.v0, w.v1);
    if (n) {
        csr_records_kernel<<<grid_1d(n, 256, cfg), 256, 0, s>>>(triples, n, lo, hi, swap, w.k0, w.v0);
        cudaError_t e = cub::DeviceRadixSort::SortPairs(w.tmp, w.tmp_bytes, dk, dv, (int) n, 0, bits_for(n_rows), s);
        if (e != cudaSuccess) return (int) e;
    }
    csr_rows_kernel<<<grid_1d(n + 1, 256, cfg), 256, 0, s>>>(dk.Current(), dv.Current(), n, n_rows, row_off, nbr, off);
    bump(cfg, 2);
    return 0;
}

}  // namespace alga

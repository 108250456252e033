// First step of the graph simplifier on the GPU (reference: GraphSimplifier::simplifyGraphOld, GraphSimplifier.cpp:110-130):
// Graph::sortEdgesByIncreasingOffset (Graph.cpp:583-614) and GraphSimplifier::cutNonAndWeaklyMetricTriangles (:228-349).
// the step right after the graph build, pure two-hop work on the CSR the build leaves on the device.
//
// The reference computes, per node i, dst[b] = min over two-hop paths i -> a -> b of w(i,a) + w(a,b) (an unordered_map
// filled by a double loop, :283-294) and marks the edge (i, b) when w(i,b) <= MAX_OFFSET_PARALLEL_PATHS and
// dst[b] == w(i,b) (:297-316); all marks are taken on the unchanged graph, then Graph::removeDirectedEdge drops every
// entry i -> b (:263-273, Graph.cpp:96-119).  Here: one thread per row, rows sorted by neighbour, so dst[b] is a merge of
// short sorted lists (binary search of b in the row of every a); kept entries are compacted by a scan and the rows
// sorted by (offset, neighbour) -- the order sortEdgesByIncreasingOffset leaves.
#include "../../include/alga_gpu.h"
#include "launch.h"

namespace alga {

namespace {

inline int grid_for(uint64_t n_items, int per_block, const LaunchCfg &cfg, int max_blocks_per_sm = 16) {
    uint64_t need = (n_items + per_block - 1) / per_block;
    const uint64_t cap = (uint64_t) cfg.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

constexpr int32_t kNoPath = 0x7FFFFFFF;

// keep[e] = 0 for the entries the reference removes; kept[i] = surviving entries of row i
__global__ void triangle_marks_kernel(const uint64_t *__restrict__ row_off, const int32_t *__restrict__ nbr,
                                      const int32_t *__restrict__ off, uint32_t n, int32_t max_offset, uint8_t *__restrict__ keep,
                                      uint32_t *__restrict__ kept) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t s = row_off[i], e = row_off[i + 1];
        uint32_t n_kept = 0;
        uint64_t k = s;
        while (k < e) {
            const int32_t b = nbr[k];
            uint64_t k_end = k + 1;
            while (k_end < e && nbr[k_end] == b) k_end++;  // every entry i -> b (one, after retainOnlySmallestOffset)
            // dst[b]: shortest two-hop path i -> a -> b
            int32_t dst = kNoPath;
            for (uint64_t k2 = s; k2 < e; k2++) {
                const int32_t a = nbr[k2];
                if (a < 0 || (uint32_t) a >= n) continue;
                uint64_t lo = row_off[a], hi = row_off[a + 1];
                const uint64_t a_end = hi;
                while (lo < hi) {  // first entry of row a with neighbour >= b
                    const uint64_t mid = (lo + hi) >> 1;
                    if (nbr[mid] < b) lo = mid + 1;
                    else hi = mid;
                }
                for (; lo < a_end && nbr[lo] == b; lo++) {
                    const int32_t d = off[k2] + off[lo];
                    dst = d < dst ? d : dst;
                }
            }
            bool remove = false;
            if (dst != kNoPath)
                for (uint64_t q = k; q < k_end; q++) remove |= off[q] <= max_offset && off[q] == dst;
            for (uint64_t q = k; q < k_end; q++) keep[q] = remove ? 0 : 1;
            if (!remove) n_kept += (uint32_t) (k_end - k);
            k = k_end;
        }
        kept[i] = n_kept;
    }
}

// out_first / out_second receive (offset, neighbour): the row sort that follows orders by (first, second)
__global__ void triangle_compact_kernel(const uint64_t *__restrict__ row_off, const int32_t *__restrict__ nbr,
                                        const int32_t *__restrict__ off, const uint8_t *__restrict__ keep, uint32_t n,
                                        const uint64_t *__restrict__ new_off, int32_t *__restrict__ out_nbr,
                                        int32_t *__restrict__ out_off) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        uint64_t w = new_off[i];
        for (uint64_t k = row_off[i]; k < row_off[i + 1]; k++) {
            if (!keep[k]) continue;
            out_nbr[w] = nbr[k];
            out_off[w] = off[k];
            w++;
        }
    }
}

}  // namespace

void launch_triangle_marks(const uint64_t *row_off, const int32_t *nbr, const int32_t *off, uint32_t n, int32_t max_offset,
                           uint8_t *keep, uint32_t *kept, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    triangle_marks_kernel<<<grid_for(n, 128, cfg), 128, 0, s>>>(row_off, nbr, off, n, max_offset, keep, kept);
    if (cfg.launches) *cfg.launches += 1;
}

void launch_triangle_compact(const uint64_t *row_off, const int32_t *nbr, const int32_t *off, const uint8_t *keep, uint32_t n,
                             const uint64_t *new_off, int32_t *out_nbr, int32_t *out_off, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    triangle_compact_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(row_off, nbr, off, keep, n, new_off, out_nbr, out_off);
    if (cfg.launches) *cfg.launches += 1;
}

}  // namespace alga

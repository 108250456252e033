/* TEST INFRASTRUCTURE ONLY -- see oracle.h.
 *
 * Restatement of the first step of GraphSimplifier::simplifyGraphOld (GraphSimplifier.cpp:110-130):
 * Graph::sortEdgesByIncreasingOffset (Graph.cpp:583-614) and GraphSimplifier::cutNonAndWeaklyMetricTriangles
 * (:228-349) -- per node i the map dst[b] = min over i -> a -> b of w(i,a) + w(a,b) (:283-294), then every edge (i, b) with
 * w <= MAX_OFFSET_PARALLEL_PATHS and dst[b] == w is collected (:297-316) and Graph::removeDirectedEdge drops all entries
 * i -> b (:263-273, Graph.cpp:96-119).  The result is returned as (src, dst, offset) triples sorted by
 * (src, offset, dst); the reference's swap-and-pop leaves each row in another order, the edge set is what is compared.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static int cmp_off_dst(const void *a, const void *b) {
    const int32_t *x = (const int32_t *) a, *y = (const int32_t *) b;
    if (x[0] != y[0]) return x[0] < y[0] ? -1 : 1;
    if (x[2] != y[2]) return x[2] < y[2] ? -1 : 1;
    return x[1] < y[1] ? -1 : (x[1] > y[1] ? 1 : 0);
}

/* edges_in: n_in triples grouped by src in ascending order (rows); returns malloc'ed triples, *n_out = count */
int32_t *oracle_cut_triangles(const int32_t *edges_in, uint64_t n_in, uint32_t n_nodes, int32_t max_offset, uint64_t *n_out) {
    uint64_t *row = (uint64_t *) calloc((size_t) n_nodes + 1, sizeof(uint64_t));
    for (uint64_t e = 0; e < n_in; e++) row[edges_in[3 * e] + 1]++;
    for (uint32_t i = 0; i < n_nodes; i++) row[i + 1] += row[i];
    int32_t *dst = (int32_t *) malloc((size_t) (n_nodes ? n_nodes : 1) * sizeof(int32_t)); /* dense stand-in for the unordered_map */
    for (uint32_t i = 0; i < n_nodes; i++) dst[i] = -1;
    uint8_t *drop = (uint8_t *) calloc(n_in ? n_in : 1, 1);
    for (uint32_t i = 0; i < n_nodes; i++) {
        for (uint64_t k = row[i]; k < row[i + 1]; k++) {
            const int32_t a = edges_in[3 * k + 1];
            for (uint64_t j = row[a]; j < row[a + 1]; j++) {
                const int32_t b = edges_in[3 * j + 1], d = edges_in[3 * k + 2] + edges_in[3 * j + 2];
                if (dst[b] < 0 || d < dst[b]) dst[b] = d;
            }
        }
        for (uint64_t k = row[i]; k < row[i + 1]; k++) {
            const int32_t b = edges_in[3 * k + 1], w = edges_in[3 * k + 2];
            if (w > max_offset) continue;
            if (dst[b] >= 0 && dst[b] == w) /* removeDirectedEdge(i, b): every entry i -> b */
                for (uint64_t q = row[i]; q < row[i + 1]; q++)
                    if (edges_in[3 * q + 1] == b) drop[q] = 1;
        }
        for (uint64_t k = row[i]; k < row[i + 1]; k++) { /* dst.clear() */
            const int32_t a = edges_in[3 * k + 1];
            for (uint64_t j = row[a]; j < row[a + 1]; j++) dst[edges_in[3 * j + 1]] = -1;
        }
    }
    uint64_t m = 0;
    for (uint64_t e = 0; e < n_in; e++) m += !drop[e];
    int32_t *out = (int32_t *) malloc((size_t) (m ? m : 1) * 12);
    uint64_t w = 0;
    for (uint64_t e = 0; e < n_in; e++)
        if (!drop[e]) memcpy(out + 3 * w++, edges_in + 3 * e, 12);
    qsort(out, m, 12, cmp_off_dst);
    free(row), free(dst), free(drop);
    *n_out = m;
    return out;
}

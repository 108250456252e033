/* TEST INFRASTRUCTURE ONLY -- see oracle.h.
 *
 * Restatement of GraphCreatorPrefSuf in the reference's canonical (--threads=1) order:
 * overlap length L ascending, suffix read id ascending, prefix read id ascending.
 */
#include "oracle.h"

#include <stdlib.h>
#include <string.h>

#define M64 1000000000000000003ULL /* Params.cpp:721 MAX_HASH_CONSIDERED */
#define M32 1000000007U            /* GraphCreatorPrefSuf.h:42 MAX_ADDITIONAL_HASH */
#define SMALL_EDGES_KEPT 3         /* GraphCreatorPrefSuf.h:62 SOES */

typedef struct {
    int32_t nbr, off;
} edge_t;

typedef struct {
    edge_t *e;
    uint32_t size, cap;
} row_t;

static inline uint32_t nt_at(const oracle_reads *r, uint32_t i, uint32_t j) {
    const uint32_t *w = r->words + r->word_off[i];
    return (w[j >> 4] >> ((j & 15u) * 2u)) & 3u;
}

static void row_push(row_t *row, int32_t nbr, int32_t off) {
    if (row->size == row->cap) {
        row->cap = row->cap ? row->cap * 2 : 4;
        row->e = (edge_t *) realloc(row->e, (size_t) row->cap * sizeof(edge_t));
    }
    row->e[row->size].nbr = nbr;
    row->e[row->size].off = off;
    row->size++;
}

static int edge_cmp(const void *a, const void *b) {
    const edge_t *x = (const edge_t *) a, *y = (const edge_t *) b;
    if (x->nbr != y->nbr) return x->nbr < y->nbr ? -1 : 1;
    if (x->off != y->off) return x->off < y->off ? -1 : 1;
    return 0;
}

/* Graph::reverseGraphInPlace (Graph.cpp:926-971): edge (j -> d, off) becomes (d -> j, off). */
static row_t *transpose(row_t *g, uint32_t n) {
    row_t *t = (row_t *) calloc(n ? n : 1, sizeof(row_t));
    for (uint32_t j = 0; j < n; j++) {
        for (uint32_t k = 0; k < g[j].size; k++) row_push(&t[g[j].e[k].nbr], (int32_t) j, g[j].e[k].off);
        free(g[j].e);
    }
    free(g);
    return t;
}

/* Graph::retainOnlySmallestOffset (Graph.cpp:348-387): sort by (nbr, off), keep the first entry per nbr. */
static void retain_smallest_offset(row_t *g, uint32_t n) {
    for (uint32_t i = 0; i < n; i++) {
        row_t *row = &g[i];
        if (row->size < 2) continue;
        qsort(row->e, row->size, sizeof(edge_t), edge_cmp);
        uint32_t w = 1;
        for (uint32_t k = 1; k < row->size; k++)
            if (row->e[k].nbr != row->e[w - 1].nbr) row->e[w++] = row->e[k];
        row->size = w;
    }
}

/* a[d .. d+o) == b[0 .. o) ?  -- the block copy + shift + mismatchBounded test of
 * GraphCreatorPrefSuf.cpp:434-451 (Bitset.cpp:116-163, 879-903), stated on nucleotides. */
static int left_part_matches(const oracle_reads *r, uint32_t a, uint32_t d, uint32_t b, uint32_t o) {
    for (uint32_t k = 0; k < o; k++)
        if (nt_at(r, a, d + k) != nt_at(r, b, k)) return 0;
    return 1;
}

int32_t *oracle_prefsuf(const oracle_reads *r, int32_t min_overlap, int32_t rs_min_overlap, int32_t min_offset,
                        int32_t max_len_cap, uint64_t *n_edges) {
    const uint32_t n = r->n;
    *n_edges = 0;
    if (min_overlap < 1) return NULL;

    uint64_t *p64 = (uint64_t *) calloc(n ? n : 1, 8), *s64 = (uint64_t *) calloc(n ? n : 1, 8);
    uint32_t *p32 = (uint32_t *) calloc(n ? n : 1, 4), *s32 = (uint32_t *) calloc(n ? n : 1, 4);
    uint8_t *to = (uint8_t *) malloc(n ? n : 1), *from = (uint8_t *) malloc(n ? n : 1);
    uint8_t *mark = (uint8_t *) calloc(n ? n : 1, 1);
    int64_t maxlen = 0; /* calculateMaxReadLength, GraphCreatorPrefSuf.cpp:52-56 */
    for (uint32_t i = 0; i < n; i++) {
        to[i] = r->len_nt[i] && r->align_to[i];
        from[i] = r->len_nt[i] && r->align_from[i];
        if ((int64_t) r->len_nt[i] > maxlen) maxlen = r->len_nt[i];
    }

    /* createInitialStateJob, GraphCreatorPrefSuf.cpp:179-211: roll lengths 1 .. min_overlap-1 */
    for (uint32_t i = 0; i < n; i++) {
        if (!(to[i] || from[i])) continue;
        const int64_t len = r->len_nt[i];
        uint64_t f64 = 1;
        uint32_t f32 = 1;
        for (int64_t l = 1; l <= (int64_t) min_overlap - 1; l++) {
            if (to[i]) {
                if (l > len) {
                    to[i] = 0;
                } else {
                    uint32_t c = nt_at(r, i, (uint32_t) (l - 1));
                    p64[i] = (p64[i] + c * f64) % M64;
                    p32[i] = (uint32_t) (((uint64_t) p32[i] + (uint64_t) c * f32) % M32);
                }
            }
            if (from[i]) {
                if (l > len - min_offset) {
                    from[i] = 0;
                } else {
                    uint32_t c = nt_at(r, i, (uint32_t) (len - l));
                    s64[i] = (s64[i] * 4 + c) % M64;
                    s32[i] = (uint32_t) (((uint64_t) s32[i] * 4 + c) % M32);
                }
            }
            f64 = (f64 * 4) % M64;
            f32 = (uint32_t) (((uint64_t) f32 * 4) % M32);
        }
    }
    uint64_t f64 = 1;
    uint32_t f32 = 1;
    for (int32_t l = 0; l < min_overlap - 1; l++) {
        f64 = (f64 * 4) % M64;
        f32 = (uint32_t) (((uint64_t) f32 * 4) % M32);
    }

    /* bucket table keyed by the low bits of the prefix hash (putKmersIntoBucketsJob, :323-332).  The
     * bucket count only partitions the scan; ids inside a bucket are in ascending order. */
    uint32_t nb = 4;
    while (nb < n / 2) nb <<= 1;
    uint32_t *head = (uint32_t *) malloc(((size_t) nb + 1) * 4);
    uint32_t *ids = (uint32_t *) malloc((n ? n : 1) * 4);

    row_t *g = (row_t *) calloc(n ? n : 1, sizeof(row_t));
    int32_t *scratch = NULL;
    uint32_t scratch_cap = 0;

    int64_t maxL = maxlen < max_len_cap ? maxlen : max_len_cap; /* :92 */
    int64_t cur = (int64_t) min_overlap - 1;
    while (cur <= maxL) { /* :94 */
        cur++;
        const int64_t L = cur;

        /* updatePrexihHashJob, :347-354 */
        for (uint32_t i = 0; i < n; i++) {
            if (!to[i]) continue;
            if (L > (int64_t) r->len_nt[i]) {
                to[i] = 0;
                continue;
            }
            uint32_t c = nt_at(r, i, (uint32_t) (L - 1));
            p64[i] = (p64[i] + c * f64) % M64;
            p32[i] = (uint32_t) (((uint64_t) p32[i] + (uint64_t) c * f32) % M32);
        }
        memset(head, 0, ((size_t) nb + 1) * 4);
        for (uint32_t i = 0; i < n; i++)
            if (to[i]) head[(p64[i] & (nb - 1)) + 1]++;
        for (uint32_t k = 0; k < nb; k++) head[k + 1] += head[k];
        for (uint32_t i = 0; i < n; i++)
            if (to[i]) ids[head[p64[i] & (nb - 1)]++] = i;
        for (uint32_t k = nb; k > 0; k--) head[k] = head[k - 1];
        head[0] = 0;

        if (L == rs_min_overlap) { /* :288-296 */
            g = transpose(g, n);
            retain_smallest_offset(g, n);
        }

        /* nextPrefSufIterationJobAddEdges, :356-488 */
        for (uint32_t b = 0; b < n; b++) {
            if (!from[b]) continue;
            const int64_t lenb = r->len_nt[b];
            if (L > lenb - min_offset) {
                from[b] = 0;
                continue;
            }
            uint32_t cnt = nt_at(r, b, (uint32_t) (lenb - L));
            s64[b] = (s64[b] * 4 + cnt) % M64;
            s32[b] = (uint32_t) (((uint64_t) s32[b] * 4 + cnt) % M32);
            const int32_t o = (int32_t) (lenb - L);
            const uint32_t bk = (uint32_t) (s64[b] & (nb - 1));
            for (uint32_t q = head[bk]; q < head[bk + 1]; q++) {
                const uint32_t c = ids[q];
                if (c == b || p64[c] != s64[b] || p32[c] != s32[b]) continue;
                if (L < rs_min_overlap) { /* phase 1, :397-402 */
                    row_t *row = &g[b];
                    if (row->size == SMALL_EDGES_KEPT) {
                        memmove(row->e, row->e + 1, (row->size - 1) * sizeof(edge_t));
                        row->size--;
                    }
                    row_push(row, (int32_t) c, o);
                } else { /* phase 2, :403-483 */
                    row_t *row = &g[c];
                    uint32_t nrm = 0;
                    if (scratch_cap < row->size + 1) {
                        scratch_cap = 2 * (row->size + 1);
                        scratch = (int32_t *) realloc(scratch, (size_t) scratch_cap * 4);
                    }
                    if (o > 0) {
                        for (uint32_t k = 0; k < row->size; k++) {
                            const int32_t a = row->e[k].nbr;
                            const int32_t d = row->e[k].off - o;
                            if (d < 0 || (uint32_t) a == b) continue;
                            if (lenb + d - (int64_t) r->len_nt[a] < 0) continue; /* getRightOffset, Read.h:88 */
                            if (left_part_matches(r, (uint32_t) a, (uint32_t) d, b, (uint32_t) o)) {
                                if (!mark[a]) {
                                    mark[a] = 1;
                                    scratch[nrm++] = a;
                                }
                            }
                        }
                    }
                    mark[b] = 1;
                    uint32_t w = 0;
                    for (uint32_t k = 0; k < row->size; k++)
                        if (!mark[row->e[k].nbr]) row->e[w++] = row->e[k];
                    row->size = w;
                    for (uint32_t k = 0; k < nrm; k++) mark[scratch[k]] = 0;
                    mark[b] = 0;
                    row_push(row, (int32_t) b, o);
                }
            }
        }
        f64 = (f64 * 4) % M64;
        f32 = (uint32_t) (((uint64_t) f32 * 4) % M32);
    }

    g = transpose(g, n);            /* :107 */
    retain_smallest_offset(g, n);   /* main.cpp:291 */

    uint64_t E = 0;
    for (uint32_t i = 0; i < n; i++) E += g[i].size;
    int32_t *out = (int32_t *) malloc((E ? E : 1) * 12);
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; i++) {
        for (uint32_t k = 0; k < g[i].size; k++) {
            out[3 * w] = (int32_t) i;
            out[3 * w + 1] = g[i].e[k].nbr;
            out[3 * w + 2] = g[i].e[k].off;
            w++;
        }
        free(g[i].e);
    }
    free(g);
    free(scratch);
    free(head);
    free(ids);
    free(p64);
    free(s64);
    free(p32);
    free(s32);
    free(to);
    free(from);
    free(mark);
    *n_edges = E;
    return out;
}

void oracle_fingerprints(const oracle_reads *r, int32_t L, uint64_t *pre64, uint32_t *pre32, uint64_t *suf64,
                         uint32_t *suf32) {
    for (uint32_t i = 0; i < r->n; i++) {
        const uint32_t len = r->len_nt[i];
        if (L < 1 || (uint32_t) L > len) continue;
        uint64_t h64 = 0, f64 = 1, g64 = 0;
        uint32_t h32 = 0, f32 = 1, g32 = 0;
        for (int32_t l = 1; l <= L; l++) {
            uint32_t c = nt_at(r, i, (uint32_t) (l - 1));     /* updatePrefixHash, :213-223 */
            h64 = (h64 + c * f64) % M64;
            h32 = (uint32_t) (((uint64_t) h32 + (uint64_t) c * f32) % M32);
            f64 = (f64 * 4) % M64;
            f32 = (uint32_t) (((uint64_t) f32 * 4) % M32);
            uint32_t e = nt_at(r, i, len - (uint32_t) l);     /* updateSuffixHash, :225-236 */
            g64 = (g64 * 4 + e) % M64;
            g32 = (uint32_t) (((uint64_t) g32 * 4 + e) % M32);
        }
        pre64[i] = h64;
        pre32[i] = h32;
        suf64[i] = g64;
        suf32[i] = g32;
    }
}

void oracle_free(void *p) { free(p); }

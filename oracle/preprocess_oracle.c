/* TEST INFRASTRUCTURE ONLY -- see oracle.h.
 *
 * Restatement of ReadPreprocess::getPrefixReads (src/IO/ReadPreprocess.cpp:13-77): sort all reads in the order of
 * getSortedReads (:79-152: bit strings compared LSB first -- the first differing bit decides, a 0 sorts first --, then
 * the shorter read, then the smaller id), then mark read i when its successor starts with it
 * (Bitset::mismatch, Bitset.cpp:858-877): with type 1 (PREF_READS_ONLY_DUPLICATES) only if both have the same length,
 * with type 2 (PREF_READS_ALL_PREFIX_READS, the default) always, and then also its reverse complement (id ^ 1,
 * Read.cpp:229-236) if the successor is longer.
 */
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

static const oracle_reads *g_r;

static inline uint32_t nblocks(uint32_t i) { return (uint32_t) (g_r->word_off[i + 1] - g_r->word_off[i]); }

static int cmp_reads(const void *pa, const void *pb) {
    const uint32_t a = *(const uint32_t *) pa, b = *(const uint32_t *) pb;
    const uint32_t *wa = g_r->words + g_r->word_off[a], *wb = g_r->words + g_r->word_off[b];
    const uint32_t ma = nblocks(a), mb = nblocks(b), m = ma < mb ? ma : mb;
    for (uint32_t p = 0; p < m; p++) {
        if (wa[p] != wb[p]) {
            const int ind = __builtin_ctz(wa[p] ^ wb[p]);
            return ((wa[p] >> ind) & 1u) ? 1 : -1; /* a's bit is 1 => b's is 0 => b first */
        }
    }
    if (g_r->len_nt[a] != g_r->len_nt[b]) return g_r->len_nt[a] < g_r->len_nt[b] ? -1 : 1;
    return a < b ? -1 : (a > b ? 1 : 0);
}

/* Bitset::mismatch >> 1 */
static uint32_t lcp_nt(uint32_t a, uint32_t b) {
    const uint32_t *wa = g_r->words + g_r->word_off[a], *wb = g_r->words + g_r->word_off[b];
    const uint32_t ma = nblocks(a), mb = nblocks(b), m = ma < mb ? ma : mb;
    uint64_t ind = 1000000000ull;
    for (uint32_t p = 0; p < m; p++) {
        if (wa[p] != wb[p]) {
            ind = (uint64_t) p * 32u + (uint64_t) __builtin_ctz(wa[p] ^ wb[p]);
            break;
        }
    }
    const uint64_t sa = 2ull * g_r->len_nt[a], sb = 2ull * g_r->len_nt[b], ms = sa < sb ? sa : sb;
    return (uint32_t) ((ind < ms ? ind : ms) >> 1);
}

void oracle_prefix_reads(const oracle_reads *r, int32_t remove_type, uint8_t *mask) {
    memset(mask, 0, r->n);
    uint32_t *ids = (uint32_t *) malloc((size_t) (r->n ? r->n : 1) * sizeof(uint32_t));
    uint32_t n = 0;
    for (uint32_t i = 0; i < r->n; i++)
        if (r->len_nt[i]) ids[n++] = i;
    g_r = r;
    qsort(ids, n, sizeof(uint32_t), cmp_reads);
    for (uint32_t i = 0; i + 1 < n; i++) {
        const uint32_t a = ids[i], b = ids[i + 1];
        const uint32_t l = lcp_nt(a, b);
        if (remove_type == 1) {
            if (l == r->len_nt[a] && r->len_nt[a] == r->len_nt[b]) mask[a] = 1;
        } else if (remove_type == 2 && l == r->len_nt[a]) {
            mask[a] = 1;
            if (r->len_nt[a] < r->len_nt[b]) mask[a ^ 1u] = 1;
        }
    }
    free(ids);
}

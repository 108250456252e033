/* TEST INFRASTRUCTURE ONLY -- see oracle.h.
 *
 * Restatement of the candidate verification used by the error-rate supplement:
 * AlignmentControllerHybrid::canAlign (AlignmentControllerHybrid.cpp:46-83) with the reference
 * defaults USE_LCS_LOW_ERROR_FILTER=1, USE_ACLER_INSTEAD_OF_ACLCS=1 (Params.cpp:702-703), i.e. the
 * verdict is AlignmentControllerLowErrorRate::canAlign (AlignmentControllerLowErrorRate.cpp:15-49); with
 * USE_ACLER_INSTEAD_OF_ACLCS=0 (lcs_rate_pct > 0) the pairs it rejects go on to AlignmentControllerLCS::canAlign
 * (AlignmentControllerLCS.cpp:30-59, banded LCS :61-150).
 */
#include "oracle.h"

static inline uint32_t bit_at(const oracle_reads *r, uint32_t i, uint64_t bit) {
    /* bits past the end of a read are zero (Bitset tail / operator<<= zero fill, Bitset.cpp:116-163) */
    if (bit >= 2ull * r->len_nt[i]) return 0;
    return (r->words[r->word_off[i] + (bit >> 5)] >> (bit & 31u)) & 1u;
}

static inline uint32_t nt_at(const oracle_reads *r, uint32_t i, int64_t pos) {
    return bit_at(r, i, 2ull * (uint64_t) pos) | (bit_at(r, i, 2ull * (uint64_t) pos + 1) << 1);
}

/* AlignmentControllerLCS::calculateLCS :61-150: longest common subsequence of r1 and r2 inside the band
 * |q - (p - offset)| <= E.  The reference keeps the table in hash maps, so a cell it never wrote reads as 0: that is
 * every cell outside the band or outside r2.  Two rows of 2E+3 cells are all it takes. */
static int64_t banded_lcs(const oracle_reads *r, int32_t a, int32_t b, int32_t off, int32_t E) {
    const int64_t la = r->len_nt[a], lb = r->len_nt[b];
    int64_t prev[64], cur[64];                                      /* index d + E + 1, d = q - (p - off) in [-E-1, E+1] */
    if (E < 0 || 2 * E + 3 > 64) return 0;
    for (int k = 0; k < 2 * E + 3; k++) prev[k] = cur[k] = 0;
    const int64_t p_beg = off - E > 0 ? off - E : 0;
    const int64_t p_last = la - 1 < lb - 1 + off ? la - 1 : lb - 1 + off;   /* :147 */
    for (int64_t pp = p_beg; pp <= p_last; pp++) {
        for (int k = 0; k < 2 * E + 3; k++) cur[k] = 0;
        for (int d = -E; d <= E; d++) {
            const int64_t q = pp - off + d;
            if (q < 0 || q > lb - 1) continue;
            int64_t v = 0;
            if (nt_at(r, (uint32_t) a, pp) == nt_at(r, (uint32_t) b, q)) {
                v = (pp > 0 && q > 0) ? prev[d + E + 1] + 1 : 1;   /* lcs[p-1][q-1]: the same diagonal of the row before */
            } else {
                if (pp > 0 && v < prev[d + 1 + E + 1]) v = prev[d + 1 + E + 1];   /* lcs[p-1][q] */
                if (q > 0 && v < cur[d - 1 + E + 1]) v = cur[d - 1 + E + 1];       /* lcs[p][q-1] */
            }
            cur[d + E + 1] = v;
        }
        for (int k = 0; k < 2 * E + 3; k++) prev[k] = cur[k];
    }
    if (p_last < p_beg) return 0;
    const int64_t q_last = lb - 1 < p_last - off + E ? lb - 1 : p_last - off + E;  /* :148 */
    const int64_t d = q_last - (p_last - off);
    return (d < -E || d > E) ? 0 : prev[d + E + 1];
}

static int can_align(const oracle_reads *r, int32_t a, int32_t b, int32_t off, const oracle_verify_params *p) {
    const int64_t la = r->len_nt[a], lb = r->len_nt[b];
    if (la == 0 || lb == 0) return 0;
    if (100ll * off > (int64_t) p->max_offset_pct * la) return 0;          /* Hybrid :50-52 */
    if (off < p->min_offset) return 0;                                      /* :54 */
    const int64_t ov = (la < lb + off ? la : lb + off) - off;               /* Read.h:81-83 */
    if (ov < p->min_overlap_area) return 0;                                 /* :56-57 */
    if (lb + off - la < 0) return 0;                                        /* :59 getRightOffset */
    if (off < 0) return 0;
    /* ACLER :33-47 -- x = (a >> 2*off) XOR b over the first 2*ov bits; mismatches counted in bits */
    int64_t diff = 0, head = 0, tail = 0;
    const int64_t nb = 2 * ov;
    for (int64_t k = 0; k < nb; k++) {
        uint32_t x = bit_at(r, (uint32_t) a, (uint64_t) (k + 2ll * off)) ^ bit_at(r, (uint32_t) b, (uint64_t) k);
        diff += x;
        if (k <= 2ll * p->same_ends) head += x;                 /* count(0, SAME<<1): bits 0..2*SAME inclusive */
        if (k >= 2 * (ov - p->same_ends)) tail += x;            /* count((ov-SAME)<<1, 2*ov-1) */
    }
    const int64_t sim = (nb - diff) >> 1;
    if (head == 0 && tail == 0 && 100 * sim >= (int64_t) p->threshold_pct * ov) return 1;
    if (p->lcs_rate_pct <= 0) return 0;                                     /* Hybrid :72: USE_ACLER_INSTEAD_OF_ACLCS */
    /* AlignmentControllerLCS::canAlign :30-59 (its two filters repeat the ones above) */
    return 100 * banded_lcs(r, a, b, off, p->lcs_band) > (int64_t) p->lcs_rate_pct * ov;
}

void oracle_verify_pairs(const oracle_reads *r, const int32_t *pairs, uint64_t n_pairs,
                         const oracle_verify_params *p, uint8_t *verdict) {
    for (uint64_t i = 0; i < n_pairs; i++)
        verdict[i] = (uint8_t) can_align(r, pairs[3 * i], pairs[3 * i + 1], pairs[3 * i + 2], p);
}

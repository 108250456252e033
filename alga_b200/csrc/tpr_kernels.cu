// Thread-per-read overlap search kernels (sm_100a): the fast path of GraphCreatorPrefSuf phase 1 and phase 2.
//
// One THREAD owns one read and walks its overlap lengths one per iteration; the 32 reads of a warp advance in
// lock step.  Per iteration a thread extracts its K-nucleotide seed window from shared memory, hashes it and
// loads one 32-byte bucket (the bucket of the next iteration is already in flight).  Tag hits are rare and
// expensive (a 2-bit compare of the whole overlap), so the compares are handed to the warp: the lanes are cut
// into groups of 8/10/16/32, each group verifies one requesting thread's candidate, one 32-bit word per lane,
// straight from the requester's staged read in shared memory.  All bookkeeping (the 3 best phase-1 edges, the
// surviving phase-2 arrivals) lives in the owning thread's registers, so the per-read overhead is thread
// instructions, not warp instructions -- the warp-per-read version of these kernels was issue-bound at 560-670
// warp instructions per read (profiles/ncu_full_r01c_summary.txt).
//
// Reads the fast path cannot take (longer than 512 nt, a window with more than two tag matches, offsets above
// 32, more than 4 surviving arrivals, a source id that occurs twice for one target, ...) are queued for the
// generic kernels of prefsuf_kernels.cu, which replay GraphCreatorPrefSuf.cpp:356-488 literally.
#include "launch.h"

namespace alga {

namespace {

constexpr int kTpr = 256;    // threads = reads per tile
constexpr int kIdCap = 32;   // candidate + in-neighbour ids remembered per target (duplicate detection)
constexpr int kSurv = 4;     // surviving arrivals kept in registers per target

inline int tile_grid(uint64_t n_items, const LaunchCfg &cfg, int blocks_per_sm) {
    uint64_t need = (n_items + kTpr - 1) / kTpr;
    uint64_t cap = (uint64_t) cfg.sm_count * blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}
inline void bump(const LaunchCfg &cfg) {
    if (cfg.launches) (*cfg.launches)++;
}

__device__ __forceinline__ int warp_max(int v) {
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(kFull, v, d));
    return v;
}

// ids of the entries of one bucket whose tag matches (first two) and how many there are; true if the bucket is full
__device__ __forceinline__ bool eval_bucket(const SeedTable &t, const uint32_t (&e)[8], uint32_t tag, uint32_t &c0,
                                            uint32_t &c1, int &n) {
    uint32_t hm = 0;
#pragma unroll
    for (int s = 0; s < kSlotsPerBucket; s++) hm |= ((e[s] ^ tag) <= t.id_mask ? 1u : 0u) << s;
    if (hm) {
#pragma unroll
        for (int s = 0; s < kSlotsPerBucket; s++) {
            if (hm & (1u << s)) {
                if (n == 0) c0 = e[s] & t.id_mask;
                else if (n == 1) c1 = e[s] & t.id_mask;
                n++;
            }
        }
    }
    return e[kSlotsPerBucket - 1] != kEmptySlot;  // buckets fill front to back: a full one chains on
}

// Finish the probe whose first bucket is in e[]: chained buckets are walked on demand (p ~ 1e-3 per probe).
__device__ __forceinline__ void finish_probe(const SeedTable &t, uint32_t (&e)[8], uint32_t tag, uint32_t bk, uint32_t &c0,
                                             uint32_t &c1, int &n) {
    n = 0;
    c0 = c1 = kNone;
    bool full = eval_bucket(t, e, tag, c0, c1, n);
    while (full) {
        bk = (bk + 1 == t.n_buckets) ? 0u : bk + 1;
        load_bucket(t.slots + (uint64_t) bk * kSlotsPerBucket, e);
        full = eval_bucket(t, e, tag, c0, c1, n);
    }
}

// ------------------------------------------------------------------------------------------------------------
// Warp-cooperative verification of phase-1 candidates: requester r asks "prefix(cand, L) == suffix(own_r, L)?".
// wown = staged reads of the warp (read of lane r at wown + r * wp).
template <bool UNIFORM>
__device__ __forceinline__ bool coop_verify_suffix(const ReadsDev &R, const uint32_t *wown, int wp, const GroupGeom &q,
                                                   bool want, uint32_t cand, uint32_t self, uint32_t o, int32_t L,
                                                   int lane) {
    unsigned req = __ballot_sync(kFull, want);
    bool result = false;
    const unsigned lt = (1u << lane) - 1u;
    while (req) {
        unsigned chunk = 0;
        int src = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < q.V && req) {
                const int l = __ffs(req) - 1;
                if (i == q.g) src = l;
                chunk |= 1u << l;
                req &= req - 1;
            }
        }
        const int ng = __popc(chunk);
        const uint32_t c_s = __shfl_sync(kFull, cand, src), b_s = __shfl_sync(kFull, self, src);
        const uint32_t o_s = __shfl_sync(kFull, o, src);
        const int32_t L_s = __shfl_sync(kFull, L, src);
        const bool act = q.g < ng && q.g < q.V;
        bool bad = false;
        if (act) {
            const uint32_t nbits = 2u * (uint32_t) L_s, nw = (nbits + 31u) >> 5;
            if (q.k == 0 && (c_s == b_s || (!UNIFORM && (int64_t) R.len[c_s] < L_s))) bad = true;
            if ((uint32_t) q.k < nw) {
                const uint32_t *ow = wown + src * wp + ((2u * o_s) >> 5) + q.k;
                uint32_t x = __funnelshift_r(ow[0], ow[1], (2u * o_s) & 31u) ^ __ldg(read_ptr(R, c_s) + q.k);
                if ((uint32_t) q.k == nw - 1 && (nbits & 31u)) x &= (1u << (nbits & 31u)) - 1u;
                bad |= x != 0;
            }
        }
        const unsigned okbits = group_ok(q, act, bad);
        if ((chunk >> lane) & 1u) result = (okbits >> __popc(chunk & lt)) & 1u;
    }
    return result;
}

// Phase-2 flavour: requester r asks "suffix(cand, L) == prefix(own_r, L)?", o = len(cand) - L.
__device__ __forceinline__ bool coop_verify_prefix(const ReadsDev &R, const uint32_t *wown, int wp, const GroupGeom &q,
                                                   bool want, uint32_t cand, uint32_t o, int32_t L, int lane) {
    unsigned req = __ballot_sync(kFull, want);
    bool result = false;
    const unsigned lt = (1u << lane) - 1u;
    while (req) {
        unsigned chunk = 0;
        int src = 0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (i < q.V && req) {
                const int l = __ffs(req) - 1;
                if (i == q.g) src = l;
                chunk |= 1u << l;
                req &= req - 1;
            }
        }
        const int ng = __popc(chunk);
        const uint32_t b_s = __shfl_sync(kFull, cand, src), o_s = __shfl_sync(kFull, o, src);
        const int32_t L_s = __shfl_sync(kFull, L, src);
        const bool act = q.g < ng && q.g < q.V;
        bool bad = false;
        if (act) {
            const uint32_t nbits = 2u * (uint32_t) L_s, nw = (nbits + 31u) >> 5;
            if ((uint32_t) q.k < nw) {
                const uint32_t *qb = read_ptr(R, b_s) + ((2u * o_s) >> 5) + q.k;
                uint32_t x = __funnelshift_r(__ldg(qb), __ldg(qb + 1), (2u * o_s) & 31u) ^ wown[src * wp + q.k];
                if ((uint32_t) q.k == nw - 1 && (nbits & 31u)) x &= (1u << (nbits & 31u)) - 1u;
                bad = x != 0;
            }
        }
        const unsigned okbits = group_ok(q, act, bad);
        if ((chunk >> lane) & 1u) result = (okbits >> __popc(chunk & lt)) & 1u;
    }
    return result;
}

// ------------------------------------------------------------------------------------------------------------
// Phase 1 (GraphCreatorPrefSuf.cpp:397-402 in closed form): source read b walks L from min(rs-1, len) downwards
// and keeps the first 3 confirmed (L, c) -- within one L the larger c first -- = "the last 3 pushes".
template <bool UNIFORM>
__global__ void __launch_bounds__(kTpr, 4)
phase1_tpr_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, int wp, int nw_max, int2 *__restrict__ fwd,
                  uint64_t *__restrict__ fwd_t, uint32_t *__restrict__ fwd_pos, uint32_t *__restrict__ indeg,
                  uint32_t *__restrict__ hard_queue, uint32_t *n_hard, int force_hard) {
    extern __shared__ uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t *own = smem + tid * wp;
    const uint32_t *wown = smem + (tid & ~31) * wp;
    const GroupGeom q = group_geom(nw_max, lane);
    const uint64_t n_tiles = ((uint64_t) (hi - lo) + kTpr - 1) / kTpr;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t bb = (uint64_t) lo + tile * kTpr + tid;
        const bool inr = bb < hi;
        const uint32_t b = inr ? (uint32_t) bb : lo;
        const uint32_t lenb = inr ? (UNIFORM ? P.uniform_len : R.len[b]) : 0u;
        int64_t l_hi64 = (int64_t) lenb - P.min_offset;
        if (l_hi64 > P.rs - 1) l_hi64 = P.rs - 1;
        if (l_hi64 > P.max_l) l_hi64 = P.max_l;
        const int32_t l_hi = (int32_t) l_hi64;
        bool active = inr && lenb != 0 && flag_from(R, b) && l_hi >= P.lmin;
        bool hard = false;
        if (active && (force_hard || lenb > (uint32_t) kOwnWords * 16u)) {
            hard = true;
            active = false;
        }
        __syncwarp();
        {
            const uint32_t nw = active ? (lenb + 15u) >> 4 : 0u;
            const uint32_t *p = read_ptr(R, b);
            for (int w = 0; w < wp; w++) own[w] = (uint32_t) w < nw ? __ldg(p + w) : 0u;
        }
        __syncwarp();
        int found = 0;
        uint32_t sc0 = kNone, sc1 = kNone, sc2 = kNone, so0 = 0, so1 = 0, so2 = 0;
        const int n_iter = warp_max(active ? l_hi - P.lmin + 1 : 0);
        uint32_t e[8];
        uint32_t tag = 0, bk = 0;
        if (active) {
            const uint64_t h = mix64(sbits64(own, 2u * (lenb - (uint32_t) l_hi)) & P.seed_mask);
            tag = tag_of(T, h);
            bk = bucket_of(h, T.n_buckets);
            load_bucket(T.slots + (uint64_t) bk * kSlotsPerBucket, e);
        }
        for (int it = 0; it < n_iter; it++) {
            const int32_t L = l_hi - it;
            const bool live = active && !hard && found < kSmallEdgesKept && L >= P.lmin;
            if (!__any_sync(kFull, live)) break;
            uint32_t c0 = kNone, c1 = kNone;
            int n = 0;
            if (live) finish_probe(T, e, tag, bk, c0, c1, n);
            if (live && L - 1 >= P.lmin) {  // bucket of the next length goes in flight before the compares
                const uint64_t h = mix64(sbits64(own, 2u * (lenb - (uint32_t) (L - 1))) & P.seed_mask);
                tag = tag_of(T, h);
                bk = bucket_of(h, T.n_buckets);
                load_bucket(T.slots + (uint64_t) bk * kSlotsPerBucket, e);
            }
            if (n > 2) hard = true;
            if (n == 2 && c1 > c0) {  // within one length the larger target id is the later push
                const uint32_t x = c0;
                c0 = c1;
                c1 = x;
            }
            const uint32_t o = lenb - (uint32_t) L;
            const bool ok0 = coop_verify_suffix<UNIFORM>(R, wown, wp, q, live && !hard && n > 0, c0, b, o, L, lane);
            bool ok1 = false;
            if (__any_sync(kFull, live && !hard && n > 1))
                ok1 = coop_verify_suffix<UNIFORM>(R, wown, wp, q, live && !hard && n > 1, c1, b, o, L, lane);
            if (ok0) {
                if (found == 0) sc0 = c0, so0 = o;
                else if (found == 1) sc1 = c0, so1 = o;
                else if (found == 2) sc2 = c0, so2 = o;
                found++;
            }
            if (ok1) {
                if (found == 0) sc0 = c1, so0 = o;
                else if (found == 1) sc1 = c1, so1 = o;
                else if (found == 2) sc2 = c1, so2 = o;
                found++;
            }
        }
        if (hard) {
            hard_queue[atomicAdd(n_hard, 1u)] = b;
        } else if (inr) {
            const uint64_t s0 = (uint64_t) (b - lo) * kSmallEdgesKept;
#pragma unroll
            for (int k = 0; k < kSmallEdgesKept; k++) {
                const uint32_t c = k == 0 ? sc0 : (k == 1 ? sc1 : sc2), o = k == 0 ? so0 : (k == 1 ? so1 : so2);
                if (k < found) {
                    fwd[s0 + k] = make_int2((int32_t) c, (int32_t) o);
                    fwd_t[s0 + k] = overhang_tail_own(own, o);
                    if (indeg) fwd_pos[s0 + k] = atomicAdd(indeg + c, 1u);
                } else {
                    fwd[s0 + k] = make_int2(-1, 0);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// Phase 2 (GraphCreatorPrefSuf.cpp:403-483 in closed form): target read c walks L from len downwards, i.e. from its
// LAST arrival to its first.  An arrival (b, o) stays in c's list unless a later arrival j with o_j > 0 has an
// overhang that is a suffix of b's overhang (a[oa-oj .. oa) == b_j[0 .. oj), right offset >= 0); the relation is
// transitive, so it is enough to test against the arrivals that survived so far, and a candidate that fails this
// test is dropped without ever comparing its overlap.  Survivors are confirmed by the warp.  Entries of the
// transposed phase-1 graph (rows) are tested against the survivors at the end.
template <bool UNIFORM>
__global__ void __launch_bounds__(kTpr, 3)
phase2_tpr_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, int wp, int nw_max,
                  const uint32_t *__restrict__ rev_off, const RevEntry *__restrict__ rev, Phase2Out out, int force_hard) {
    extern __shared__ uint32_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    uint32_t *own = smem + tid * wp;
    const uint32_t *wown = smem + (tid & ~31) * wp;
    uint32_t *ids = smem + kTpr * wp + tid;  // ids[k * kTpr]: k-th remembered id of this thread
    const GroupGeom q = group_geom(nw_max, lane);
    const int32_t l_lo = P.rs > P.lmin ? P.rs : P.lmin;
    const uint64_t n_tiles = ((uint64_t) (hi - lo) + kTpr - 1) / kTpr;
    for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const uint64_t cc = (uint64_t) lo + tile * kTpr + tid;
        const bool inr = cc < hi;
        const uint32_t c = inr ? (uint32_t) cc : lo;
        uint32_t r0 = 0, deg = 0;
        if (inr) {
            r0 = rev_off[c - lo];
            deg = rev_off[c - lo + 1] - r0;
        }
        const uint32_t lenc = inr ? (UNIFORM ? P.uniform_len : R.len[c]) : 0u;
        int32_t l_hi = (int32_t) (lenc < (uint32_t) P.max_l ? lenc : (uint32_t) P.max_l);
        const bool active = inr && lenc != 0 && flag_to(R, c) && l_hi >= l_lo;
        const bool part = active || deg > 0;  // has something to emit
        bool hard = part && (force_hard || deg > (uint32_t) kIdCap);
        __syncwarp();
        {
            uint32_t nw = 0;
            if (active && !hard) {
                const uint32_t need = (uint32_t) ((2 * l_hi + 31) >> 5), have = (lenc + 15u) >> 4;
                nw = need < have ? need : have;
            }
            const uint32_t *p = read_ptr(R, c);
            for (int w = 0; w < wp; w++) own[w] = (uint32_t) w < nw ? __ldg(p + w) : 0u;
        }
        __syncwarp();
        int n_ids = 0, ns = 0;
        uint32_t s_id[kSurv], s_o[kSurv], s_len[kSurv];
        uint64_t s_t[kSurv];
#pragma unroll
        for (int s = 0; s < kSurv; s++) s_id[s] = kNone, s_o[s] = 0, s_len[s] = 0, s_t[s] = 0;
        const int n_iter = warp_max(active && !hard ? l_hi - l_lo + 1 : 0);
        uint32_t e[8];
        uint32_t tag = 0, bk = 0;
        if (active && !hard) {
            const uint64_t h = mix64(sbits64(own, 2u * (uint32_t) (l_hi - P.seed_nt)) & P.seed_mask);
            tag = tag_of(T, h);
            bk = bucket_of(h, T.n_buckets);
            load_bucket(T.slots + (uint64_t) bk * kSlotsPerBucket, e);
        }
        for (int it = 0; it < n_iter; it++) {
            const int32_t L = l_hi - it;
            const bool live = active && !hard && L >= l_lo;
            if (!__any_sync(kFull, live)) break;
            uint32_t b0 = kNone, b1 = kNone;
            int n = 0;
            if (live) finish_probe(T, e, tag, bk, b0, b1, n);
            if (live && L - 1 >= l_lo) {
                const uint64_t h = mix64(sbits64(own, 2u * (uint32_t) (L - 1 - P.seed_nt)) & P.seed_mask);
                tag = tag_of(T, h);
                bk = bucket_of(h, T.n_buckets);
                load_bucket(T.slots + (uint64_t) bk * kSlotsPerBucket, e);
            }
            if (n > 2) hard = true;
            if (n == 2 && b1 > b0) {  // walking backwards: within one length the larger source id arrived later
                const uint32_t x = b0;
                b0 = b1;
                b1 = x;
            }
            const int n_pass = __any_sync(kFull, live && !hard && n > 1) ? 2 : 1;
            for (int k = 0; k < n_pass; k++) {
                const uint32_t cand = k ? b1 : b0;
                bool want = false;
                uint32_t o = 0, lenb = lenc;
                uint64_t t = 0;
                if (live && !hard && n > k && cand != c) {
                    if (!UNIFORM) lenb = R.len[cand];
                    if ((int64_t) lenb - P.min_offset >= L) {
                        o = lenb - (uint32_t) L;
                        if (o > 32u || n_ids >= kIdCap) {
                            hard = true;
                        } else {
                            ids[n_ids * kTpr] = cand;
                            n_ids++;
                            t = overhang_tail(read_ptr(R, cand), o);
                            bool removed = false;
#pragma unroll
                            for (int s = 0; s < kSurv; s++) {
                                if (s < ns && s_o[s] > 0 && o >= s_o[s] &&
                                    (UNIFORM || (int64_t) s_len[s] + (int64_t) (o - s_o[s]) - (int64_t) lenb >= 0) &&
                                    ((t ^ s_t[s]) >> (64u - 2u * s_o[s])) == 0)
                                    removed = true;
                            }
                            want = !removed;
                        }
                    }
                }
                const bool ok = coop_verify_prefix(R, wown, wp, q, want, cand, o, L, lane);
                if (ok) {
                    if (ns >= kSurv) {
                        hard = true;
                    } else {
#pragma unroll
                        for (int s = 0; s < kSurv; s++)
                            if (s == ns) s_id[s] = cand, s_o[s] = o, s_len[s] = lenb, s_t[s] = t;
                        ns++;
                    }
                }
            }
        }
        // in-neighbours from phase 1 (row of the transposed graph): kept unless a surviving arrival removes them
        uint32_t rowmask = 0;
        if (part && !hard) {
            for (uint32_t r = 0; r < deg; r++) {
                const RevEntry en = rev[r0 + r];
                const uint32_t a = (uint32_t) en.b, oa = (uint32_t) en.o;
                for (int k = 0; k < n_ids; k++)
                    if (ids[k * kTpr] == a) hard = true;  // the same read twice for one target: generic path
                if (n_ids >= kIdCap) {
                    hard = true;
                    break;
                }
                ids[n_ids * kTpr] = a;
                n_ids++;
                const uint64_t ta = en.t;
                const uint32_t lena = UNIFORM ? lenc : R.len[a];
                bool removed = false;
#pragma unroll
                for (int s = 0; s < kSurv; s++) {
                    if (s < ns && s_o[s] > 0 && oa >= s_o[s] &&
                        (UNIFORM || (int64_t) s_len[s] + (int64_t) (oa - s_o[s]) - (int64_t) lena >= 0) &&
                        ((ta ^ s_t[s]) >> (64u - 2u * s_o[s])) == 0)
                        removed = true;
                }
                if (!removed) rowmask |= 1u << r;
            }
            // a surviving arrival whose read occurs a second time among the candidates (same-id replacement rule)
            for (int k = 0; k < n_ids; k++) {
                const uint32_t v = ids[k * kTpr];
                int cnt = 0;
#pragma unroll
                for (int s = 0; s < kSurv; s++) cnt += (s < ns && s_id[s] == v) ? 1 : 0;
                if (cnt) {
                    for (int k2 = k + 1; k2 < n_ids; k2++)
                        if (ids[k2 * kTpr] == v) hard = true;
                }
            }
        }
        // emit: one atomicAdd on the edge counter per warp
        uint32_t n_out = (part && !hard) ? (uint32_t) ns + __popc(rowmask) : 0u;
        uint32_t incl = n_out;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(kFull, incl, d);
            if (lane >= d) incl += y;
        }
        const uint32_t total = __shfl_sync(kFull, incl, 31);
        unsigned long long base = 0;
        if (total) {
            if (lane == 31) base = atomicAdd(out.n_edges, (unsigned long long) total);
            base = __shfl_sync(kFull, base, 31);
        }
        if (n_out) {
            unsigned long long pos = base + incl - n_out;
#pragma unroll
            for (int s = 0; s < kSurv; s++) {
                if (s < ns) {
                    if (pos < out.edge_cap) {
                        out.triples[3 * pos] = (int32_t) s_id[s];
                        out.triples[3 * pos + 1] = (int32_t) c;
                        out.triples[3 * pos + 2] = (int32_t) s_o[s];
                    }
                    if (out.outdeg) atomicAdd(out.outdeg + s_id[s], 1u);
                    pos++;
                }
            }
            for (uint32_t m = rowmask; m; m &= m - 1) {
                const RevEntry en = rev[r0 + (__ffs(m) - 1)];
                if (pos < out.edge_cap) {
                    out.triples[3 * pos] = en.b;
                    out.triples[3 * pos + 1] = (int32_t) c;
                    out.triples[3 * pos + 2] = en.o;
                }
                if (out.outdeg) atomicAdd(out.outdeg + (uint32_t) en.b, 1u);
                pos++;
            }
        }
        if (hard) out.spill_queue[atomicAdd(out.n_spill, 1u)] = c;
    }
}

inline int smem_words_per_read(uint32_t max_len_nt, int need_words) {
    int w = (int) ((max_len_nt + 15u) >> 4);
    if (w > kOwnWords) w = kOwnWords;
    if (need_words > 0 && w > need_words) w = need_words;
    return (w + 2) | 1;  // two pad words; odd stride: lanes of a warp fall into distinct banks
}

}  // namespace

void launch_phase1_tpr(const ReadsDev &R, const SeedTable &prefix, const PsDev &P, uint32_t max_len_nt, uint32_t lo,
                       uint32_t hi, int2 *fwd, uint64_t *fwd_t, uint32_t *fwd_pos, uint32_t *indeg, uint32_t *hard_queue,
                       uint32_t *n_hard, int force_hard, cudaStream_t s, const LaunchCfg &cfg) {
    if (hi <= lo) return;
    const int wp = smem_words_per_read(max_len_nt, 0);
    int64_t lmax = (int64_t) P.rs - 1;
    if (lmax > P.max_l) lmax = P.max_l;
    if (lmax > (int64_t) max_len_nt) lmax = max_len_nt;
    if (lmax < 1) lmax = 1;
    if (lmax > kOwnWords * 16) lmax = kOwnWords * 16;
    const int nw_max = (int) ((2 * lmax + 31) >> 5);
    const size_t smem = (size_t) kTpr * wp * sizeof(uint32_t);
    const int grid = tile_grid(hi - lo, cfg, 6);
    if (P.uniform_len) {
        cudaFuncSetAttribute(phase1_tpr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        phase1_tpr_kernel<true><<<grid, kTpr, smem, s>>>(R, prefix, P, lo, hi, wp, nw_max, fwd, fwd_t, fwd_pos, indeg,
                                                         hard_queue, n_hard, force_hard);
    } else {
        cudaFuncSetAttribute(phase1_tpr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        phase1_tpr_kernel<false><<<grid, kTpr, smem, s>>>(R, prefix, P, lo, hi, wp, nw_max, fwd, fwd_t, fwd_pos, indeg,
                                                          hard_queue, n_hard, force_hard);
    }
    bump(cfg);
}

void launch_phase2_tpr(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t max_len_nt, uint32_t lo,
                       uint32_t hi, const uint32_t *rev_off, const RevEntry *rev, const Phase2Out &out, int force_hard,
                       cudaStream_t s, const LaunchCfg &cfg) {
    if (hi <= lo) return;
    int64_t lmax = P.max_l;
    if (lmax > (int64_t) max_len_nt) lmax = max_len_nt;
    if (lmax < 1) lmax = 1;
    const int nw_max = (int) ((2 * lmax + 31) >> 5);
    const int wp = smem_words_per_read(max_len_nt, nw_max);
    const size_t smem = (size_t) kTpr * (wp + kIdCap) * sizeof(uint32_t);
    const int grid = tile_grid(hi - lo, cfg, 5);
    if (P.uniform_len) {
        cudaFuncSetAttribute(phase2_tpr_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        phase2_tpr_kernel<true><<<grid, kTpr, smem, s>>>(R, suffix, P, lo, hi, wp, nw_max, rev_off, rev, out,
                                                         force_hard);
    } else {
        cudaFuncSetAttribute(phase2_tpr_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        phase2_tpr_kernel<false><<<grid, kTpr, smem, s>>>(R, suffix, P, lo, hi, wp, nw_max, rev_off, rev, out,
                                                          force_hard);
    }
    bump(cfg);
}

}  // namespace alga

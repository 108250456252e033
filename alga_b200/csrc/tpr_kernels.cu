// Thread-per-read overlap search kernels (sm_100a): the fast path of GraphCreatorPrefSuf phase 1 and phase 2.
//
// One THREAD owns one read, one WARP owns a tile of 32 consecutive reads (no block-level barriers).  The work of a
// read is split into two loops with very different instruction mixes:
//
//   probe    walk the overlap lengths, one 32-byte bucket of the seed index per length.  The K-nucleotide seed
//            window slides through a 96-bit register window over the read staged in shared memory (two funnel
//            shifts per length), the hash is four IMADs.  Random buckets must not pass through L1 (filling its
//            lines with them halves the rate: 2.5 vs 1.4 ms in phase 2) and are the one structure worth keeping in
//            L2 (evict_last).  Phase 2 fetches them by cp.async.cg (2 x 16 bytes) into a per-thread ring in shared
//            memory kRing2 lengths ahead of their use, so kRing2 random sectors per thread are in flight without
//            costing registers; phase 1, whose lanes stop after about a third of their lengths, keeps one bucket in
//            flight in registers (LDG.E.NA.256).  A bucket is tested with eight XOR + a min tree (an entry with the
//            right tag XORs to its bare read id, everything else to something larger).  Tag hits (0.3 per length)
//            are only QUEUED; in phase 2 the first 64 bits of the hit read (all its overhang tail needs) follow by
//            cp.async as well.
//   resolve  take the queued hits in order.  Exact 2-bit compares of whole overlaps are thread-local: every lane
//            verifies its own candidate, words of the candidate straight from L1/L2, words of the own read from
//            shared memory -- at 3 candidates per read in phase 1 nearly all lanes are busy.
//
// Two passes: the first over all reads allows two tag matches per window; reads that see more (same start position,
// different sequencing errors) are queued, and a queue of at least kSecondPassMin reads is run again with four
// matches per window (template parameter MAXM) before the rest goes to the generic kernels.
//
// History (profiles/): the first thread-per-read version interleaved probe and resolve per
// length and verified with warp-cooperative groups; ncu showed it issue-bound at 300-440 warp instructions per
// (warp, length) with 9 of 32 lanes inside the hit branch (r01d).  Splitting the loops (r01h) halved the
// instructions and left the kernels latency-bound on the bucket loads (27 % of the stall samples at 20 warps per
// SM, one bucket in flight per thread); the cp.async ring (r01i) removed that.
//
// Reads the fast path cannot take at all (longer than 512 nt, a window with more than four tag matches, offsets above
// 32, more queued arrivals than fit, a source id that occurs twice for one target, ...) go to the generic kernels of
// prefsuf_kernels.cu, which replay GraphCreatorPrefSuf.cpp:356-488 literally.
#include <algorithm>

#include "launch.h"

namespace alga {

namespace {

constexpr int kTpr = 128;     // threads per block = 4 independent warps
constexpr int kWarps = kTpr / 32;
// tuning knobs (scripts/build_variant.sh builds A/B variants with -D...)
#ifndef ALGA_P1_BLOCKS
#define ALGA_P1_BLOCKS 5
#endif
#ifndef ALGA_P2_BLOCKS
#define ALGA_P2_BLOCKS 4
#endif
#ifndef ALGA_Q2
#define ALGA_Q2 20
#endif
#ifndef ALGA_RING2
#define ALGA_RING2 3
#endif
constexpr int kQ2 = ALGA_Q2;  // arrivals queued per target in phase 2
constexpr int kSurv = 4;      // surviving arrivals kept in registers per target
constexpr int kRowFast = 32;  // longest transposed row the fast phase-2 kernel takes
constexpr int kRing2 = ALGA_RING2;     // buckets in flight per thread in phase 2 (static ring index: the loop is unrolled)

inline int warp_tile_grid(uint64_t n_items, const LaunchCfg &cfg, int blocks_per_sm) {
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

// ---- asynchronous copies ------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t l2_evict_last_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void cp_async8(uint64_t *smem_dst, const uint32_t *gmem_src) {  // 8-byte aligned on both sides
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- bucket evaluation ----------------------------------------------------------------------------------------
// An entry with the probed tag XORs to its bare read id (<= id_mask), any other entry to something larger: the
// minimum over the bucket is the smallest matching id, if there is one.
__device__ __forceinline__ uint32_t bucket_min(const uint32_t (&e)[8], uint32_t tag) {
    uint32_t m = e[0] ^ tag;
#pragma unroll
    for (int s = 1; s < kSlotsPerBucket; s++) m = min(m, e[s] ^ tag);
    return m;
}

// ids of the entries of one bucket whose tag matches (the first kMaxMatch) and how many there are; true if the bucket
// is full
// (MAXM = reads that share a K-nucleotide seed and overlap at one length, i.e. start at the same position: 2 in the
// first pass over all reads, 4 in the second pass over the reads the first one gave up on)
template <int MAXM>
__device__ __forceinline__ bool eval_bucket(const SeedTable &t, const uint32_t (&e)[8], uint32_t tag, uint32_t (&c)[MAXM],
                                            int &n) {
#pragma unroll
    for (int s = 0; s < kSlotsPerBucket; s++) {
        if ((e[s] ^ tag) <= t.id_mask) {
#pragma unroll
            for (int k = 0; k < MAXM; k++)
                if (k == n) c[k] = e[s] & t.id_mask;
            n++;
        }
    }
    return e[kSlotsPerBucket - 1] != kEmptySlot;  // buckets fill front to back: a full one chains on
}

__device__ __forceinline__ void order_desc(uint32_t &a, uint32_t &b) {
    const uint32_t hi = max(a, b), lo = min(a, b);
    a = hi, b = lo;
}

// Matches of a probe whose first bucket is in e[] and holds at least one match or is full, largest id first (the order
// of arrival inside one overlap length, seen backwards).  The common case -- one match, bucket not full -- is answered
// by the minimum alone; otherwise collect the matches and walk the chain.  n > MAXM: the caller gives up.
template <int MAXM>
__device__ __forceinline__ void probe_matches(const SeedTable &t, uint32_t (&e)[8], uint32_t tag, uint32_t bk, uint32_t m,
                                              uint32_t (&c)[MAXM], int &n) {
    int cnt = 0;
#pragma unroll
    for (int s = 0; s < kSlotsPerBucket; s++) cnt += ((e[s] ^ tag) <= t.id_mask) ? 1 : 0;
    const bool full = e[kSlotsPerBucket - 1] != kEmptySlot;
#pragma unroll
    for (int k = 0; k < MAXM; k++) c[k] = 0u;
    if (cnt == 1 && !full) {
        n = 1;
        c[0] = m;
        return;
    }
    n = 0;
    bool more = eval_bucket<MAXM>(t, e, tag, c, n);
    while (more) {
        bk = next_bucket(t, bk);
        load_bucket(t.slots + (uint64_t) bk * kSlotsPerBucket, e);
        more = eval_bucket<MAXM>(t, e, tag, c, n);
    }
    if (n > 1 && n <= MAXM) {  // unused slots hold 0 and sink to the end (n says how many are real)
        if (MAXM == 2) {
            order_desc(c[0], c[1]);
        } else {
            static_assert(MAXM == 2 || MAXM == 4, "sorting network");
            order_desc(c[0], c[1]);
            order_desc(c[MAXM - 2], c[MAXM - 1]);
            order_desc(c[0], c[MAXM - 2]);
            order_desc(c[1], c[MAXM - 1]);
            order_desc(c[1], c[MAXM - 2]);
        }
    }
}

// 64 bits of the staged read starting `s` bits into the register window (w0, w1, w2)
__device__ __forceinline__ uint64_t window_key(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t s) {
    return (uint64_t) __funnelshift_r(w0, w1, s) | ((uint64_t) __funnelshift_r(w1, w2, s) << 32);
}

// Stage the (up to 32) reads of a warp's tile: `sw` words each at stride `wp`, the rest of the stride zeroed.
// FAST (fixed stride in HBM): the tile is one contiguous range, copied with fully coalesced loads.
// `listed`: the tile is not a contiguous id range (second pass over a queue): every lane copies its own read `my_id`.
template <bool FAST>
__device__ __forceinline__ void stage_warp(const ReadsDev &R, uint32_t *wown, int wp, int sw, uint64_t first,
                                           uint32_t n_valid, uint32_t my_words, int lane, bool listed = false,
                                           uint32_t my_id = 0) {
    uint32_t *own = wown + lane * wp;
    __syncwarp();
    if (listed) {
        const uint32_t *p = (uint32_t) lane < n_valid ? read_ptr(R, my_id) : R.words;
        const uint32_t nw = FAST ? ((uint32_t) lane < n_valid ? (uint32_t) sw : 0u) : my_words;
        for (int w = 0; w < wp; w++) own[w] = (uint32_t) w < nw ? __ldg(p + w) : 0u;
    } else if (FAST) {
        for (int w = min(sw, (int) R.stride); w < wp; w++) own[w] = 0u;
        const uint32_t *base = R.words + first * R.stride;
        if ((R.stride & 3u) == 0 && ((uintptr_t) base & 15u) == 0) {
            // sector-aligned slots: the tile is one contiguous range, 16 bytes per lane and load
            const uint32_t q_per_read = R.stride >> 2, total = n_valid * q_per_read;
            const uint4 *b4 = reinterpret_cast<const uint4 *>(base);
            for (uint32_t j = lane; j < total; j += 32) {
                const uint32_t r = j / q_per_read, w = (j - r * q_per_read) << 2;
                if (w < (uint32_t) sw) {
                    const uint4 v = __ldg(b4 + j);
                    uint32_t *d = wown + r * wp + w;
                    d[0] = v.x;
                    if (w + 1 < (uint32_t) sw) d[1] = v.y;
                    if (w + 2 < (uint32_t) sw) d[2] = v.z;
                    if (w + 3 < (uint32_t) sw) d[3] = v.w;
                }
            }
        } else if ((uint32_t) sw == R.stride) {
            const uint32_t total = n_valid * (uint32_t) sw;
            uint32_t r = (uint32_t) lane / (uint32_t) sw, w = (uint32_t) lane - r * (uint32_t) sw;
            const uint32_t dr = 32u / (uint32_t) sw, dw = 32u - dr * (uint32_t) sw;
            for (uint32_t j = lane; j < total; j += 32) {
                wown[r * wp + w] = __ldg(base + j);
                r += dr;
                w += dw;
                if (w >= (uint32_t) sw) {
                    w -= sw;
                    r++;
                }
            }
        } else {
            if ((uint32_t) lane < n_valid)
                for (int w = 0; w < sw; w++) own[w] = __ldg(base + (uint64_t) lane * R.stride + w);
        }
    } else {
        const uint32_t *p = (uint32_t) lane < n_valid ? read_ptr(R, (uint32_t) (first + lane)) : R.words;
        for (int w = 0; w < wp; w++) own[w] = (uint32_t) w < my_words ? __ldg(p + w) : 0u;
    }
    __syncwarp();
}

// prefix(cand, L) == own[o .. o + L) ?   (phase 1: suffix of the own read against the prefix of the candidate)
__device__ __forceinline__ bool verify_own_suffix(const ReadsDev &R, const uint32_t *own, uint32_t cand, uint32_t o,
                                                  int32_t L) {
    const uint32_t *pc = read_ptr(R, cand);
    const uint32_t nbits = 2u * (uint32_t) L, nw = (nbits + 31u) >> 5, sh = (2u * o) & 31u;
    const uint32_t *ow = own + ((2u * o) >> 5);
    for (uint32_t k0 = 0; k0 < nw; k0 += 4) {
        uint32_t g[4];
#pragma unroll
        for (int j = 0; j < 4; j++) g[j] = k0 + j < nw ? __ldg(pc + k0 + j) : 0u;
        uint32_t diff = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t k = k0 + j;
            if (k < nw) {
                uint32_t x = __funnelshift_r(ow[k], ow[k + 1], sh) ^ g[j];
                if (k == nw - 1 && (nbits & 31u)) x &= (1u << (nbits & 31u)) - 1u;
                diff |= x;
            }
        }
        if (diff) return false;
    }
    return true;
}

// cand[o .. o + L) == own[0 .. L) ?   (phase 2: suffix of the candidate against the prefix of the own read)
__device__ __forceinline__ bool verify_own_prefix(const ReadsDev &R, const uint32_t *own, uint32_t cand, uint32_t o,
                                                  int32_t L) {
    const uint32_t *pb = read_ptr(R, cand) + ((2u * o) >> 5);
    const uint32_t nbits = 2u * (uint32_t) L, nw = (nbits + 31u) >> 5, sh = (2u * o) & 31u;
    for (uint32_t k0 = 0; k0 < nw; k0 += 4) {
        uint32_t g[5];
#pragma unroll
        for (int j = 0; j < 5; j++) g[j] = k0 + j <= nw ? __ldg(pb + k0 + j) : 0u;
        uint32_t diff = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t k = k0 + j;
            if (k < nw) {
                uint32_t x = __funnelshift_r(g[j], g[j + 1], sh) ^ own[k];
                if (k == nw - 1 && (nbits & 31u)) x &= (1u << (nbits & 31u)) - 1u;
                diff |= x;
            }
        }
        if (diff) return false;
    }
    return true;
}

// The same test for sector-aligned read slots (fast path): the candidate's words arrive as whole 32-byte sectors, one
// request each, and stay unshifted in registers; the OWN read (shared memory, any index is cheap there) is shifted up
// by the offset instead.  Word j of the candidate lines up with bits [32 j - 2 o, +32) of the own read.
__device__ __forceinline__ bool verify_own_prefix_aligned(const ReadsDev &R, const uint32_t *own, uint32_t cand, uint32_t o,
                                                          int32_t L) {
    const uint32_t *pb = R.words + (uint64_t) cand * R.stride;
    const int32_t nb = 2 * L, sh2 = 2 * (int32_t) o;
    const int32_t n_words = (sh2 + nb + 31) >> 5;  // candidate words that hold compared bits
    for (int32_t s0 = 0; s0 < n_words; s0 += 16) {  // two sectors per round, both requested before the first compare
        uint32_t g[2][8];
        load8_na(pb + s0, g[0]);
        if (s0 + 8 < n_words) load8_na(pb + s0 + 8, g[1]);
        uint32_t diff = 0;
#pragma unroll
        for (int h = 0; h < 2; h++) {
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const int32_t p = 32 * (s0 + 8 * h + j) - sh2;  // own-read bit that meets bit 0 of this candidate word
                if (s0 + 8 * h < n_words && p + 32 > 0 && p < nb) {
                    uint32_t cw, mask = 0xFFFFFFFFu;
                    if (p >= 0) {
                        const uint32_t w = (uint32_t) p >> 5;
                        cw = __funnelshift_r(own[w], own[w + 1], (uint32_t) p & 31u);
                    } else {
                        cw = own[0] << (uint32_t) (-p);
                        mask <<= (uint32_t) (-p);
                    }
                    if (nb - p < 32) mask &= (1u << (uint32_t) (nb - p)) - 1u;
                    diff |= (g[h][j] ^ cw) & mask;
                }
            }
        }
        if (diff) return false;
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// Phase 1 (GraphCreatorPrefSuf.cpp:397-402 in closed form): source read b walks L from min(rs-1, len) downwards
// and keeps the first 3 confirmed (L, c) -- within one L the larger c first -- = "the last 3 pushes".
//
// Lanes pause once they hold 3 candidates, so their positions in the walk differ; with so few lengths per read
// (about 11 of the 34 possible) a deep prefetch ring mostly fetches buckets nobody tests, so this kernel keeps ONE
// bucket in flight per lane, in registers.  Shared memory per warp: own reads [32][wp].
// id_list != nullptr: second pass -- the reads are id_list[0 .. *n_list) instead of [lo, hi)
template <bool FAST, int MAXM, int MINI>  // MINI: 0 window-hash buckets, 1 minimizer buckets, 2 ... with a sliding minimum
__global__ void __launch_bounds__(kTpr, ALGA_P1_BLOCKS)
phase1_tpr_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, const uint32_t *__restrict__ id_list,
                  const uint32_t *__restrict__ n_list, int wp, Phase1Out out, uint32_t *__restrict__ hard_queue,
                  uint32_t *n_hard, int force_hard) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    uint32_t *wown = smem + wib * 32 * wp;
    const uint32_t *own = wown + lane * wp;
    const bool listed = id_list != nullptr;
    const uint64_t n_items = listed ? (uint64_t) *n_list : (uint64_t) (hi - lo);
    if (listed && n_items < kSecondPassMin) return;  // a short queue is cheaper in the generic kernel (one tile = 40 us)
    const uint64_t pol = l2_evict_last_policy();
    const uint64_t n_tiles = (n_items + 31) / 32;
    const uint64_t warp_id = (uint64_t) blockIdx.x * kWarps + wib, n_warps = (uint64_t) gridDim.x * kWarps;
    for (uint64_t tile = warp_id; tile < n_tiles; tile += n_warps) {
        const uint64_t first = (uint64_t) lo + tile * 32;
        const uint32_t n_valid = (uint32_t) min((uint64_t) 32, n_items - tile * 32);
        const bool inr = (uint32_t) lane < n_valid;
        const uint32_t b = inr ? (listed ? id_list[tile * 32 + lane] : (uint32_t) (first + lane)) : lo;
        const uint32_t lenb = inr ? (FAST ? P.uniform_len : R.len[b]) : 0u;
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
        stage_warp<FAST>(R, wown, wp, wp - 2, first, n_valid, active ? (lenb + 15u) >> 4 : 0u, lane, listed, b);

        int conf = 0, np = 0;
        constexpr int kPend = 2 + MAXM;  // fewer than 3 held when a window is probed, up to MAXM more from it
        constexpr int kBatch = MAXM == 2 ? 4 : 3;  // candidates verified together (their loads overlap)
        uint32_t sc[kSmallEdgesKept], so[kSmallEdgesKept], pc[kPend];
        int32_t pl[kPend];
#pragma unroll
        for (int k = 0; k < kSmallEdgesKept; k++) sc[k] = kNone, so[k] = 0;
#pragma unroll
        for (int k = 0; k < kPend; k++) pc[k] = 0, pl[k] = 0;

        // probe state: the bucket of length `Lc` is in flight in e[]
        int32_t Lc = l_hi;
        bool more = active;
        uint32_t e[8], tag = 0, bk = 0, w0 = 0, w1 = 0, w2 = 0, sh = 0;
        int wb = 0;
        SlidingMinimizer smin;  // MINI == 2 only: the window moves up by one nucleotide per length
        smin.best = 0, smin.pos = 0;
        if (more) {
            const uint32_t p = 2u * (lenb - (uint32_t) Lc);
            wb = (int) (p >> 5);
            sh = p & 31u;
            w0 = own[wb], w1 = own[wb + 1], w2 = own[wb + 2];
            const uint64_t win = window_key(w0, w1, w2, sh) & P.seed_mask, h = mix64(win);
            tag = tag_of(T, h);
            if (MINI == 2) {
                smin.reset(win, (uint32_t) P.seed_nt, T.min_m);
                bk = bucket_of(mix64((uint64_t) smin.best), T.n_buckets);
            } else {
                bk = bucket_index<MINI>(T, win, h, (uint32_t) P.seed_nt);
            }
            load_bucket_na(T.slots + (uint64_t) bk * kSlotsPerBucket, e, pol);
        }
        while (true) {
            // ---- probe until every lane has 3 candidates (confirmed + pending) or ran out of lengths
            while (true) {
                const bool need = more && !hard && conf + np < kSmallEdgesKept;
                if (!__any_sync(kFull, need)) break;
                if (need) {
                    const uint32_t m = bucket_min(e, tag);
                    uint32_t cm[MAXM];
                    int n = 0;
                    const bool e7_walked = MINI == 2 && e[7] != kEmptySlot;  // a full bucket: probe_matches may load its chain into e[]
                    if (m <= T.id_mask || e[7] != kEmptySlot) probe_matches<MAXM>(T, e, tag, bk, m, cm, n);
                    const int32_t L = Lc;
                    Lc--;
                    if (Lc >= P.lmin) {  // slide the window by one nucleotide, next bucket goes in flight
                        sh += 2u;
                        if (sh == 32u) {
                            sh = 0u;
                            wb++;
                            w0 = w1, w1 = w2, w2 = own[wb + 2];
                        }
                        const uint64_t win = window_key(w0, w1, w2, sh) & P.seed_mask, h = mix64(win);
                        tag = tag_of(T, h);
                        if (MINI == 2) {
                            const uint32_t bk_new = bucket_of(mix64((uint64_t) smin.slide_up(win, (uint32_t) P.seed_nt, T.min_m)), T.n_buckets);
                            // the same bucket as for the previous length (the usual case) is still in e[]: nothing to fetch,
                            // unless probe_matches walked its chain and left another bucket there
                            if (bk_new != bk || e7_walked) {
                                bk = bk_new;
                                load_bucket_na(T.slots + (uint64_t) bk * kSlotsPerBucket, e, pol);
                            }
                        } else {
                            bk = bucket_index<MINI>(T, win, h, (uint32_t) P.seed_nt);
                            load_bucket_na(T.slots + (uint64_t) bk * kSlotsPerBucket, e, pol);
                        }
                    } else {
                        more = false;
                    }
                    if (n > MAXM) {
                        hard = true;
                    } else if (n) {  // within one length the larger target id is the later push: cm[] is descending
#pragma unroll
                        for (int k = 0; k < kPend; k++)
                            if (k == np) pc[k] = cm[0], pl[k] = L;
                        np++;
                        if (n > 1) {  // several reads start with this seed (rare without sequencing errors)
#pragma unroll
                            for (int j = 1; j < MAXM; j++) {
                                if (j < n) {
#pragma unroll
                                    for (int k = 0; k < kPend; k++)
                                        if (k == np) pc[k] = cm[j], pl[k] = L;
                                    np++;
                                }
                            }
                        }
                    }
                }
            }
            // ---- confirm the pending candidates, every lane its own, kBatch at a time: the words of a whole batch are
            // requested before the first compare, so their (random, mostly DRAM) latencies overlap
            const int np_max = warp_max(hard ? 0 : np);
            if (np_max == 0) break;
            const int nw_max = warp_max(np && !hard ? (2 * pl[0] + 31) >> 5 : 0);
#pragma unroll
            for (int base = 0; base < kPend; base += kBatch) {
                if (base < np_max) {
                    uint32_t diff[kBatch];
#pragma unroll
                    for (int k = 0; k < kBatch; k++) diff[k] = 0u;
                    // FAST: read slots are sector aligned, eight words of a candidate = ONE 32-byte request (a prefix of up
                    // to 128 nucleotides is a single request; 4-byte loads of a sector that is still on its way from DRAM
                    // would each fetch it again)
                    constexpr int kChunk = FAST ? 8 : 4;
                    for (int k0 = 0; k0 < nw_max; k0 += kChunk) {
                        uint32_t g[kBatch][kChunk];
#pragma unroll
                        for (int k = 0; k < kBatch; k++) {
                            if (base + k < np_max) {
                                const bool on = base + k < np && !hard;
                                const uint32_t *pcand = read_ptr(R, on ? pc[base + k] : b);
                                const int nw = on ? (2 * pl[base + k] + 31) >> 5 : 0;
                                if constexpr (FAST) {
#pragma unroll
                                    for (int j = 0; j < kChunk; j++) g[k][j] = 0u;
                                    if (k0 < nw) load8_na(pcand + k0, reinterpret_cast<uint32_t(&)[8]>(g[k]));
                                } else {
#pragma unroll
                                    for (int j = 0; j < kChunk; j++) g[k][j] = k0 + j < nw ? __ldg(pcand + k0 + j) : 0u;
                                }
                            }
                        }
#pragma unroll
                        for (int k = 0; k < kBatch; k++) {
                            if (base + k < np_max && base + k < np && !hard) {
                                const uint32_t nbits = 2u * (uint32_t) pl[base + k], nw = (nbits + 31u) >> 5;
                                const uint32_t o2 = 2u * (lenb - (uint32_t) pl[base + k]), shv = o2 & 31u;
                                const uint32_t *ow = own + (o2 >> 5);
#pragma unroll
                                for (int j = 0; j < kChunk; j++) {
                                    const uint32_t w = (uint32_t) (k0 + j);
                                    if (w < nw) {
                                        uint32_t x = __funnelshift_r(ow[w], ow[w + 1], shv) ^ g[k][j];
                                        if (w == nw - 1 && (nbits & 31u)) x &= (1u << (nbits & 31u)) - 1u;
                                        diff[k] |= x;
                                    }
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int k = 0; k < kBatch; k++) {
                        if (base + k < np && !hard && conf < kSmallEdgesKept) {
                            const uint32_t cand = pc[base + k];
                            const int32_t L = pl[base + k];
                            if (diff[k] == 0 && cand != b && (FAST || (int64_t) R.len[cand] >= L)) {
#pragma unroll
                                for (int q = 0; q < kSmallEdgesKept; q++)
                                    if (q == conf) sc[q] = cand, so[q] = lenb - (uint32_t) L;
                                conf++;
                            }
                        }
                    }
                }
            }
            np = 0;
        }

        // ---- emit
        if (out.mode == 0) {
            if (!hard) {
                uint32_t pos[kSmallEdgesKept];
#pragma unroll
                for (int k = 0; k < kSmallEdgesKept; k++)
                    pos[k] = k < conf ? atomicAdd(out.indeg + (sc[k] - out.c_base), 1u) : 0u;
#pragma unroll
                for (int k = 0; k < kSmallEdgesKept; k++) {
                    if (k < conf) {
                        const uint32_t c = sc[k];
                        RevEntry r;
                        r.b = (int32_t) b;
                        r.o = (int32_t) so[k];
                        r.t = overhang_tail_own(own, so[k]);
                        if (pos[k] < out.row_cap) {
                            out.rows[(uint64_t) (c - out.c_base) * out.row_cap + pos[k]] = r;
                        } else {
                            const uint32_t i = atomicAdd(out.n_list, 1u);
                            if (i < out.list_cap) {
                                Edge1 x;
                                x.c = (int32_t) (c - out.c_base), x.b = r.b, x.o = r.o, x.pad = 0, x.t = r.t;
                                out.list[i] = x;
                            }
                        }
                    }
                }
            }
        } else if (out.mode == 2) {
            // sharded: every edge goes to the segment of the rank that owns its target read
#pragma unroll
            for (int k = 0; k < kSmallEdgesKept; k++) {
                const bool valid = !hard && k < conf;
                if (__any_sync(kFull, valid)) {
                    const uint32_t d = valid ? shard_of(out.sh, sc[k]) : 0u;
                    const uint32_t pos = shard_reserve(out.sh, valid, d, lane);
                    if (valid && pos < out.sh.cap) {
                        Edge1 x;
                        x.c = (int32_t) sc[k], x.b = (int32_t) b, x.o = (int32_t) so[k], x.pad = 0;
                        x.t = overhang_tail_own(own, so[k]);
                        reinterpret_cast<Edge1 *>(out.sh.seg)[(uint64_t) d * out.sh.cap + pos] = x;
                    }
                }
            }
        } else {
            const uint32_t n_out = hard ? 0u : (uint32_t) conf;
            uint32_t incl = n_out;
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += y;
            }
            const uint32_t total = __shfl_sync(kFull, incl, 31);
            uint32_t base = 0;
            if (total) {
                if (lane == 31) base = atomicAdd(out.n_list, total);
                base = __shfl_sync(kFull, base, 31);
            }
            uint32_t pos = base + incl - n_out;
#pragma unroll
            for (int k = 0; k < kSmallEdgesKept; k++) {
                if ((uint32_t) k < n_out) {
                    if (pos < out.list_cap) {
                        Edge1 x;
                        x.c = (int32_t) sc[k], x.b = (int32_t) b, x.o = (int32_t) so[k], x.pad = 0;
                        x.t = overhang_tail_own(own, so[k]);
                        out.list[pos] = x;
                    }
                    pos++;
                }
            }
        }
        if (hard) hard_queue[atomicAdd(n_hard, 1u)] = b;
    }
}

// ------------------------------------------------------------------------------------------------------------
// Phase 2 (GraphCreatorPrefSuf.cpp:403-483 in closed form): target read c walks L from len downwards, i.e. from its
// LAST arrival to its first.  An arrival (b, o) stays in c's list unless a later arrival j with o_j > 0 has an
// overhang that is a suffix of b's overhang (a[oa-oj .. oa) == b_j[0 .. oj), right offset >= 0); the relation is
// transitive, so it is enough to test against the arrivals that survived so far, and a candidate that fails this
// test is dropped without ever comparing its overlap.  Entries of the transposed phase-1 graph (rows) are tested
// against the survivors at the end.
//
// Shared memory per warp: own reads [32][wp] | queue: ids [kQ2][32], heads [kQ2][32] (u64), lengths [kQ2][32] (u16).
// The buckets of the next kRing2 lengths are in flight in REGISTERS, one 32-byte request each (LDG.E.256): two 16-byte
// cp.async of one sector are two requests, and L2 fetches a sector that is still on its way from DRAM once per request
// (scripts/probes/random_requests.cu: 19.8 vs 39.4 G buckets/s).
// id_list != nullptr: second pass -- the targets are id_list[0 .. *n_list) instead of [lo, hi)
template <bool FAST, int MAXM, int MINI>  // MINI: 0 window-hash buckets, 1 minimizer buckets, 2 ... with a sliding minimum
__global__ void __launch_bounds__(kTpr, ALGA_P2_BLOCKS)
phase2_tpr_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, const uint32_t *__restrict__ id_list,
                  const uint32_t *__restrict__ n_list, int wp, RowsView rows, Phase2Out out, int force_hard) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int own_words = ((kWarps * 32 * wp + 3) & ~3);
    constexpr int kQueueWords = kQ2 * 32 * 3 + kQ2 * 32 / 2;
    uint32_t *wown = smem + wib * 32 * wp;
    const uint32_t *own = wown + lane * wp;
    uint32_t *q_id = smem + own_words + wib * kQueueWords + lane;                                // q_id[k * 32]
    uint64_t *q_t = reinterpret_cast<uint64_t *>(q_id - lane + kQ2 * 32) + lane;                 // q_t[k * 32]: first 64 bits of the hit read
    uint16_t *q_l = reinterpret_cast<uint16_t *>(q_id - lane + kQ2 * 32 * 3) + lane;             // q_l[k * 32]
    const uint64_t pol = l2_evict_last_policy();
    const int32_t l_lo = P.rs > P.lmin ? P.rs : P.lmin;
    const bool csr = rows_are_csr(rows);
    const bool listed = id_list != nullptr;
    const uint64_t n_items = listed ? (uint64_t) *n_list : (uint64_t) (hi - lo);
    if (listed && n_items < kSecondPassMin) return;  // a short queue is cheaper in the generic kernel (one tile = 40 us)
    const uint64_t n_tiles = (n_items + 31) / 32;
    const uint64_t warp_id = (uint64_t) blockIdx.x * kWarps + wib, n_warps = (uint64_t) gridDim.x * kWarps;
    for (uint64_t tile = warp_id; tile < n_tiles; tile += n_warps) {
        const uint64_t first = (uint64_t) lo + tile * 32;
        const uint32_t n_valid = (uint32_t) min((uint64_t) 32, n_items - tile * 32);
        const bool inr = (uint32_t) lane < n_valid;
        const uint32_t c = inr ? (listed ? id_list[tile * 32 + lane] : (uint32_t) (first + lane)) : lo;
        uint32_t deg = 0;
        const RevEntry *row = rows.rev;
        if (inr) row = get_row(rows, csr, c - lo, deg);
        const uint32_t lenc = inr ? (FAST ? P.uniform_len : R.len[c]) : 0u;
        const int32_t l_hi = (int32_t) (lenc < (uint32_t) P.max_l ? lenc : (uint32_t) P.max_l);
        const bool active = inr && lenc != 0 && flag_to(R, c) && l_hi >= l_lo;
        const bool part = active || deg > 0;  // has something to emit
        bool hard = part && (force_hard || deg > (uint32_t) kRowFast || (!csr && deg > rows.cap));
        {
            uint32_t nw = 0;
            if (active && !hard) {
                const uint32_t need = (uint32_t) ((2 * l_hi + 31) >> 5), have = (lenc + 15u) >> 4;
                nw = need < have ? need : have;
            }
            stage_warp<FAST>(R, wown, wp, wp - 2, first, n_valid, nw, lane, listed, c);
        }

        // ---- probe: queue every tag hit (b, L); FAST: the first 64 bits of b follow by cp.async
        int qn = 0;
        const bool walk = active && !hard;
        const int n_iter = warp_max(walk ? l_hi - l_lo + 1 : 0);
        uint32_t er[kRing2][8], tagr[kRing2], bkr[kRing2], w0 = 0, w1 = 0, w2 = 0, sh = 0;
        int wb = 0;
        int32_t Lp = l_hi;  // next length to prefetch; (w0, w1, w2, sh, wb) = register window at Lp
        SlidingMinimizer smin;  // MINI == 2 only
        smin.best = 0, smin.pos = 0;
        bool smin_on = false;
        if (walk) {
            const uint32_t p = 2u * (uint32_t) (l_hi - P.seed_nt);
            wb = (int) (p >> 5);
            sh = p & 31u;
            w0 = own[wb], w1 = own[wb + 1], w2 = own[wb + 2];
        }
#pragma unroll
        for (int j = 0; j < kRing2; j++) {
            tagr[j] = 0, bkr[j] = 0;
#pragma unroll
            for (int q = 0; q < 8; q++) er[j][q] = kEmptySlot;
        }
        auto prefetch = [&](uint32_t (&e_out)[8], uint32_t &tag_out, uint32_t &bk_out) {
            if (walk && Lp >= l_lo) {
                const uint64_t win = window_key(w0, w1, w2, sh) & P.seed_mask, h = mix64(win);
                tag_out = tag_of(T, h);
                if (MINI == 2) {  // the window moves down by one nucleotide per call
                    if (smin_on) smin.slide_down(win, (uint32_t) P.seed_nt, T.min_m);
                    else smin.reset(win, (uint32_t) P.seed_nt, T.min_m);
                    smin_on = true;
                    bk_out = bucket_of(mix64((uint64_t) smin.best), T.n_buckets);
                } else {
                    bk_out = bucket_index<MINI>(T, win, h, (uint32_t) P.seed_nt);
                }
                load_bucket_na(T.slots + (uint64_t) bk_out * kSlotsPerBucket, e_out, pol);
                Lp--;
                if (sh == 0u) {  // slide the window down by one nucleotide
                    sh = 32u;
                    wb--;
                    w2 = w1, w1 = w0, w0 = wb >= 0 ? own[wb] : 0u;
                }
                sh -= 2u;
            }
        };
#pragma unroll
        for (int j = 0; j < kRing2; j++) prefetch(er[j], tagr[j], bkr[j]);

        for (int it0 = 0; it0 < n_iter; it0 += kRing2) {
#pragma unroll
            for (int j = 0; j < kRing2; j++) {
                const int32_t L = l_hi - (it0 + j);
                if (walk && L >= l_lo) {
                    uint32_t e[8];
#pragma unroll
                    for (int q = 0; q < 8; q++) e[q] = er[j][q];
                    const uint32_t tag = tagr[j], bk = bkr[j];
                    const uint32_t m = bucket_min(e, tag);
                    uint32_t bm[MAXM];
                    int n = 0;
                    if (!hard && (m <= T.id_mask || e[7] != kEmptySlot)) probe_matches<MAXM>(T, e, tag, bk, m, bm, n);
                    prefetch(er[j], tagr[j], bkr[j]);
                    if (n > MAXM) {
                        hard = true;
                    } else if (n) {  // walking backwards: within one length the larger source id arrived later: bm[] is descending
                        for (int k = 0; k < n; k++) {  // n == 1 unless several reads end with this seed
                            uint32_t cand = bm[0];
#pragma unroll
                            for (int t = 1; t < MAXM; t++)
                                if (t == k) cand = bm[t];
                            if (cand == c) continue;
                            if (qn >= kQ2) {
                                hard = true;
                                break;
                            }
                            q_id[qn * 32] = cand;
                            q_l[qn * 32] = (uint16_t) L;
                            if (FAST) cp_async8(q_t + qn * 32, R.words + (uint64_t) cand * R.stride);  // one request
                            qn++;
                        }
                    }
                }
            }
        }
        cp_async_wait_all();

        // ---- resolve the queued arrivals, last arrival first
        int ns = 0;
        uint32_t s_id[kSurv], s_o[kSurv], s_len[kSurv];
        uint64_t s_t[kSurv];
#pragma unroll
        for (int s = 0; s < kSurv; s++) s_id[s] = kNone, s_o[s] = 0, s_len[s] = 0, s_t[s] = 0;
        const int q_max = warp_max(hard ? 0 : qn);
        for (int k = 0; k < q_max; k++) {
            bool want = false;
            uint32_t cand = kNone, o = 0, lenb = lenc;
            int32_t L = 0;
            uint64_t t = 0;
            if (k < qn && !hard) {
                cand = q_id[k * 32];
                L = (int32_t) q_l[k * 32];
                if (!FAST) lenb = R.len[cand];
                if ((int64_t) lenb - P.min_offset >= L) {
                    o = lenb - (uint32_t) L;
                    if (o > 32u) {
                        hard = true;
                    } else {
                        if (FAST) {
                            const uint64_t head = q_t[k * 32];
                            t = o ? head << (64u - 2u * o) : 0ull;
                        } else {
                            t = overhang_tail(read_ptr(R, cand), o);
                        }
                        bool removed = false;
#pragma unroll
                        for (int s = 0; s < kSurv; s++) {
                            if (s < ns && s_o[s] > 0 && o >= s_o[s] &&
                                (FAST || (int64_t) s_len[s] + (int64_t) (o - s_o[s]) - (int64_t) lenb >= 0) &&
                                ((t ^ s_t[s]) >> (64u - 2u * s_o[s])) == 0)
                                removed = true;
                        }
                        want = !removed;
                    }
                } else {
                    q_id[k * 32] = kNone;  // too short for this length: not an arrival
                }
            }
            if (__any_sync(kFull, want)) {
                if (want) {
                    if (FAST ? verify_own_prefix_aligned(R, own, cand, o, L) : verify_own_prefix(R, own, cand, o, L)) {
                        if (ns >= kSurv) {
                            hard = true;
                        } else {
#pragma unroll
                            for (int s = 0; s < kSurv; s++)
                                if (s == ns) s_id[s] = cand, s_o[s] = o, s_len[s] = lenb, s_t[s] = t;
                            ns++;
                        }
                    } else {
                        q_id[k * 32] = kNone;  // seed matched, overlap did not: not an arrival
                    }
                }
            }
        }

        // ---- in-neighbours from phase 1 (row of the transposed graph): kept unless a surviving arrival removes them
        uint32_t rowmask = 0;
        if (part && !hard) {
            const bool pairs = !csr && (rows.cap & 1u) == 0;  // fixed-capacity rows start on a sector: two entries per request
            for (uint32_t r0 = 0; r0 < deg; r0 += 2) {
                RevEntry en2[2];
                if (pairs) {
                    uint32_t q[8];
                    load8_na(reinterpret_cast<const uint32_t *>(row + r0), q);
                    en2[0].b = (int32_t) q[0], en2[0].o = (int32_t) q[1], en2[0].t = (uint64_t) q[2] | ((uint64_t) q[3] << 32);
                    en2[1].b = (int32_t) q[4], en2[1].o = (int32_t) q[5], en2[1].t = (uint64_t) q[6] | ((uint64_t) q[7] << 32);
                } else {
                    en2[0] = row[r0];
                    en2[1] = r0 + 1 < deg ? row[r0 + 1] : en2[0];
                }
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const uint32_t r = r0 + h;
                    if (r < deg) {
                        const RevEntry en = en2[h];
                        const uint32_t oa = (uint32_t) en.o;
                        const uint32_t lena = FAST ? lenc : R.len[(uint32_t) en.b];
                        bool removed = false;
#pragma unroll
                        for (int s = 0; s < kSurv; s++) {
                            if (s < ns && s_o[s] > 0 && oa >= s_o[s] &&
                                (FAST || (int64_t) s_len[s] + (int64_t) (oa - s_o[s]) - (int64_t) lena >= 0) &&
                                ((en.t ^ s_t[s]) >> (64u - 2u * s_o[s])) == 0)
                                removed = true;
                        }
                        if (!removed) rowmask |= 1u << r;
                    }
                }
            }
            // Same-id replacement rule (an arrival of read x also removes any older entry of x): whatever is about
            // to be emitted must be the only occurrence of its read among the arrivals and the row, else the
            // generic kernel decides.  Entries that are dropped anyway need no such check.
            for (uint32_t m = rowmask; m; m &= m - 1) {
                const uint32_t r = (uint32_t) __ffs(m) - 1u;
                const uint32_t a = (uint32_t) row[r].b;
                for (int k = 0; k < qn; k++)
                    if (q_id[k * 32] == a) hard = true;
                for (uint32_t r2 = 0; r2 < deg; r2++)
                    if (r2 != r && (uint32_t) row[r2].b == a) hard = true;
            }
            if (ns) {
                int cnt[kSurv];
#pragma unroll
                for (int s = 0; s < kSurv; s++) cnt[s] = 0;
                for (int k = 0; k < qn; k++) {
                    const uint32_t v = q_id[k * 32];
#pragma unroll
                    for (int s = 0; s < kSurv; s++) cnt[s] += (s < ns && s_id[s] == v) ? 1 : 0;
                }
#pragma unroll
                for (int s = 0; s < kSurv; s++)
                    if (cnt[s] > 1) hard = true;
            }
        }

        if (out.sh.world > 1) {
            // ---- emit, sharded: every surviving edge goes to the segment of the rank that owns its source read
            const bool emit = part && !hard;
#pragma unroll
            for (int sidx = 0; sidx < kSurv; sidx++) {
                const bool valid = emit && sidx < ns;
                if (__any_sync(kFull, valid)) {
                    const uint32_t d = valid ? shard_of(out.sh, s_id[sidx]) : 0u;
                    const uint32_t pos = shard_reserve(out.sh, valid, d, lane);
                    if (valid && pos < out.sh.cap) {
                        int32_t *t = reinterpret_cast<int32_t *>(out.sh.seg) + ((uint64_t) d * out.sh.cap + pos) * 3;
                        t[0] = (int32_t) s_id[sidx], t[1] = (int32_t) c, t[2] = (int32_t) s_o[sidx];
                    }
                }
            }
            uint32_t m = emit ? rowmask : 0u;
            while (__any_sync(kFull, m != 0u)) {
                const bool valid = m != 0u;
                RevEntry en;
                en.b = 0, en.o = 0, en.t = 0;
                if (valid) en = row[__ffs(m) - 1];
                m &= m - 1;
                const uint32_t d = valid ? shard_of(out.sh, (uint32_t) en.b) : 0u;
                const uint32_t pos = shard_reserve(out.sh, valid, d, lane);
                if (valid && pos < out.sh.cap) {
                    int32_t *t = reinterpret_cast<int32_t *>(out.sh.seg) + ((uint64_t) d * out.sh.cap + pos) * 3;
                    t[0] = en.b, t[1] = (int32_t) c, t[2] = en.o;
                }
            }
            if (hard) out.spill_queue[atomicAdd(out.n_spill, 1u)] = c;
            continue;
        }
        // ---- emit: one atomicAdd on the edge counter per warp
        const uint32_t n_out = (part && !hard) ? (uint32_t) ns + __popc(rowmask) : 0u;
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
                const RevEntry en = row[__ffs(m) - 1];
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

// the FAST instantiations: equal-length reads, none removed, no flags, read slots on 32-byte sector boundaries
// (alga_ps_plan_run makes such a copy; other callers fall back to the general instantiations)
inline bool fast_layout(const ReadsDev &R, const PsDev &P) {
    return P.uniform_len && !R.word_off && (R.stride & 7u) == 0 && ((uintptr_t) R.words & 31u) == 0;
}

inline int stride_words(int words) {
    return (words + 2) | 1;  // two pad words; odd stride: lanes of a warp fall into distinct banks
}

}  // namespace

// id_list / n_list: nullptr = first pass over [lo, hi) (two matches per window); else second pass over the queue of the
// first one (four matches per window; the queue length is read on the device, the grid is sized for a short queue)
void launch_phase1_tpr(const ReadsDev &R, const SeedTable &prefix, const PsDev &P, uint32_t max_len_nt, uint32_t lo,
                       uint32_t hi, const uint32_t *id_list, const uint32_t *n_list, const Phase1Out &out,
                       uint32_t *hard_queue, uint32_t *n_hard, int force_hard, cudaStream_t s, const LaunchCfg &cfg) {
    if (hi <= lo) return;
    int w = (int) ((max_len_nt + 15u) >> 4);
    if (w > kOwnWords) w = kOwnWords;
    const int wp = stride_words(w);
    const size_t smem = (size_t) kWarps * 32 * wp * sizeof(uint32_t);
    const bool fast = fast_layout(R, P);
    if (!id_list) {
        const int grid = warp_tile_grid(hi - lo, cfg, ALGA_P1_BLOCKS);
        auto k = !prefix.min_m ? (fast ? phase1_tpr_kernel<true, 2, 0> : phase1_tpr_kernel<false, 2, 0>)
                 : !cfg.min_slide ? (fast ? phase1_tpr_kernel<true, 2, 1> : phase1_tpr_kernel<false, 2, 1>)
                                  : (fast ? phase1_tpr_kernel<true, 2, 2> : phase1_tpr_kernel<false, 2, 2>);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, prefix, P, lo, hi, nullptr, nullptr, wp, out, hard_queue, n_hard, force_hard);
    } else {
        const int grid = warp_tile_grid(std::min<uint64_t>(hi - lo, (uint64_t) cfg.sm_count * kTpr * 2), cfg, 2);
        auto k = !prefix.min_m ? (fast ? phase1_tpr_kernel<true, 4, 0> : phase1_tpr_kernel<false, 4, 0>)
                 : !cfg.min_slide ? (fast ? phase1_tpr_kernel<true, 4, 1> : phase1_tpr_kernel<false, 4, 1>)
                                  : (fast ? phase1_tpr_kernel<true, 4, 2> : phase1_tpr_kernel<false, 4, 2>);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, prefix, P, lo, hi, id_list, n_list, wp, out, hard_queue, n_hard, force_hard);
    }
    bump(cfg);
}

void launch_phase2_tpr(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t max_len_nt, uint32_t lo,
                       uint32_t hi, const uint32_t *id_list, const uint32_t *n_list, const RowsView &rows,
                       const Phase2Out &out, int force_hard, cudaStream_t s, const LaunchCfg &cfg) {
    if (hi <= lo) return;
    int64_t lmax = P.max_l;
    if (lmax > (int64_t) max_len_nt) lmax = max_len_nt;
    if (lmax < 1) lmax = 1;
    int w = (int) ((2 * lmax + 31) >> 5);  // words of the longest prefix that takes part
    const int have = (int) ((max_len_nt + 15u) >> 4);
    if (w > have) w = have;
    const int wp = stride_words(w);
    const size_t smem = (size_t) (((kWarps * 32 * wp + 3) & ~3) + kWarps * (kQ2 * 32 * 3 + kQ2 * 32 / 2)) * sizeof(uint32_t);
    const bool fast = fast_layout(R, P);
    if (!id_list) {
        const int grid = warp_tile_grid(hi - lo, cfg, ALGA_P2_BLOCKS);
        auto k = !suffix.min_m ? (fast ? phase2_tpr_kernel<true, 2, 0> : phase2_tpr_kernel<false, 2, 0>)
                 : !cfg.min_slide ? (fast ? phase2_tpr_kernel<true, 2, 1> : phase2_tpr_kernel<false, 2, 1>)
                                  : (fast ? phase2_tpr_kernel<true, 2, 2> : phase2_tpr_kernel<false, 2, 2>);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, suffix, P, lo, hi, nullptr, nullptr, wp, rows, out, force_hard);
    } else {
        const int grid = warp_tile_grid(std::min<uint64_t>(hi - lo, (uint64_t) cfg.sm_count * kTpr * 2), cfg, 2);
        auto k = !suffix.min_m ? (fast ? phase2_tpr_kernel<true, 4, 0> : phase2_tpr_kernel<false, 4, 0>)
                 : !cfg.min_slide ? (fast ? phase2_tpr_kernel<true, 4, 1> : phase2_tpr_kernel<false, 4, 1>)
                                  : (fast ? phase2_tpr_kernel<true, 4, 2> : phase2_tpr_kernel<false, 4, 2>);
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, suffix, P, lo, hi, id_list, n_list, wp, rows, out, force_hard);
    }
    bump(cfg);
}

}  // namespace alga

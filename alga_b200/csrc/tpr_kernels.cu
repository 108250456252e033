// Thread-per-read overlap search kernels (sm_100a): the fast path of GraphCreatorPrefSuf phase 1 and phase 2.
//
// One THREAD owns one read, one WARP owns a tile of 32 consecutive reads (no block-level barriers).  The work of a
// read is split into two loops with very different instruction mixes:
//
//   probe    walk the overlap lengths.  The bucket of a K-nucleotide seed window is chosen by the MINIMIZER of the window
//            (common.cuh SeedTable), so consecutive lengths form RUNS that share one 128-byte bucket: about five runs for
//            the 29 lengths of phase 2, two or three for the dozen lengths phase 1 looks at.  Per round every lane works
//            on ONE run: its bucket was fetched kDepth rounds earlier by the whole warp -- eight cp.async instructions,
//            each covering four buckets with eight consecutive 16-byte lanes, i.e. one request per 128-byte line (a
//            random access costs one request whatever it carries: scripts/probes/random_coop.cu) -- into a staging area
//            in shared memory; the lane then tests the 16-bit tags of the bucket against the window hash of every length
//            of the run (SIMD halfword compares).  All lanes of a warp therefore wait for memory once per RUN, not once
//            per length: with one bucket per length the warp moved at the pace of "some lane has a load outstanding",
//            i.e. one DRAM latency per length (ncu at config-4 size: 61 us per tile in phase 1).  Tag hits are only
//            QUEUED; in phase 2 the first 64 bits of the hit read (all its overhang tail needs) follow by cp.async.
//   resolve  take the queued hits in order.  Exact 2-bit compares of whole overlaps are thread-local: every lane
//            verifies its own candidate -- read slots are sector aligned, a candidate arrives in one or two 32-byte
//            requests -- against its own read in shared memory.
//
// Two passes: the first over all reads allows two tag matches per window; reads that see more (same start position,
// different sequencing errors) are queued, and a queue of at least kSecondPassMin reads is run again with four
// matches per window (template parameter MAXM) before the rest goes to the generic kernels.
//
// History (profiles/): r01d interleaved probe and resolve per length (issue-bound); r01h split the
// loops; r01i fetched one 32-byte bucket per length through a cp.async ring (2.7 ms for config 2, but 148 ms for config 4:
// ncu showed 158 DRAM sectors per read in phase 2, most of them the same sector fetched again by 4- and 16-byte loads,
// and one exposed DRAM latency per length in phase 1); r2a cut the requests (sector-aligned read slots, 32-byte loads:
// 115 ms); this version probes by runs (99 ms), keeps its hot loops inside the SM's 32 KB instruction cache (find_runs rolled,
// one compare block per verification, rare paths out of line: 93 ms -- phase 2 had been bound by instruction fetch), and leaves
// index and CSR to sorted_stages.cu (85 ms).
//
// Reads the fast path cannot take at all (longer than 512 nt, a window with more than four tag matches, offsets above
// 32, more queued arrivals than fit, a source id that occurs twice for one target, ...) go to the generic kernels of
// prefsuf_kernels.cu, which replay GraphCreatorPrefSuf.cpp:356-488 literally.
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "launch.h"

namespace alga {

namespace {

constexpr int kTpr = 64;      // threads per block = 2 independent warps (no block-level barriers: small blocks only make the
constexpr int kWarps = kTpr / 32;  // shared-memory budget of an SM divisible)
// tuning knobs (scripts/build_variant.sh builds A/B variants with -D...)
#ifndef ALGA_P1_BLOCKS
#define ALGA_P1_BLOCKS 8
#endif
#ifndef ALGA_P2_BLOCKS
#define ALGA_P2_BLOCKS 7
#endif
#ifndef ALGA_Q2
#define ALGA_Q2 16
#endif
constexpr int kQ2 = ALGA_Q2;  // arrivals queued per target in phase 2
constexpr int kSurv = 4;      // surviving arrivals kept in registers per target
constexpr int kRowFast = 32;  // longest transposed row the fast phase-2 kernel takes
constexpr int kDepth = 2;     // runs (= buckets) in flight per lane
constexpr int kStageSlotWords = 8 * 32 * 4;             // one staging slot of a warp: [piece][lane] x 16 bytes = 32 buckets
constexpr int kStageWords = kDepth * kStageSlotWords;   // per warp
#ifndef ALGA_MAXRUNS
#define ALGA_MAXRUNS 16
#endif
constexpr int kMaxRuns = ALGA_MAXRUNS;   // runs per read the fast kernels keep (reads with more go to the generic kernels)
constexpr int kRunWords = kMaxRuns * 32 + kMaxRuns * 8;  // per warp: bucket [kMaxRuns][32] (u32) + first length index [kMaxRuns][32] (u8)

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
__device__ __forceinline__ void cp_async16(uint32_t *smem_dst, const uint32_t *gmem_src, uint64_t pol) {
    const uint32_t d = (uint32_t) __cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(d), "l"(gmem_src), "l"(pol) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ---- runs of overlap lengths that share a bucket --------------------------------------------------------------
// seed window of the own read (shared memory) for overlap length L.  UP (phase 1): the window starts at len - L, it moves
// up the read as L falls; else (phase 2): it ends at L, it moves down.
template <bool UP>
__device__ __forceinline__ uint64_t seed_window(const uint32_t *own, const PsDev &P, uint32_t len, int32_t L) {
    const uint32_t p = UP ? 2u * (len - (uint32_t) L) : 2u * (uint32_t) (L - P.seed_nt);
    return sbits64(own, p) & P.seed_mask;
}

// The walk over the overlap lengths l_hi, l_hi - 1, ... l_lo of one read, cut into runs of lengths whose seed windows have
// the same minimizer (= the same bucket): run k starts at length l_hi - first[k] and uses bucket bk[k].  All lanes step
// through their lengths together: the kNM scrambled m-mers of the current window sit in registers, oldest first; a step
// shifts them down by one, appends the m-mer that enters the window and takes the minimum of all -- no data-dependent
// rescan, so no lane waits for another one's.  ONE rolled loop does the first window and the steps: the version unrolled
// by kNM (m-mers in fixed slots, no shifting) was a third of the instructions either phase kernel keeps hot -- 10 KB of
// 35 KB -- and the kernels were bound by instruction fetch (32 KB instruction cache per SM; ncu r2f: 19 % of the fetches of
// phase 2 missed it and kept the GPC's instruction cache 93 % busy).  Returns the number of runs (may exceed kMaxRuns: the
// caller gives the read up); the table of the warp is indexed [k * 32 + lane].
template <bool UP>
__device__ __forceinline__ int find_runs(const uint32_t *own, const PsDev &P, const SeedTable &T, uint32_t len, int32_t l_hi,
                                         int32_t l_lo, bool on, uint32_t *run_bk, uint8_t *run_first, int lane) {
    const int n_len = on ? l_hi - l_lo + 1 : 0;
    const int n_max = warp_max(n_len);
    const uint32_t m = T.min_m;
    // m-mer i of the walk: the window of length l_hi holds i = 0 .. kNM - 1, step t brings in i = kNM - 1 + t and drops i = t - 1.
    // UP (phase 1): the window starts at len - l_hi and moves up; else (phase 2): it ends at l_hi and moves down
    const int32_t p0 = UP ? (int32_t) len - l_hi : l_hi - P.seed_nt + kNM - 1;
    uint32_t H[kNM];
#pragma unroll
    for (int j = 0; j < kNM; j++) H[j] = 0xFFFFFFFFu;
    uint32_t prev = 0;
    int n_runs = 0;
#pragma unroll 1
    for (int i = 0; i < kNM - 1 + n_max; i++) {
        const int t = i - (kNM - 1);
        const bool act = n_len > 0 && t < n_len;
        const int32_t pos = act ? (UP ? p0 + i : p0 - i) : 0;
        const uint32_t hm = scrambled_mmer(sbits64(own, 2u * (uint32_t) pos), 0, m);
#pragma unroll
        for (int j = 0; j + 1 < kNM; j++) H[j] = H[j + 1];
        H[kNM - 1] = hm;
        if (t >= 0) {  // the same for all lanes
            uint32_t v[kNM];  // a tree, not a chain: the minimum is on the critical path of every step
#pragma unroll
            for (int j = 0; j < kNM; j++) v[j] = H[j];
#pragma unroll
            for (int w = 1; w < kNM; w <<= 1)
#pragma unroll
                for (int j = 0; j + w < kNM; j += 2 * w) v[j] = min(v[j], v[j + w]);
            if (act && (t == 0 || v[0] != prev)) {
                prev = v[0];
                if (n_runs < kMaxRuns) {
                    run_bk[n_runs * 32 + lane] = bucket_of(mix64((uint64_t) v[0]), T.n_buckets);
                    run_first[n_runs * 32 + lane] = (uint8_t) t;
                }
                n_runs++;
            }
        }
    }
    return n_runs;
}

// Warp-collective fetch: every lane with `need` gets its 128-byte bucket `bk` into its place of the staging slot.  Eight
// instructions; in each, the 8 lanes 8q .. 8q+7 cover the 8 x 16 bytes of the bucket of lane 4 i + q: one request per line
// (the ids must come with the tags: a line of which only the first 48 bytes were asked for arrives without its upper
// sectors, and the ids of the matching entries then cost a DRAM round trip each -- measured, r2d).
__device__ __forceinline__ void fetch_buckets(const SeedTable &T, uint32_t *slot, bool need, uint32_t bk, int lane, uint64_t pol) {
    const unsigned needm = __ballot_sync(kFull, need);
    if (needm) {
        const int piece = lane & 7;
#pragma unroll
        for (int i = 0; i < 8; i++) {
            const int owner = 4 * i + (lane >> 3);
            const uint32_t obk = __shfl_sync(kFull, bk, owner);
            if ((needm >> owner) & 1u)
                cp_async16(slot + ((piece * 32 + owner) << 2), T.slots + (uint64_t) obk * kBucketWords + piece * 4, pol);
        }
    }
    cp_async_commit();
}
// this lane's piece p (16 bytes) of a staged bucket
__device__ __forceinline__ uint4 staged_piece(const uint32_t *slot, int p, int lane) {
    return *reinterpret_cast<const uint4 *>(slot + ((p * 32 + lane) << 2));
}
// id of entry i of the staged bucket of lane `lane`
__device__ __forceinline__ uint32_t staged_id(const uint32_t *slot, int i, int lane) {
    return slot[(((3 + (i >> 2)) * 32 + lane) << 2) + (i & 3)];
}
struct BucketTags {
    uint32_t cnt;     // inserts that chose the bucket (> kBucketCap: the chain goes on in the next bucket)
    uint32_t tw[10];  // two 16-bit tags per word
};
__device__ __forceinline__ void load_tags(const uint32_t *slot, int lane, BucketTags &b) {
    const uint4 q0 = staged_piece(slot, 0, lane), q1 = staged_piece(slot, 1, lane), q2 = staged_piece(slot, 2, lane);
    b.cnt = q0.x;
    b.tw[0] = q0.z, b.tw[1] = q0.w, b.tw[2] = q1.x, b.tw[3] = q1.y, b.tw[4] = q1.z, b.tw[5] = q1.w;
    b.tw[6] = q2.x, b.tw[7] = q2.y, b.tw[8] = q2.z, b.tw[9] = q2.w;
}

__device__ __forceinline__ void order_desc(uint32_t &a, uint32_t &b) {
    const uint32_t hi = max(a, b), lo = min(a, b);
    a = hi, b = lo;
}

// Entries of the staged bucket whose tag equals that of a window with hash h: bit k = entry 2 k, bit 16 + k = entry 2 k + 1.
// Only the 16-byte pieces that hold entries are looked at (a bucket holds 3 to 6 on average: entries 0-3 sit in the first piece
// next to the count, 4-11 in the second, 12-19 in the third); the tag compare was 9 % / 12 % of the instructions of phase 2 / 1.
__device__ __forceinline__ uint32_t match_words(const uint32_t *tw, int k0, int k1, __half2 t2) {
    uint32_t acc = 0;
#pragma unroll
    for (int k = k0; k < k1; k++) {
        const uint32_t w = tw[k - k0];
        acc += (__heq2_mask(*reinterpret_cast<const __half2 *>(&w), t2) & 0x00010001u) << k;  // distinct bits: + is |
    }
    return acc;
}
__device__ __forceinline__ __half2 tag_pair(uint64_t h) {
    const uint32_t tt = tag_of(h) * 0x10001u;
    return *reinterpret_cast<const __half2 *>(&tt);
}
__device__ __forceinline__ uint32_t match_mask(const BucketTags &b, uint64_t h) {  // tags in registers (phase 1: many lengths per bucket)
    const __half2 t2 = tag_pair(h);
    uint32_t acc = match_words(b.tw, 0, 2, t2);
    if (b.cnt > 4u) acc += match_words(b.tw + 2, 2, 6, t2);
    if (b.cnt > 12u) acc += match_words(b.tw + 6, 6, 10, t2);
    return acc;
}
// the same straight from the staging slot (phase 2: one length per item); also returns the count of the bucket
__device__ __forceinline__ uint32_t match_mask_staged(const uint32_t *slot, int lane, uint64_t h, uint32_t &cnt) {
    const __half2 t2 = tag_pair(h);
    const uint4 q0 = staged_piece(slot, 0, lane);
    cnt = q0.x;
    const uint32_t t0[2] = {q0.z, q0.w};
    uint32_t acc = match_words(t0, 0, 2, t2);
    if (cnt > 4u) {
        const uint4 q1 = staged_piece(slot, 1, lane);
        const uint32_t t1[4] = {q1.x, q1.y, q1.z, q1.w};
        acc += match_words(t1, 2, 6, t2);
    }
    if (cnt > 12u) {
        const uint4 q2 = staged_piece(slot, 2, lane);
        const uint32_t t2w[4] = {q2.x, q2.y, q2.z, q2.w};
        acc += match_words(t2w, 6, 10, t2);
    }
    return acc;
}
// Rare path, kept out of line (inlined, the twenty compare-and-call sites of probe_seed_at sit in the middle of the hottest loop
// of either kernel): the reads in the bucket chain from bk on whose tag matches that of h -> ids[0 .. min(n, cap)); returns n.
__device__ __noinline__ int chain_matches(SeedTable T, uint64_t h, uint32_t bk, uint32_t *ids, int cap) {
    int n = 0;
    probe_seed_at(T, h, bk, [&](uint32_t id) {
        if (n < cap) ids[n] = id;
        n++;
    });
    return n;
}

// The reads behind a match mask -- and behind the chain of the bucket, if it overflowed -- : ids in c[0 .. min(n, MAXM)),
// largest first (the order of arrival inside one overlap length, seen backwards).  n > MAXM: the caller gives up.
// (MAXM = reads that share a K-nucleotide seed and overlap at one length, i.e. start at the same position: 2 in the first
// pass over all reads, 4 in the second pass over the reads the first one gave up on.)
template <int MAXM>
__device__ __forceinline__ void collect_matches(const SeedTable &T, const BucketTags &b, const uint32_t *slot, int lane, uint32_t bk,
                                                uint64_t h, uint32_t mask, uint32_t (&c)[MAXM], int &n) {
    n = 0;
#pragma unroll
    for (int k = 0; k < MAXM; k++) c[k] = 0u;
    auto add = [&](uint32_t id) {
#pragma unroll
        for (int k = 0; k < MAXM; k++)
            if (k == n) c[k] = id;
        n++;
    };
    while (mask) {
        const int bit = __ffs(mask) - 1;
        mask &= mask - 1;
        add(staged_id(slot, ((bit & 15) << 1) | (bit >> 4), lane));
    }
    if (b.cnt > (uint32_t) kBucketCap) {  // rare: plain loads
        uint32_t ids[MAXM + 1];
        const int m = chain_matches(T, h, next_bucket(T, bk), ids, MAXM + 1);
        for (int i = 0; i < m && i < MAXM + 1; i++) add(ids[i]);
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
    // one sector per round, the next one requested before this one is compared (one compare block in the code: the kernels
    // are short of instruction cache: find_runs)
    uint32_t g[8], gn[8];
    load8_na(pb, g);
#pragma unroll 1
    for (int32_t s0 = 0; s0 < n_words; s0 += 8) {
        const bool more = s0 + 8 < n_words;
        if (more) load8_na(pb + s0 + 8, gn);
        uint32_t diff = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const int32_t p = 32 * (s0 + j) - sh2;  // own-read bit that meets bit 0 of this candidate word
            if (p + 32 > 0 && p < nb) {
                uint32_t cw, mask = 0xFFFFFFFFu;
                if (p >= 0) {
                    const uint32_t w = (uint32_t) p >> 5;
                    cw = __funnelshift_r(own[w], own[w + 1], (uint32_t) p & 31u);
                } else {
                    cw = own[0] << (uint32_t) (-p);
                    mask <<= (uint32_t) (-p);
                }
                if (nb - p < 32) mask &= (1u << (uint32_t) (nb - p)) - 1u;
                diff |= (g[j] ^ cw) & mask;
            }
        }
        if (diff) return false;
        if (more) {
#pragma unroll
            for (int j = 0; j < 8; j++) g[j] = gn[j];
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------
// Phase 1 (GraphCreatorPrefSuf.cpp:397-402 in closed form): source read b walks L from min(rs-1, len) downwards
// and keeps the first 3 confirmed (L, c) -- within one L the larger c first -- = "the last 3 pushes".
//
// A lane stops walking once it holds 3 candidates (confirmed + pending) -- in the middle of a run if need be, the run
// then stays in its staging slot -- and goes on only if the exact compares reject some of them.  The lanes of a warp
// therefore stand at different runs; `head` says which of its two staging slots holds the older run of a lane.
// Shared memory per warp: own reads [32][wp] | staging [kDepth][8][32] x 16 B | run table.
// id_list != nullptr: second pass -- the reads are id_list[0 .. *n_list) instead of [lo, hi)
template <bool FAST, int MAXM>
__global__ void __launch_bounds__(kTpr, ALGA_P1_BLOCKS)
phase1_tpr_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, const uint32_t *__restrict__ id_list,
                  const uint32_t *__restrict__ n_list, int wp, Phase1Out out, uint32_t *__restrict__ hard_queue,
                  uint32_t *n_hard, int force_hard) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int own_words = ((kWarps * 32 * wp + 3) & ~3);
    uint32_t *wown = smem + wib * 32 * wp;
    const uint32_t *own = wown + lane * wp;
    uint32_t *stage = smem + own_words + wib * kStageWords;
    uint32_t *run_bk = smem + own_words + kWarps * kStageWords + wib * kRunWords;
    uint8_t *run_first = reinterpret_cast<uint8_t *>(run_bk + kMaxRuns * 32);
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

        // the walk of every lane, cut into runs (shared-memory table of the warp)
        int n_runs = find_runs<true>(own, P, T, lenb, l_hi, P.lmin, active, run_bk, run_first, lane);
        if (n_runs > kMaxRuns || (active && l_hi - P.lmin > 255)) {
            hard = true;
            n_runs = 0;
        }
        __syncwarp();
        // run k lives in staging slot k & 1.  kc: the run this lane consumes next (from length Lc on), kf: the run it fetches
        // next; kc <= kf <= kc + 2.  Round j (the same for all lanes) serves the runs of parity j.
        int kc = 0, kf = 0;
        int32_t Lc = l_hi;
        int j = 0;
        while (true) {
            // ---- probe, at most one run per lane and round, until every lane has 3 candidates (confirmed + pending) or ran out
            while (true) {
                const bool need = active && !hard && conf + np < kSmallEdgesKept && kc < n_runs;
                if (!__any_sync(kFull, need)) break;
                cp_async_wait_group<kDepth - 1>();  // every fetch but the latest has landed (this lane's part of it)
                __syncwarp();                       // ... and everybody else's
                uint32_t *slot = stage + j * kStageSlotWords;
                if (need && (kc & 1) == j && kf > kc) {
                    BucketTags bt;
                    load_tags(slot, lane, bt);
                    const uint32_t bk = run_bk[kc * 32 + lane];
                    const int32_t l_last = kc + 1 < n_runs ? l_hi - (int32_t) run_first[(kc + 1) * 32 + lane] + 1 : P.lmin;
                    int32_t L = Lc;
                    for (; L >= l_last && conf + np < kSmallEdgesKept && !hard; L--) {
                        const uint64_t h = mix64(seed_window<true>(own, P, lenb, L));
                        const uint32_t mask = match_mask(bt, h);
                        if (mask || bt.cnt > (uint32_t) kBucketCap) {
                            uint32_t cm[MAXM];
                            int n = 0;
                            collect_matches<MAXM>(T, bt, slot, lane, bk, h, mask, cm, n);
                            if (n > MAXM) {
                                hard = true;
                            } else {  // within one length the larger target id is the later push: cm[] is descending
#pragma unroll
                                for (int q = 0; q < MAXM; q++) {
                                    if (q < n) {
#pragma unroll
                                        for (int k = 0; k < kPend; k++)
                                            if (k == np) pc[k] = cm[q], pl[k] = L;
                                        np++;
                                    }
                                }
                            }
                        }
                    }
                    Lc = L;  // enough candidates for now: the rest of the run waits in its slot
                    if (L < l_last) kc++;  // the run is used up
                }
                __syncwarp();  // nobody reads this slot any more
                // fetch ahead: the next run of parity j, if its slot is free and the lane still looks for candidates
                const bool fetch = active && !hard && kf < n_runs && (kf & 1) == j && kf < kc + 2 && conf + np < kSmallEdgesKept;
                fetch_buckets(T, slot, fetch, fetch ? run_bk[kf * 32 + lane] : 0u, lane, pol);
                if (fetch) kf++;
                j ^= 1;
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

        cp_async_wait_all();  // buckets fetched ahead and never used must have landed before the next tile reuses the slots

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
                            // one 16-byte store = one request (the member-wise copy was two 8-byte stores: ncu r2f, 340 M requests)
                            *reinterpret_cast<uint4 *>(out.rows + (uint64_t) (c - out.c_base) * out.row_cap + pos[k]) =
                                make_uint4((uint32_t) r.b, (uint32_t) r.o, (uint32_t) r.t, (uint32_t) (r.t >> 32));
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
// Shared memory per warp: own reads [32][wp] | staging [kDepth][8][32] x 16 B | queue: ids [kQ2][32], heads [kQ2][32]
// (u64), lengths [kQ2][32] (u16).  Every lane walks all its runs, run k of every lane sits in staging slot k & 1.
// id_list != nullptr: second pass -- the targets are id_list[0 .. *n_list) instead of [lo, hi)
template <bool FAST, int MAXM>
__global__ void __launch_bounds__(kTpr, ALGA_P2_BLOCKS)
phase2_tpr_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, const uint32_t *__restrict__ id_list,
                  const uint32_t *__restrict__ n_list, int wp, RowsView rows, Phase2Out out, int force_hard) {
    extern __shared__ __align__(16) uint32_t smem[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    const int own_words = ((kWarps * 32 * wp + 3) & ~3);
    constexpr int kQueueWords = kQ2 * 32 + kQ2 * 32 / 2;  // ids + lengths (u16); the heads (u64) reuse the staging area
    uint32_t *wown = smem + wib * 32 * wp;
    const uint32_t *own = wown + lane * wp;
    uint32_t *stage = smem + own_words + wib * kStageWords;
    uint32_t *run_bk = smem + own_words + kWarps * kStageWords + wib * kRunWords;
    uint8_t *run_first = reinterpret_cast<uint8_t *>(run_bk + kMaxRuns * 32);
    uint32_t *ctl = smem + own_words + kWarps * (kStageWords + kRunWords) + wib * 72;  // queue lengths [32] | overflow flags [32] | lanes [32] (u8)
    uint32_t *q_id = smem + own_words + kWarps * (kStageWords + kRunWords + 72) + wib * kQueueWords + lane;  // q_id[k * 32]
    uint16_t *q_l = reinterpret_cast<uint16_t *>(q_id - lane + kQ2 * 32) + lane;                 // q_l[k * 32]
    uint64_t *q_t = reinterpret_cast<uint64_t *>(stage) + lane;                                  // q_t[k * 32]: first 64 bits of the hit read
    static_assert(kQ2 * 32 * 2 <= kStageWords, "the heads of the queued reads must fit the staging area");
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

        // ---- probe: queue every tag hit (b, L); FAST: the first 64 bits of b follow by cp.async.
        // The lengths of run k of all 32 reads of the tile form one flat list of (read, length) items that the lanes of
        // the warp share out among themselves -- a lane works on whatever item comes next, not on "its" read (the reads
        // and the staged buckets sit in shared memory, any lane reaches them) -- so every lane is busy in every step
        // although the runs of the reads differ in length (one read per lane kept 11 of 32 lanes busy: ncu, r2c).
        int qn = 0;
        const bool walk = active && !hard;
        int n_runs = find_runs<false>(own, P, T, lenc, l_hi, l_lo, walk, run_bk, run_first, lane);
        if (n_runs > kMaxRuns || (walk && l_hi - l_lo > 255)) {
            hard = true;
            n_runs = 0;
        }
        ctl[lane] = 0u;       // arrivals queued for the read of this lane (by any lane)
        ctl[32 + lane] = 0u;  // ... more than the queue holds
        __syncwarp();
        const bool on = walk && !hard;
        uint32_t *q_id0 = q_id - lane;
        uint16_t *q_l0 = q_l - lane;
        // run k lives in staging slot k & 1; round t fetches run t and works through run t - 2 of every read
        for (int t = 0;; t++) {
            const int j = t & 1, kc = t - 2;
            if (!__any_sync(kFull, on && kc < n_runs)) break;
            cp_async_wait_group<kDepth - 1>();  // every fetch but the latest has landed (this lane's part of it)
            __syncwarp();                       // ... and everybody else's
            uint32_t *slot = stage + j * kStageSlotWords;
            int n_mine = 0;        // lengths of run kc of the read of this lane
            int32_t lf_mine = 0;   // ... the first (longest) of them
            if (on && kc >= 0 && kc < n_runs) {
                const int f0 = run_first[kc * 32 + lane];
                const int f1 = kc + 1 < n_runs ? (int) run_first[(kc + 1) * 32 + lane] : l_hi - l_lo + 1;
                n_mine = f1 - f0;
                lf_mine = l_hi - f0;
            }
            int incl = n_mine;
            for (int d = 1; d < 32; d <<= 1) {
                const int y = __shfl_up_sync(kFull, incl, d);
                if (lane >= d) incl += y;
            }
            const int total = __shfl_sync(kFull, incl, 31), off_mine = incl - n_mine;
            // the reads that take part in this round, in lane order: lane i holds the lane of the i-th of them
            const unsigned activem = __ballot_sync(kFull, n_mine > 0);
            uint8_t *nth = reinterpret_cast<uint8_t *>(ctl + 64);
            if (n_mine > 0) nth[__popc(activem & ((1u << lane) - 1u))] = (uint8_t) lane;
            __syncwarp();
            for (int base = 0; base < total; base += 32) {
                const int item = base + lane;
                // the read of this item = the last participating read whose first item is not beyond it: reads that start
                // inside this window of 32 items mark their place, the others are counted
                const bool inwin = n_mine > 0 && off_mine >= base && off_mine < base + 32;
                const unsigned marks = __reduce_or_sync(kFull, inwin ? 1u << (off_mine - base) : 0u);
                const int before = __popc(__ballot_sync(kFull, n_mine > 0 && off_mine < base));
                const int rank = before + __popc(marks & (0xFFFFFFFFu >> (31 - lane))) - 1;
                const int r = nth[rank & 31];
                const int off_r = __shfl_sync(kFull, off_mine, r);
                const int32_t lf_r = __shfl_sync(kFull, lf_mine, r);
                const uint32_t c_r = __shfl_sync(kFull, c, r);
                if (item < total) {
                    const int32_t L = lf_r - (item - off_r);
                    const uint64_t h = mix64(seed_window<false>(wown + r * wp, P, 0u, L));
                    auto push = [&](uint32_t cand) {
                        if (cand == c_r) return;
                        const uint32_t pos = atomicAdd(ctl + r, 1u);
                        if (pos < (uint32_t) kQ2) {
                            q_id0[pos * 32 + r] = cand;
                            q_l0[pos * 32 + r] = (uint16_t) L;
                        } else {
                            ctl[32 + r] = 1u;
                        }
                    };
                    uint32_t bt_cnt;
                    uint32_t mask = match_mask_staged(slot, r, h, bt_cnt);
                    while (mask) {
                        const int bit = __ffs(mask) - 1;
                        mask &= mask - 1;
                        push(staged_id(slot, ((bit & 15) << 1) | (bit >> 4), r));
                    }
                    if (bt_cnt > (uint32_t) kBucketCap) {  // rare
                        uint32_t ids[8];
                        const int m = chain_matches(T, h, next_bucket(T, run_bk[kc * 32 + r]), ids, 8);
                        if (m > 8) ctl[32 + r] = 1u;  // the generic kernel takes the read
                        for (int i = 0; i < m && i < 8; i++) push(ids[i]);
                    }
                }
            }
            __syncwarp();  // nobody reads this slot any more
            const bool fetch = on && t < n_runs;
            fetch_buckets(T, slot, fetch, fetch ? run_bk[t * 32 + lane] : 0u, lane, pol);
        }
        cp_async_wait_all();
        __syncwarp();
        qn = (int) min(ctl[lane], (uint32_t) kQ2);
        if (ctl[32 + lane]) hard = true;
        // The items of a step were queued in no particular order: bring the queue of every read into the order of arrival
        // seen backwards -- longer overlap first, within one length the larger source id first (insertion sort: the queue
        // is almost in order, only the arrivals of one step are mixed).
        {
            const int q_top = warp_max(hard ? 0 : qn);
            for (int i = 1; i < q_top; i++) {
                if (i < qn && !hard) {
                    const uint32_t id_i = q_id[i * 32];
                    const uint16_t l_i = q_l[i * 32];
                    int k = i - 1;
                    while (k >= 0 && (q_l[k * 32] < l_i || (q_l[k * 32] == l_i && q_id[k * 32] < id_i))) {
                        q_id[(k + 1) * 32] = q_id[k * 32];
                        q_l[(k + 1) * 32] = q_l[k * 32];
                        k--;
                    }
                    q_id[(k + 1) * 32] = id_i;
                    q_l[(k + 1) * 32] = l_i;
                }
            }
            // FAST: the first 64 bits of every queued read (all its overhang tail needs), one 8-byte request each, all in
            // flight together; they land in the staging area, which the probe no longer needs
            if (FAST) {
#pragma unroll 1
                for (int k = 0; k < q_top; k++)
                    if (k < qn && !hard) cp_async8(q_t + k * 32, R.words + (uint64_t) q_id[k * 32] * R.stride);
                cp_async_wait_all();
            }
        }

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
#pragma unroll 1
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
    const size_t smem = (size_t) (((kWarps * 32 * wp + 3) & ~3) + kWarps * (kStageWords + kRunWords)) * sizeof(uint32_t);
    const bool fast = fast_layout(R, P);
    if (!id_list) {
        const int grid = warp_tile_grid(hi - lo, cfg, ALGA_P1_BLOCKS);
        auto k = fast ? phase1_tpr_kernel<true, 2> : phase1_tpr_kernel<false, 2>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, prefix, P, lo, hi, nullptr, nullptr, wp, out, hard_queue, n_hard, force_hard);
    } else {
        const int grid = warp_tile_grid(std::min<uint64_t>(hi - lo, (uint64_t) cfg.sm_count * kTpr * 2), cfg, 2);
        auto k = fast ? phase1_tpr_kernel<true, 4> : phase1_tpr_kernel<false, 4>;
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
    const size_t smem = (size_t) (((kWarps * 32 * wp + 3) & ~3) + kWarps * (kStageWords + kRunWords + 72) +
                                  kWarps * (kQ2 * 32 + kQ2 * 32 / 2)) * sizeof(uint32_t);
    const bool fast = fast_layout(R, P);
    if (!id_list) {
        const int grid = warp_tile_grid(hi - lo, cfg, ALGA_P2_BLOCKS);
        auto k = fast ? phase2_tpr_kernel<true, 2> : phase2_tpr_kernel<false, 2>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, suffix, P, lo, hi, nullptr, nullptr, wp, rows, out, force_hard);
    } else {
        const int grid = warp_tile_grid(std::min<uint64_t>(hi - lo, (uint64_t) cfg.sm_count * kTpr * 2), cfg, 2);
        auto k = fast ? phase2_tpr_kernel<true, 4> : phase2_tpr_kernel<false, 4>;
        cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
        k<<<grid, kTpr, smem, s>>>(R, suffix, P, lo, hi, id_list, n_list, wp, rows, out, force_hard);
    }
    bump(cfg);
}

}  // namespace alga

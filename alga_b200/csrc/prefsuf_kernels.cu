// Overlap search + transitive reduction kernels of the GraphCreatorPrefSuf hot path (sm_100a).
//
// The reference rebuilds a hash table of all length-L prefix fingerprints for every overlap length L
// and probes it with every length-L suffix fingerprint (GraphCreatorPrefSuf.cpp:238-315, 356-488).
// Here each read is indexed ONCE by a seed (its first K nucleotides for the prefix side, its last K
// for the suffix side, K = min(min_overlap, 32)); every (read, L) pair then costs one 32-byte bucket
// probe plus an exact 2-bit compare of the full overlap on a tag hit.  Acceptance by exact compare
// equals the reference's acceptance by (mod 10^18+3, mod 10^9+7) fingerprint equality
// (GraphCreatorPrefSuf.cpp:385-387) up to fingerprint collisions (~1e-27 per pair).
//
// Work decomposition: one warp per read, the 32 lanes probe 32 overlap lengths at once.
//   phase 1 (L <  rs): warp = suffix read b, probes the prefix index, keeps the 3 largest (L, c)
//                      = "the last 3 pushes" of GraphCreatorPrefSuf.cpp:397-402 in canonical order.
//   phase 2 (L >= rs): warp = prefix read c, probes the suffix index for L ascending and replays the
//                      per-target reduction of GraphCreatorPrefSuf.cpp:403-483 in (L, b) order on an
//                      in-neighbour list kept in shared memory (global memory for spilled targets).
#include "launch.h"

namespace alga {

namespace {

constexpr int kWarpsPerBlock = 8;
constexpr int kThreads = kWarpsPerBlock * 32;

inline int grid_for(uint64_t n_items, int per_block, const LaunchCfg &cfg, int max_blocks_per_sm = 16) {
    uint64_t need = (n_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t) cfg.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}
inline void bump(const LaunchCfg &cfg) {
    if (cfg.launches) (*cfg.launches)++;
}

// ------------------------------------------------------------------------------------------------
__global__ void read_stats_kernel(ReadsDev R, int lmin, int min_offset, ReadStats *stats) {
    uint32_t mx = 0, mn = 0xFFFFFFFFu, np = 0, ns = 0;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < R.n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t len = R.len[i];
        mx = max(mx, len);
        mn = min(mn, len);
        if (len && flag_to(R, i) && (int64_t) len >= lmin) np++;
        if (len && flag_from(R, i) && (int64_t) len - min_offset >= lmin) ns++;
    }
    for (int d = 16; d; d >>= 1) {
        mx = max(mx, __shfl_xor_sync(kFull, mx, d));
        mn = min(mn, __shfl_xor_sync(kFull, mn, d));
        np += __shfl_xor_sync(kFull, np, d);
        ns += __shfl_xor_sync(kFull, ns, d);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&stats->max_len, mx);
        atomicMin(&stats->min_len, mn);
        atomicAdd(&stats->n_prefix, np);
        atomicAdd(&stats->n_suffix, ns);
    }
}

// ------------------------------------------------------------------------------------------------
// Seed index build: one thread per read, two inserts (prefix side, suffix side).
// [b_lo, b_hi): only seeds whose bucket falls in this range are inserted (sharded build: the rank's own slice)
// which: bit 0 = prefix table, bit 1 = suffix table
__global__ void build_index_kernel(ReadsDev R, PsDev P, SeedTable tp, SeedTable ts, uint32_t lo, uint32_t hi, uint32_t b_lo,
                                   uint32_t b_hi, int which) {
    for (uint64_t i = (uint64_t) lo + blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < hi; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t len = P.uniform_len ? P.uniform_len : R.len[i];
        if (len == 0 || (int64_t) len < P.lmin) continue;
        const uint32_t *p = read_ptr(R, (uint32_t) i);
        if ((which & 1) && flag_to(R, (uint32_t) i)) {
            const uint64_t win = bits64(p, 0) & P.seed_mask, h = mix64(win);
            const uint32_t bk = bucket_index_rt(tp, win, h, (uint32_t) P.seed_nt);
            if (bk >= b_lo && bk < b_hi) insert_seed_at(tp, h, bk, (uint32_t) i);
        }
        if ((which & 2) && flag_from(R, (uint32_t) i) && (int64_t) len - P.min_offset >= P.lmin) {
            const uint64_t win = bits64(p, 2u * (len - (uint32_t) P.seed_nt)) & P.seed_mask, h = mix64(win);
            const uint32_t bk = bucket_index_rt(ts, win, h, (uint32_t) P.seed_nt);
            if (bk >= b_lo && bk < b_hi) insert_seed_at(ts, h, bk, (uint32_t) i);
        }
    }
}

// Sharded build, equal-length reads: the two seeds of a read -- bucket and tag on either side -- are computed ONCE, by the
// rank that owns the read (from its shard in the caller's layout), and travel with the shard; every rank then inserts
// out of these 12-byte records the seeds that fall into its slice of the bucket space.  (Each rank scanning all reads
// for its slice cost 5.4 ms of a 21 ms build on 8 GPUs: the minimizers of 2 x 57 M windows, eight times over.)
__global__ void seed_keys_kernel(const uint32_t *__restrict__ words, uint32_t stride, uint32_t n, PsDev P, SeedTable tp,
                                 SeedTable ts, uint32_t *__restrict__ keys) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t *p = words + i * stride;
        const uint32_t len = P.uniform_len;
        const uint64_t wp = bits64(p, 0) & P.seed_mask, hp = mix64(wp);
        const uint64_t ws = bits64(p, 2u * (len - (uint32_t) P.seed_nt)) & P.seed_mask, hs = mix64(ws);
        keys[3 * i] = bucket_index_rt(tp, wp, hp, (uint32_t) P.seed_nt);
        keys[3 * i + 1] = bucket_index_rt(ts, ws, hs, (uint32_t) P.seed_nt);
        keys[3 * i + 2] = tag_of(hp) | (tag_of(hs) << 16);
    }
}
__global__ void index_keys_kernel(const uint32_t *__restrict__ keys, uint32_t first_id, uint32_t n, SeedTable tp, SeedTable ts,
                                  uint32_t b_lo, uint32_t b_hi) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t bp = keys[3 * i], bs = keys[3 * i + 1], tg = keys[3 * i + 2];
        if (bp >= b_lo && bp < b_hi) insert_tag_at(tp, tg & 0xFFFFu, bp, first_id + (uint32_t) i);
        if (bs >= b_lo && bs < b_hi) insert_tag_at(ts, tg >> 16, bs, first_id + (uint32_t) i);
    }
}

// ------------------------------------------------------------------------------------------------
// Sharded runs: what the other ranks produced for this rank's reads is read straight out of their exchange
// workspaces over NVLink (peer pointers), no staging copy.
struct PeerSegs {
    const void *seg[8];      // peer p's segment for this rank
    const uint32_t *cnt[8];  // peer p's entry count for this rank
    uint32_t cap;
    int world;
};

// Peer memory is not cached on this side of the link, so every load instruction costs its own NVLink sectors: the
// pull kernels move tiles of the peer's segment into shared memory with fully coalesced 16-byte loads and work from
// there.
constexpr int kPullTile = 1024;  // entries per block and round

__device__ __forceinline__ void pull_tile(const void *src, uint64_t first, uint32_t count, uint32_t entry_bytes, uint4 *sh) {
    // segments start 16-byte aligned and entry_bytes * kPullTile is a multiple of 16
    const uint4 *g = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(src) + first * entry_bytes);
    const uint32_t n16 = (count * entry_bytes + 15u) >> 4;
    for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) sh[i] = g[i];
    __syncthreads();
}

// phase-1 edges whose target this rank owns -> rows of the transposed graph (blockIdx.y = peer).  Four entries per
// thread and round: their four atomics are in flight together.
__global__ void __launch_bounds__(256) pull_rows_kernel(PeerSegs ps, Phase1Out out) {
    __shared__ uint4 sh[kPullTile * sizeof(Edge1) / 16];
    const int p = blockIdx.y;
    uint32_t n = *ps.cnt[p];
    if (n > ps.cap) n = ps.cap;
    const Edge1 *tile = reinterpret_cast<const Edge1 *>(sh);
    for (uint64_t first = (uint64_t) blockIdx.x * kPullTile; first < n; first += (uint64_t) gridDim.x * kPullTile) {
        const uint32_t count = (uint32_t) min((uint64_t) kPullTile, (uint64_t) n - first);
        __syncthreads();
        pull_tile(ps.seg[p], first, count, sizeof(Edge1), sh);
        Edge1 e[4];
        uint32_t pos[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t i = threadIdx.x + j * 256;
            if (i < count) {
                e[j] = tile[i];
                pos[j] = atomicAdd(out.indeg + ((uint32_t) e[j].c - out.c_base), 1u);
            }
        }
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint32_t i = threadIdx.x + j * 256;
            if (i < count) {
                const uint32_t ci = (uint32_t) e[j].c - out.c_base;
                if (pos[j] < out.row_cap) {
                    // one 16-byte store = one request (a member-wise copy of the entry is two)
                    *reinterpret_cast<uint4 *>(out.rows + (uint64_t) ci * out.row_cap + pos[j]) =
                        make_uint4((uint32_t) e[j].b, (uint32_t) e[j].o, (uint32_t) e[j].t, (uint32_t) (e[j].t >> 32));
                } else {
                    const uint32_t k = atomicAdd(out.n_list, 1u);
                    if (k < out.list_cap) {
                        Edge1 x = e[j];
                        x.c = (int32_t) ci;
                        out.list[k] = x;
                    }
                }
            }
        }
    }
}

// surviving edges whose source this rank owns -> local triples + out-degrees (blockIdx.y = peer)
__global__ void __launch_bounds__(256) pull_triples_kernel(PeerSegs ps, uint64_t cap_local, int32_t *__restrict__ triples,
                                                           unsigned long long *n_total, uint32_t *outdeg) {
    __shared__ uint4 sh[kPullTile * 12 / 16];
    const int p = blockIdx.y;
    uint64_t base = 0, total = 0;
    uint32_t n = 0;
    for (int q = 0; q < ps.world; q++) {
        uint32_t c = *ps.cnt[q];
        if (c > ps.cap) c = ps.cap;
        if (q < p) base += c;
        if (q == p) n = c;
        total += c;
    }
    if (p == 0 && blockIdx.x == 0 && threadIdx.x == 0) *n_total = total;
    const int32_t *tile = reinterpret_cast<const int32_t *>(sh);
    for (uint64_t first = (uint64_t) blockIdx.x * kPullTile; first < n; first += (uint64_t) gridDim.x * kPullTile) {
        const uint32_t count = (uint32_t) min((uint64_t) kPullTile, (uint64_t) n - first);
        __syncthreads();
        pull_tile(ps.seg[p], first, count, 12, sh);
        // the local copy is contiguous: written as plain words, coalesced; the host pulls again if it did not fit
        for (uint32_t w = threadIdx.x; w < 3 * count; w += blockDim.x) {
            const uint64_t dst = 3 * (base + first) + w;
            if (dst < 3 * cap_local) triples[dst] = tile[w];
        }
        if (base + first + count <= cap_local)
            for (uint32_t i = threadIdx.x; i < count; i += blockDim.x) atomicAdd(outdeg + (uint32_t) tile[3 * i], 1u);  // global id
    }
}

// ------------------------------------------------------------------------------------------------
// Phase 1, generic path.  Canonical order of the reference pushes is (L asc, c asc) and only the last 3 survive,
// so the result is the 3 largest (L, c): scan L downwards 32 lengths at a time and stop at 3 hits.
__device__ __forceinline__ void phase1_generic_read(const ReadsDev &R, const SeedTable &T, const PsDev &P, uint32_t b,
                                                    const Phase1Out &out, int lane) {
    const uint32_t lenb = R.len[b];
    int32_t rc[kSmallEdgesKept], ro[kSmallEdgesKept];
#pragma unroll
    for (int k = 0; k < kSmallEdgesKept; k++) rc[k] = -1, ro[k] = 0;
    int found = 0;
    int64_t l_hi = (int64_t) lenb - P.min_offset;
    if (l_hi > P.rs - 1) l_hi = P.rs - 1;
    if (l_hi > P.max_l) l_hi = P.max_l;
    const uint32_t *pb = read_ptr(R, b);
    if (lenb != 0 && flag_from(R, b) && l_hi >= P.lmin) {
        for (int32_t l_top = (int32_t) l_hi; l_top >= P.lmin && found < kSmallEdgesKept; l_top -= 32) {
            const int32_t L = l_top - lane;
            // per-lane: the (up to) 3 largest matching c at this L, t0 > t1 > t2 (kNone = empty)
            uint32_t t0 = kNone, t1 = kNone, t2 = kNone;
            int nh = 0;
            if (L >= P.lmin) {
                const uint32_t o = lenb - (uint32_t) L;
                const uint64_t w = bits64(pb, 2u * o) & P.seed_mask;
                probe_seed_window(T, w, (uint32_t) P.seed_nt, [&](uint32_t c) {
                    if (c == b) return;
                    if ((int64_t) R.len[c] < L) return;
                    // prefix(c, L) == suffix(b, L)
                    if (!equal_bits_aligned(pb, 2u * o, read_ptr(R, c), 2u * (uint32_t) L)) return;
                    nh++;
                    if (t0 == kNone || c > t0) { t2 = t1; t1 = t0; t0 = c; }
                    else if (t1 == kNone || c > t1) { t2 = t1; t1 = c; }
                    else if (t2 == kNone || c > t2) { t2 = c; }
                });
                if (nh > kSmallEdgesKept) nh = kSmallEdgesKept;
            }
            unsigned m = __ballot_sync(kFull, nh > 0);
            while (m && found < kSmallEdgesKept) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int cnt = __shfl_sync(kFull, nh, src);
                const uint32_t s0 = __shfl_sync(kFull, t0, src);
                const uint32_t s1 = __shfl_sync(kFull, t1, src);
                const uint32_t s2 = __shfl_sync(kFull, t2, src);
                const int32_t off = (int32_t) lenb - (l_top - src);
                for (int k = 0; k < cnt && found < kSmallEdgesKept; k++) {
                    const uint32_t c = k == 0 ? s0 : (k == 1 ? s1 : s2);
#pragma unroll
                    for (int q = 0; q < kSmallEdgesKept; q++)
                        if (q == found) rc[q] = (int32_t) c, ro[q] = off;
                    found++;
                }
            }
        }
    }
    if (lane < kSmallEdgesKept) {
        int32_t c = -1, o = 0;
#pragma unroll
        for (int q = 0; q < kSmallEdgesKept; q++)
            if (q == lane) c = rc[q], o = ro[q];
        if (c >= 0) emit_edge1(out, b, (uint32_t) c, (uint32_t) o, overhang_tail(pb, (uint32_t) o));
    }
}

// generic phase 1 over a queue of source reads (the reads the fast kernel handed back)
__global__ void __launch_bounds__(kThreads)
phase1_queue_kernel(ReadsDev R, SeedTable T, PsDev P, const uint32_t *__restrict__ queue,
                    const uint32_t *__restrict__ n_queue, const uint32_t *__restrict__ queue2,
                    const uint32_t *__restrict__ n_queue2, Phase1Out out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * kWarpsPerBlock;
    // queue = what the first fast pass gave up on; if it was long enough for the second fast pass to run
    // (kSecondPassMin), queue2 = what that one gave up on is what is left to do
    if (queue2 && *n_queue >= kSecondPassMin) {
        queue = queue2;
        n_queue = n_queue2;
    }
    const uint32_t n = *n_queue;
    for (uint32_t q = warp; q < n; q += n_warps) {
        phase1_generic_read(R, T, P, queue[q], out, lane);
    }
}

// phase-1 edge list -> (b, c, o) triples (staged interface)
__global__ void edges_to_triples_kernel(const Edge1 *__restrict__ list, const uint32_t *__restrict__ n_list,
                                        int32_t *__restrict__ triples) {
    const uint64_t n = *n_list;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const Edge1 e = list[i];
        triples[3 * i] = e.b;
        triples[3 * i + 1] = e.c;
        triples[3 * i + 2] = e.o;
    }
}

// Many entries overflowed their fixed-capacity rows (*n_over > kOverScanMax): rebuild the transposed graph in CSR form.  rev_off = scan(indeg);
// the first `cap` entries of every row come from the fixed rows, the rest from the overflow list, which takes its
// positions by counting indeg[c] down from the row's full size.
__global__ void rows_to_csr_kernel(const uint32_t *__restrict__ n_over, const uint32_t *__restrict__ indeg,
                                   const RevEntry *__restrict__ rows, uint32_t cap, uint32_t n_targets,
                                   const uint32_t *__restrict__ rev_off, RevEntry *__restrict__ rev, uint64_t rev_cap,
                                   uint32_t *rev_overflow) {
    if (*n_over <= kOverScanMax) return;
    if (rev_off[n_targets] > rev_cap) {  // the host reports ALGA_E_CAPACITY
        if (blockIdx.x == 0 && threadIdx.x == 0) *rev_overflow = 1u;
        return;
    }
    const uint64_t total = (uint64_t) n_targets * cap;
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < total; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t) (i / cap), j = (uint32_t) (i - (uint64_t) c * cap);
        if (j < indeg[c]) rev[rev_off[c] + j] = rows[i];
    }
}
__global__ void over_to_csr_kernel(const uint32_t *__restrict__ n_over, uint32_t over_cap, const Edge1 *__restrict__ over,
                                   uint32_t *indeg, const uint32_t *__restrict__ rev_off, RevEntry *__restrict__ rev,
                                   uint32_t n_targets, uint64_t rev_cap) {
    uint32_t n = *n_over;
    if (n <= kOverScanMax || rev_off[n_targets] > rev_cap) return;
    if (n > over_cap) n = over_cap;  // the host notices the overflow of the overflow list and reruns
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const Edge1 e = over[i];
        const uint32_t pos = rev_off[e.c] + atomicSub(indeg + e.c, 1u) - 1u;
        RevEntry r;
        r.b = e.b;
        r.o = e.o;
        r.t = e.t;
        rev[pos] = r;
    }
}

// ------------------------------------------------------------------------------------------------
__global__ void count_targets_kernel(const int32_t *__restrict__ triples, uint64_t n, uint32_t lo, uint32_t hi,
                                     uint32_t *indeg) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t) triples[3 * i + 1];
        if (c >= lo && c < hi) atomicAdd(indeg + (c - lo), 1u);
    }
}

__global__ void scatter_rev_triples_kernel(ReadsDev R, const int32_t *__restrict__ triples, uint64_t n, uint32_t c_lo,
                                           uint32_t c_hi, const uint32_t *__restrict__ rev_off, uint32_t *cursor,
                                           RevEntry *__restrict__ rev) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t c = (uint32_t) triples[3 * i + 1];
        if (c < c_lo || c >= c_hi) continue;
        const uint32_t pos = rev_off[c - c_lo] + atomicSub(cursor + (c - c_lo), 1u) - 1u;
        const int32_t b = triples[3 * i], o = triples[3 * i + 2];
        RevEntry r;
        r.b = b;
        r.o = o;
        // the overhang tail does not travel with exchanged triples: rebuild it from the (replicated) read
        r.t = overhang_tail(read_ptr(R, (uint32_t) b), (uint32_t) o);
        rev[pos] = r;
    }
}

// ------------------------------------------------------------------------------------------------
// Phase 2 helpers.

// In-neighbour list of one target read: entries (a, offset of c in a, len(a)).
struct NbrList {
    uint32_t *a;
    uint32_t *o;
    uint32_t *len;
    uint32_t cap;
};

// Lane-local: among the suffix-index candidates of window (c, L), the smallest verified source read
// b > floor (floor = -1 for the first call); *n_hits = number of verified candidates > floor.
__device__ __forceinline__ uint32_t scan_hits(const ReadsDev &R, const SeedTable &T, const PsDev &P,
                                              const uint32_t *__restrict__ pc, uint32_t c, int32_t L, int64_t floor_b,
                                              int *n_hits) {
    uint32_t best = kNone;
    int nh = 0;
    const uint64_t w = bits64(pc, 2u * (uint32_t) (L - P.seed_nt)) & P.seed_mask;
    probe_seed_window(T, w, (uint32_t) P.seed_nt, [&](uint32_t b) {
        if (b == c || (int64_t) b <= floor_b) return;
        const uint32_t lenb = R.len[b];
        if ((int64_t) lenb - P.min_offset < L) return;
        // suffix(b, L) == prefix(c, L)
        if (!equal_bits_aligned(read_ptr(R, b), 2u * (lenb - (uint32_t) L), pc, 2u * (uint32_t) L)) return;
        nh++;
        if (b < best) best = b;
    });
    *n_hits = nh;
    return best;
}

// Warp-cooperative replay of one accepted overlap (b -> c, overlap L): GraphCreatorPrefSuf.cpp:403-483.
// Removes b itself and every in-neighbour a whose edge (a -> c) is implied by (a -> b) + (b -> c), i.e.
// a[d .. d+o) == b[0 .. o) with d = offset(a,c) - o, then appends (b, o).  Returns false on overflow.
__device__ __forceinline__ bool replay_hit(const ReadsDev &R, uint32_t b, int32_t L, NbrList &lst, uint32_t &cnt,
                                           int lane) {
    const uint32_t lenb = R.len[b];
    const int32_t o = (int32_t) lenb - L;
    const uint32_t *pb = read_ptr(R, b);
    uint32_t w = 0;
    for (uint32_t base = 0; base < cnt; base += 32) {
        const uint32_t j = base + lane;
        bool keep = false;
        uint32_t a = 0, oa = 0, lena = 0;
        if (j < cnt) {
            a = lst.a[j];
            oa = lst.o[j];
            lena = lst.len[j];
            bool rm = (a == b);
            if (!rm && o > 0) {
                const int64_t d = (int64_t) oa - o;
                if (d >= 0 && (int64_t) lenb + d - (int64_t) lena >= 0)
                    rm = equal_bits_aligned(read_ptr(R, a), 2u * (uint32_t) d, pb, 2u * (uint32_t) o);
            }
            keep = !rm;
        }
        const unsigned m = __ballot_sync(kFull, keep);
        __syncwarp();
        if (keep) {
            const uint32_t pos = w + __popc(m & ((1u << lane) - 1u));
            lst.a[pos] = a;
            lst.o[pos] = oa;
            lst.len[pos] = lena;
        }
        w += __popc(m);
        __syncwarp();
    }
    if (w >= lst.cap) return false;
    if (lane == 0) {
        lst.a[w] = b;
        lst.o[w] = (uint32_t) o;
        lst.len[w] = lenb;
    }
    __syncwarp();
    cnt = w + 1;
    return true;
}

// Load row c of the reversed phase-1 graph and apply Graph::retainOnlySmallestOffset (Graph.cpp:348-387):
// one entry per source read, smallest offset wins.  Returns false when the row does not fit.
__device__ __forceinline__ bool load_rev_row(const ReadsDev &R, const RowsView &rows, bool csr, uint32_t ci, uint32_t c,
                                             NbrList &lst, uint32_t &cnt, int lane) {
    cnt = 0;
    uint32_t deg;
    const RevEntry *row = get_row(rows, csr, ci, deg);
    if (deg > lst.cap) return false;
    const uint32_t in_row = (!csr && deg > rows.cap) ? rows.cap : deg;
    for (uint32_t j = lane; j < in_row; j += 32) {
        const RevEntry e = row[j];
        lst.a[j] = (uint32_t) e.b;
        lst.o[j] = (uint32_t) e.o;
        lst.len[j] = 0;
    }
    if (in_row < deg) {  // the rest of this row is in the (short) overflow list
        uint32_t w = in_row;
        const uint32_t n_over = *rows.n_over;
        for (uint32_t base = 0; base < n_over; base += 32) {
            Edge1 e;
            e.c = -1;
            if (base + lane < n_over) e = rows.over[base + lane];
            const bool mine = (uint32_t) e.c == ci;  // the list holds row indices
            const unsigned m = __ballot_sync(kFull, mine);
            if (mine) {
                const uint32_t pos = w + __popc(m & ((1u << lane) - 1u));
                if (pos < deg) {
                    lst.a[pos] = (uint32_t) e.b;
                    lst.o[pos] = (uint32_t) e.o;
                    lst.len[pos] = 0;
                }
            }
            w += __popc(m);
        }
    }
    __syncwarp();
    if (deg > 1) {
        for (uint32_t j = lane; j < deg; j += 32) {
            const uint32_t a = lst.a[j], o = lst.o[j];
            bool dup = false;
            for (uint32_t k = 0; k < deg; k++) {
                if (k == j || lst.a[k] != a) continue;
                const uint32_t ok = lst.o[k];
                if (ok < o || (ok == o && k < j)) dup = true;
            }
            if (dup) lst.len[j] = kNone;
        }
        __syncwarp();
    }
    uint32_t w = 0;
    for (uint32_t base = 0; base < deg; base += 32) {
        const uint32_t j = base + lane;
        bool keep = false;
        uint32_t a = 0, o = 0;
        if (j < deg) {
            a = lst.a[j];
            o = lst.o[j];
            keep = lst.len[j] != kNone;
        }
        const unsigned m = __ballot_sync(kFull, keep);
        __syncwarp();
        if (keep) {
            const uint32_t pos = w + __popc(m & ((1u << lane) - 1u));
            lst.a[pos] = a;
            lst.o[pos] = o;
            lst.len[pos] = R.len[a];
        }
        w += __popc(m);
        __syncwarp();
    }
    cnt = w;
    return true;
}

// All of phase 2 for one target read c on one warp.  Returns false when the list overflowed.
__device__ __forceinline__ bool phase2_target(const ReadsDev &R, const SeedTable &T, const PsDev &P, uint32_t c,
                                              const RowsView &rows, bool csr, uint32_t ci, NbrList &lst,
                                              const Phase2Out &out, int lane) {
    uint32_t cnt = 0;
    if (!load_rev_row(R, rows, csr, ci, c, lst, cnt, lane)) return false;
    const uint32_t lenc = R.len[c];
    const int32_t l_lo = P.rs > P.lmin ? P.rs : P.lmin;
    int64_t l_hi = lenc;
    if (l_hi > P.max_l) l_hi = P.max_l;
    if (lenc != 0 && flag_to(R, c) && l_hi >= l_lo) {
        const uint32_t *pc = read_ptr(R, c);
        for (int32_t l_base = l_lo; l_base <= (int32_t) l_hi; l_base += 32) {
            const int32_t L = l_base + lane;
            int nh = 0;
            uint32_t bmin = kNone;
            if (L <= (int32_t) l_hi) bmin = scan_hits(R, T, P, pc, c, L, -1, &nh);
            unsigned m = __ballot_sync(kFull, nh > 0);
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const int32_t l_src = l_base + src;
                int left = __shfl_sync(kFull, nh, src);
                uint32_t b = __shfl_sync(kFull, bmin, src);
                while (true) {
                    if (!replay_hit(R, b, l_src, lst, cnt, lane)) return false;
                    if (--left == 0) break;
                    // several source reads share this (c, L): take them in ascending id order
                    uint32_t nxt = kNone;
                    if (lane == src) {
                        int dummy;
                        nxt = scan_hits(R, T, P, pc, c, l_src, (int64_t) b, &dummy);
                    }
                    b = __shfl_sync(kFull, nxt, src);
                    if (b == kNone) break;
                }
            }
        }
    }
    // emit the surviving in-neighbours of c as forward triples (a, c, offset)
    if (cnt && out.sh.world > 1) {
        for (uint32_t j = lane; j < cnt; j += 32) emit_triple_sharded(out.sh, (int32_t) lst.a[j], (int32_t) c, (int32_t) lst.o[j]);
    } else if (cnt) {
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(out.n_edges, (unsigned long long) cnt);
        base = __shfl_sync(kFull, base, 0);
        for (uint32_t j = lane; j < cnt; j += 32) {
            const unsigned long long pos = base + j;
            const uint32_t a = lst.a[j];
            if (pos < out.edge_cap) {
                out.triples[3 * pos] = (int32_t) a;
                out.triples[3 * pos + 1] = (int32_t) c;
                out.triples[3 * pos + 2] = (int32_t) lst.o[j];
            }
            if (out.outdeg) atomicAdd(out.outdeg + a, 1u);
        }
    }
    __syncwarp();
    return true;
}

__global__ void __launch_bounds__(kThreads)
phase2_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, uint32_t hi, RowsView rows, int list_cap, Phase2Out out) {
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31;
    const int wib = threadIdx.x >> 5;
    NbrList lst;
    lst.a = smem + (size_t) wib * 3 * list_cap;
    lst.o = lst.a + list_cap;
    lst.len = lst.o + list_cap;
    lst.cap = (uint32_t) list_cap;
    const uint32_t warp = blockIdx.x * kWarpsPerBlock + wib;
    const uint32_t n_warps = gridDim.x * kWarpsPerBlock;
    const bool csr = rows_are_csr(rows);
    for (uint64_t cc = (uint64_t) lo + warp; cc < hi; cc += n_warps) {
        const uint32_t c = (uint32_t) cc;
        if (!phase2_target(R, T, P, c, rows, csr, c - lo, lst, out, lane)) {
            if (lane == 0) out.spill_queue[atomicAdd(out.n_spill, 1u)] = c;
        }
        __syncwarp();
    }
}

// spill path, pass 1: upper bound of the list length of each queued target = row size + accepted overlaps
__global__ void __launch_bounds__(kThreads)
phase2_count_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, RowsView rows, const uint32_t *__restrict__ queue,
                    uint32_t n_queue, uint32_t *__restrict__ caps) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * kWarpsPerBlock;
    for (uint32_t q = warp; q < n_queue; q += n_warps) {
        const uint32_t c = queue[q];
        const uint32_t lenc = R.len[c];
        const int32_t l_lo = P.rs > P.lmin ? P.rs : P.lmin;
        int64_t l_hi = lenc;
        if (l_hi > P.max_l) l_hi = P.max_l;
        uint32_t total = 0;
        if (lenc != 0 && flag_to(R, c) && l_hi >= l_lo) {
            const uint32_t *pc = read_ptr(R, c);
            for (int32_t L = l_lo + lane; L <= (int32_t) l_hi; L += 32) {
                int nh = 0;
                scan_hits(R, T, P, pc, c, L, -1, &nh);
                total += (uint32_t) nh;
            }
        }
        for (int d = 16; d; d >>= 1) total += __shfl_xor_sync(kFull, total, d);
        if (lane == 0) {
            uint32_t deg;
            get_row(rows, rows_are_csr(rows), c - lo, deg);
            caps[q] = total + deg + 1u;
        }
    }
}

// spill path, pass 2: same replay with the list in global memory (capacity from pass 1)
__global__ void __launch_bounds__(kThreads)
phase2_spill_kernel(ReadsDev R, SeedTable T, PsDev P, uint32_t lo, RowsView rows, const uint32_t *__restrict__ queue,
                    uint32_t n_queue, const uint64_t *__restrict__ spill_off, uint32_t *spill_store, Phase2Out out) {
    const int lane = threadIdx.x & 31;
    const uint32_t warp = blockIdx.x * kWarpsPerBlock + (threadIdx.x >> 5);
    const uint32_t n_warps = gridDim.x * kWarpsPerBlock;
    for (uint32_t q = warp; q < n_queue; q += n_warps) {
        const uint32_t c = queue[q];
        const uint64_t s0 = spill_off[q], s1 = spill_off[q + 1];
        NbrList lst;
        lst.cap = (uint32_t) (s1 - s0);
        lst.a = spill_store + 3 * s0;
        lst.o = lst.a + lst.cap;
        lst.len = lst.o + lst.cap;
        phase2_target(R, T, P, c, rows, rows_are_csr(rows), c - lo, lst, out, lane);
        __syncwarp();
    }
}

// ------------------------------------------------------------------------------------------------
// CSR assembly
__global__ void count_sources_kernel(const int32_t *__restrict__ triples, uint64_t n, uint32_t lo, uint32_t hi, int swap,
                                     uint32_t *outdeg) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t) triples[3 * i + (swap ? 1 : 0)];
        if (b >= lo && b < hi) atomicAdd(outdeg + (b - lo), 1u);
    }
}

__global__ void scatter_csr_kernel(const int32_t *__restrict__ triples, uint64_t n, uint32_t lo, uint32_t hi, int swap,
                                   const uint64_t *__restrict__ row_off, uint32_t *cursor, int32_t *__restrict__ nbr,
                                   int32_t *__restrict__ off) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t) triples[3 * i + (swap ? 1 : 0)];
        if (b < lo || b >= hi) continue;
        const uint64_t pos = row_off[b - lo] + atomicSub(cursor + (b - lo), 1u) - 1u;
        nbr[pos] = triples[3 * i + (swap ? 0 : 1)];
        off[pos] = triples[3 * i + 2];
    }
}

// The same scatter with two random accesses per edge instead of five (fewer than 2^32 edges): the cursor of a row starts at
// the END of the row (one streaming pass over the offsets), so the position is the result of the atomic alone, and
// neighbour and offset go out as ONE 8-byte store; a streaming pass splits the pairs into the two result arrays.
__global__ void end_cursor_kernel(const uint64_t *__restrict__ row_off, uint32_t n, uint32_t *__restrict__ cursor) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x)
        cursor[i] = (uint32_t) row_off[i + 1];
}
__global__ void scatter_pairs_kernel(const int32_t *__restrict__ triples, uint64_t n, uint32_t lo, uint32_t hi, int swap,
                                     uint32_t *cursor, int2 *__restrict__ pairs) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const uint32_t b = (uint32_t) triples[3 * i + (swap ? 1 : 0)];
        if (b < lo || b >= hi) continue;
        const uint32_t pos = atomicSub(cursor + (b - lo), 1u) - 1u;
        pairs[pos] = make_int2(triples[3 * i + (swap ? 0 : 1)], triples[3 * i + 2]);
    }
}
__global__ void split_pairs_kernel(const int2 *__restrict__ pairs, uint64_t n, int32_t *__restrict__ nbr, int32_t *__restrict__ off) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const int2 v = pairs[i];
        nbr[i] = v.x;
        off[i] = v.y;
    }
}

__device__ __forceinline__ bool edge_less(int32_t n1, int32_t o1, int32_t n2, int32_t o2) {
    return n1 < n2 || (n1 == n2 && o1 < o2);
}

__global__ void sort_rows_kernel(const uint64_t *__restrict__ row_off, uint32_t n_rows, int32_t *nbr, int32_t *off,
                                 uint32_t *big_rows, uint32_t *n_big) {
    for (uint64_t r = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; r < n_rows; r += (uint64_t) gridDim.x * blockDim.x) {
        const uint64_t s = row_off[r];
        const uint32_t d = (uint32_t) (row_off[r + 1] - s);
        if (d < 2) continue;
        if (d > 32) {
            big_rows[atomicAdd(n_big, 1u)] = (uint32_t) r;
            continue;
        }
        for (uint32_t i = 1; i < d; i++) {
            const int32_t kn = nbr[s + i], ko = off[s + i];
            uint32_t j = i;
            while (j > 0 && edge_less(kn, ko, nbr[s + j - 1], off[s + j - 1])) {
                nbr[s + j] = nbr[s + j - 1];
                off[s + j] = off[s + j - 1];
                j--;
            }
            nbr[s + j] = kn;
            off[s + j] = ko;
        }
    }
}

// rank sort of long rows, one block per row (entries of a row are distinct by (nbr, off) or tie-broken by index)
__global__ void sort_big_rows_kernel(const uint64_t *__restrict__ row_off, const uint32_t *__restrict__ big_rows,
                                     uint32_t n_big, int32_t *nbr, int32_t *off, int32_t *tmp_nbr, int32_t *tmp_off) {
    for (uint32_t q = blockIdx.x; q < n_big; q += gridDim.x) {
        const uint32_t r = big_rows[q];
        const uint64_t s = row_off[r];
        const uint32_t d = (uint32_t) (row_off[r + 1] - s);
        for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {
            const int32_t kn = nbr[s + i], ko = off[s + i];
            uint32_t rank = 0;
            for (uint32_t k = 0; k < d; k++) {
                const int32_t n2 = nbr[s + k], o2 = off[s + k];
                if (edge_less(n2, o2, kn, ko) || (n2 == kn && o2 == ko && k < i)) rank++;
            }
            tmp_nbr[s + rank] = kn;
            tmp_off[s + rank] = ko;
        }
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < d; i += blockDim.x) {
            nbr[s + i] = tmp_nbr[s + i];
            off[s + i] = tmp_off[s + i];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// Exclusive scan: tile sums -> scan of tile sums (single block) -> per-tile scan.
constexpr int kScanThreads = 512;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total, uint64_t *sh /*[33]*/) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
    for (int d = 1; d < 32; d <<= 1) {
        const uint64_t y = __shfl_up_sync(kFull, x, d);
        if (lane >= d) x += y;
    }
    if (lane == 31) sh[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = lane < (int) (blockDim.x >> 5) ? sh[lane] : 0;
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(kFull, s, d);
            if (lane >= d) s += y;
        }
        sh[lane] = s;  // inclusive warp totals
    }
    __syncthreads();
    const uint64_t warp_base = wid ? sh[wid - 1] : 0;
    *total = sh[(blockDim.x >> 5) - 1];
    __syncthreads();
    return warp_base + x - v;
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                      uint64_t *__restrict__ tile_sums,
                                                                      const uint32_t *__restrict__ run_if) {
    __shared__ uint64_t sh[33];
    if (run_if && *run_if <= kOverScanMax) return;
    const uint64_t base = (uint64_t) blockIdx.x * kScanTile;
    uint64_t v = 0;
    for (int k = 0; k < kScanItems; k++) {
        const uint64_t i = base + (uint64_t) k * kScanThreads + threadIdx.x;
        if (i < n) v += in[i];
    }
    uint64_t total;
    block_exclusive_scan(v, &total, sh);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_spine_kernel(uint64_t *tile_sums, uint64_t n_tiles,
                                                                  const uint32_t *__restrict__ run_if) {
    __shared__ uint64_t sh[33];
    if (run_if && *run_if <= kOverScanMax) return;
    uint64_t carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += kScanThreads) {
        const uint64_t i = base + threadIdx.x;
        const uint64_t v = i < n_tiles ? tile_sums[i] : 0;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan(v, &total, sh);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

template <class OutT>
__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(const uint32_t *__restrict__ in, OutT *__restrict__ out,
                                                                  uint64_t n, const uint64_t *__restrict__ tile_sums,
                                                                  uint64_t n_tiles, const uint32_t *__restrict__ run_if) {
    __shared__ uint64_t sh[33];
    if (run_if && *run_if <= kOverScanMax) return;
    const uint64_t base = (uint64_t) blockIdx.x * kScanTile + (uint64_t) threadIdx.x * kScanItems;
    uint32_t item[kScanItems];
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        item[k] = base + k < n ? in[base + k] : 0u;
        v += item[k];
    }
    uint64_t total;
    uint64_t run = tile_sums[blockIdx.x] + block_exclusive_scan(v, &total, sh);
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = (OutT) run;
        run += item[k];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = (OutT) tile_sums[n_tiles];
}

__global__ void fill_u64_kernel(uint64_t *p, uint64_t v, uint64_t n) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) p[i] = v;
}

}  // namespace

// ================================================================================================
// launchers

void launch_read_stats(const ReadsDev &R, int lmin, int min_offset, ReadStats *d_stats, cudaStream_t s,
                       const LaunchCfg &cfg) {
    cudaMemsetAsync(d_stats, 0, sizeof(ReadStats), s);
    cudaMemsetAsync(&d_stats->min_len, 0xFF, sizeof(uint32_t), s);
    read_stats_kernel<<<grid_for(R.n, 256, cfg), 256, 0, s>>>(R, lmin, min_offset, d_stats);
    bump(cfg);
}

void launch_build_index(const ReadsDev &R, const PsDev &P, SeedTable prefix, SeedTable suffix, uint32_t lo, uint32_t hi,
                        uint32_t b_lo, uint32_t b_hi, cudaStream_t s, const LaunchCfg &cfg, int which) {
    if (hi <= lo) return;
    build_index_kernel<<<grid_for(hi - lo, 256, cfg), 256, 0, s>>>(R, P, prefix, suffix, lo, hi, b_lo, b_hi, which);
    bump(cfg);
}

void launch_seed_keys(const uint32_t *words, uint32_t stride, uint32_t n, const PsDev &P, SeedTable prefix, SeedTable suffix,
                      uint32_t *keys, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    seed_keys_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(words, stride, n, P, prefix, suffix, keys);
    bump(cfg);
}
void launch_index_keys(const uint32_t *keys, uint32_t first_id, uint32_t n, SeedTable prefix, SeedTable suffix, uint32_t b_lo,
                       uint32_t b_hi, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    index_keys_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(keys, first_id, n, prefix, suffix, b_lo, b_hi);
    bump(cfg);
}

void launch_pull_rows(const void *const *seg, const uint32_t *const *cnt, int world, uint32_t cap, uint32_t n_expected,
                      const Phase1Out &out, cudaStream_t s, const LaunchCfg &cfg) {
    PeerSegs ps{};
    for (int p = 0; p < world; p++) ps.seg[p] = seg[p], ps.cnt[p] = cnt[p];
    ps.cap = cap;
    ps.world = world;
    dim3 grid((unsigned) grid_for(n_expected ? n_expected : 1, kPullTile, cfg, world >= 8 ? 1 : (world >= 4 ? 2 : 4)),
              (unsigned) world);
    pull_rows_kernel<<<grid, 256, 0, s>>>(ps, out);
    bump(cfg);
}

void launch_pull_triples(const void *const *seg, const uint32_t *const *cnt, int world, uint32_t cap, uint32_t n_expected,
                         uint64_t cap_local, int32_t *triples, unsigned long long *n_total, uint32_t *outdeg,
                         cudaStream_t s, const LaunchCfg &cfg) {
    PeerSegs ps{};
    for (int p = 0; p < world; p++) ps.seg[p] = seg[p], ps.cnt[p] = cnt[p];
    ps.cap = cap;
    ps.world = world;
    dim3 grid((unsigned) grid_for(n_expected ? n_expected : 1, kPullTile, cfg, world >= 8 ? 1 : (world >= 4 ? 2 : 4)),
              (unsigned) world);
    pull_triples_kernel<<<grid, 256, 0, s>>>(ps, cap_local, triples, n_total, outdeg);
    bump(cfg);
}

void launch_phase1_queue(const ReadsDev &R, const SeedTable &prefix, const PsDev &P, uint32_t n_max,
                         const uint32_t *hard_queue, const uint32_t *n_hard, const uint32_t *hard_queue2,
                         const uint32_t *n_hard2, const Phase1Out &out, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_max) return;
    // the queue length lives on the device (no host round trip; usually zero): size the grid for a short queue
    // unless the caller knows that every read is in it
    phase1_queue_kernel<<<grid_for(n_max, kWarpsPerBlock, cfg, 8), kThreads, 0, s>>>(R, prefix, P, hard_queue, n_hard,
                                                                                       hard_queue2, n_hard2, out);
    bump(cfg);
}

void launch_edges_to_triples(const Edge1 *list, const uint32_t *n_list, uint64_t n_max, int32_t *triples, cudaStream_t s,
                             const LaunchCfg &cfg) {
    if (!n_max) return;
    edges_to_triples_kernel<<<grid_for(n_max, 256, cfg), 256, 0, s>>>(list, n_list, triples);
    bump(cfg);
}

void launch_rebuild_rows_csr(const uint32_t *n_over, uint32_t over_cap, const Edge1 *over, uint32_t *indeg,
                             const RevEntry *rows, uint32_t cap, uint32_t n_targets, uint32_t *rev_off, RevEntry *rev,
                             uint64_t rev_cap, uint32_t *rev_overflow, void *scan_ws, cudaStream_t s,
                             const LaunchCfg &cfg) {
    if (!n_targets) return;
    launch_scan_u32(indeg, rev_off, n_targets, scan_ws, s, cfg, n_over);
    rows_to_csr_kernel<<<grid_for((uint64_t) n_targets * cap, 256, cfg), 256, 0, s>>>(n_over, indeg, rows, cap, n_targets,
                                                                                       rev_off, rev, rev_cap, rev_overflow);
    over_to_csr_kernel<<<grid_for(over_cap < 65536 ? over_cap : 65536, 256, cfg), 256, 0, s>>>(
        n_over, over_cap, over, indeg, rev_off, rev, n_targets, rev_cap);
    bump(cfg);
    bump(cfg);
}

void launch_count_targets(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, uint32_t *indeg,
                          cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    count_targets_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(triples, n, lo, hi, indeg);
    bump(cfg);
}

void launch_scatter_rev_triples(const ReadsDev &R, const int32_t *triples, uint64_t n, uint32_t c_lo, uint32_t c_hi,
                                const uint32_t *rev_off, uint32_t *cursor, RevEntry *rev, cudaStream_t s,
                                const LaunchCfg &cfg) {
    if (!n) return;
    scatter_rev_triples_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(R, triples, n, c_lo, c_hi, rev_off, cursor, rev);
    bump(cfg);
}

void launch_phase2(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t lo, uint32_t hi,
                   const RowsView &rows, int list_cap, const Phase2Out &out, cudaStream_t s, const LaunchCfg &cfg) {
    if (hi <= lo) return;
    const size_t smem = (size_t) kWarpsPerBlock * 3 * list_cap * sizeof(uint32_t);
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(phase2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int) smem);
    phase2_kernel<<<grid_for(hi - lo, kWarpsPerBlock, cfg, 8), kThreads, smem, s>>>(R, suffix, P, lo, hi, rows, list_cap,
                                                                                      out);
    bump(cfg);
}

void launch_phase2_count(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t lo, const RowsView &rows,
                         const uint32_t *queue, uint32_t n_queue, uint32_t *caps, cudaStream_t s,
                         const LaunchCfg &cfg) {
    if (!n_queue) return;
    phase2_count_kernel<<<grid_for(n_queue, kWarpsPerBlock, cfg, 8), kThreads, 0, s>>>(R, suffix, P, lo, rows, queue,
                                                                                         n_queue, caps);
    bump(cfg);
}

void launch_phase2_spill(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t lo, const RowsView &rows,
                         const uint32_t *queue, uint32_t n_queue, const uint64_t *spill_off, uint32_t *spill_store,
                         const Phase2Out &out, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_queue) return;
    phase2_spill_kernel<<<grid_for(n_queue, kWarpsPerBlock, cfg, 8), kThreads, 0, s>>>(
        R, suffix, P, lo, rows, queue, n_queue, spill_off, spill_store, out);
    bump(cfg);
}

void launch_count_sources(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, uint32_t *outdeg,
                          cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    count_sources_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(triples, n, lo, hi, swap, outdeg);
    bump(cfg);
}

void launch_scatter_csr(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, const uint64_t *row_off,
                        uint32_t *cursor, int32_t *nbr, int32_t *off, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    scatter_csr_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(triples, n, lo, hi, swap, row_off, cursor, nbr, off);
    bump(cfg);
}

void launch_scatter_csr_pairs(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, const uint64_t *row_off,
                              uint32_t *cursor, void *pairs, int32_t *nbr, int32_t *off, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    end_cursor_kernel<<<grid_for(hi - lo, 256, cfg), 256, 0, s>>>(row_off, hi - lo, cursor);
    scatter_pairs_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(triples, n, lo, hi, swap, cursor, (int2 *) pairs);
    split_pairs_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>((const int2 *) pairs, n, nbr, off);
    bump(cfg);
    bump(cfg);
    bump(cfg);
}

void launch_sort_rows(const uint64_t *row_off, uint32_t n_rows, int32_t *nbr, int32_t *off, uint32_t *big_rows,
                      uint32_t *n_big, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_rows) return;
    sort_rows_kernel<<<grid_for(n_rows, 256, cfg), 256, 0, s>>>(row_off, n_rows, nbr, off, big_rows, n_big);
    bump(cfg);
}

void launch_sort_big_rows(const uint64_t *row_off, const uint32_t *big_rows, uint32_t n_big, int32_t *nbr, int32_t *off,
                          int32_t *tmp_nbr, int32_t *tmp_off, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_big) return;
    sort_big_rows_kernel<<<grid_for(n_big, 1, cfg, 8), 256, 0, s>>>(row_off, big_rows, n_big, nbr, off, tmp_nbr,
                                                                      tmp_off);
    bump(cfg);
}

size_t scan_workspace_bytes(uint64_t n) {
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    return (size_t) (tiles + 2) * sizeof(uint64_t);
}

template <class OutT>
static void launch_scan_impl(const uint32_t *in, OutT *out, uint64_t n, void *workspace, cudaStream_t s,
                             const LaunchCfg &cfg, const uint32_t *run_if = nullptr) {
    uint64_t *tile_sums = (uint64_t *) workspace;
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    if (n == 0) {
        cudaMemsetAsync(out, 0, sizeof(OutT), s);
        return;
    }
    scan_tile_sums_kernel<<<(unsigned) tiles, kScanThreads, 0, s>>>(in, n, tile_sums, run_if);
    scan_spine_kernel<<<1, kScanThreads, 0, s>>>(tile_sums, tiles, run_if);
    scan_apply_kernel<OutT><<<(unsigned) tiles, kScanThreads, 0, s>>>(in, out, n, tile_sums, tiles, run_if);
    bump(cfg);
    bump(cfg);
    bump(cfg);
}

void launch_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *workspace, cudaStream_t s,
                     const LaunchCfg &cfg, const uint32_t *run_if) {
    launch_scan_impl<uint32_t>(in, out, n, workspace, s, cfg, run_if);
}
void launch_scan_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *workspace, cudaStream_t s,
                     const LaunchCfg &cfg) {
    launch_scan_impl<uint64_t>(in, out, n, workspace, s, cfg);
}

void launch_fill_u64(uint64_t *p, uint64_t v, uint64_t n, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n) return;
    fill_u64_kernel<<<grid_for(n, 256, cfg), 256, 0, s>>>(p, v, n);
    bump(cfg);
}

}  // namespace alga

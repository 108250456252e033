// Device-side building blocks shared by the overlap-graph kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace alga {

constexpr uint32_t kNone = 0xFFFFFFFFu;
constexpr int kBucketWords = 32;  // a bucket of the seed index = 128 bytes = one L2 line
constexpr int kBucketCap = 20;    // entries per bucket
#ifndef ALGA_NM
#define ALGA_NM 9
#endif
constexpr int kNM = ALGA_NM;      // m-mers per seed window: the minimizer is taken over m = seed_nt - (kNM - 1) nucleotides
constexpr int kSmallEdgesKept = 3;  // SOES, GraphCreatorPrefSuf.h:62
constexpr int kHeadWords = 4;       // cached head of a read: first 64 nucleotides
constexpr int kHeadNt = kHeadWords * 16;
constexpr int kReadPadBytes = 256;  // readable bytes required after the packed words (speculative window loads)

// Packed read set resident in HBM, reference layout (Bitset.h:38-45).  `words` is followed by at
// least kReadPadBytes readable bytes so that window / compare loads near the end stay in bounds.
struct ReadsDev {
    const uint32_t *__restrict__ words;
    const uint64_t *__restrict__ word_off;  // nullptr in fixed-stride mode
    const uint32_t *__restrict__ len;
    const uint8_t *__restrict__ from;  // may be nullptr (all true)
    const uint8_t *__restrict__ to;    // may be nullptr (all true)
    uint32_t n;
    uint32_t stride;  // words per read when word_off == nullptr
};

__device__ __forceinline__ const uint32_t *read_ptr(const ReadsDev &r, uint32_t i) {
    return r.words + (r.word_off ? r.word_off[i] : (uint64_t) i * r.stride);
}
__device__ __forceinline__ bool flag_from(const ReadsDev &r, uint32_t i) { return r.from ? r.from[i] != 0 : true; }
__device__ __forceinline__ bool flag_to(const ReadsDev &r, uint32_t i) { return r.to ? r.to[i] != 0 : true; }

// Parameters of one GraphCreatorPrefSuf run, resolved on the host.
struct PsDev {
    int32_t lmin;        // MIN_OVERLAP_PREF_SUF
    int32_t rs;          // REMOVE_SMALL_OVERLAP_EDGES_MIN_OVERLAP
    int32_t min_offset;  // MIN_OFFSET_FOR_ALIGNMENT
    int32_t max_l;       // last overlap length the reference iterates: min(max read length, cap) + 1  (GraphCreatorPrefSuf.cpp:92-95)
    int32_t seed_nt;     // K = min(lmin, 32): nucleotides hashed into the seed index
    uint32_t uniform_len;  // != 0: every read (none removed) has exactly this length -> no per-candidate length loads
    uint64_t seed_mask;  // low 2K bits
};

// One entry of the transposed phase-1 graph (row of a target read): source read, offset, overhang tail.  16 bytes =
// one 128-bit load / store.
struct __align__(16) RevEntry {
    int32_t b, o;
    uint64_t t;
};

// One phase-1 edge in list form (sharded runs, overflow of the fixed-capacity rows): b -> c at offset o with b's tail.
struct __align__(8) Edge1 {
    int32_t c, b, o, pad;
    uint64_t t;
};

// Sharded runs (one process per GPU): what a rank produces for reads owned by another rank is appended to the
// per-destination segment of the rank's own exchange workspace, which the owner later reads over NVLink.
struct ShardOut {
    int world;         // <= 1: not sharded
    uint32_t n_shard;  // reads owned per rank; owner(id) = min(id / n_shard, world - 1)
    uint32_t *cnt;     // [world] entries appended per destination (keeps counting past cap: the host checks)
    void *seg;         // world segments of `cap` entries each
    uint32_t cap;
};
__device__ __forceinline__ uint32_t shard_of(const ShardOut &sh, uint32_t id) {
    const uint32_t d = id / sh.n_shard;
    return d < (uint32_t) sh.world ? d : (uint32_t) sh.world - 1u;
}
// Warp-aggregated reservation of one slot per valid lane in the segment of its destination (one atomicAdd per
// distinct destination in the warp).  Every lane of the warp must call it.
__device__ __forceinline__ uint32_t shard_reserve(const ShardOut &sh, bool valid, uint32_t dest, int lane) {
    const unsigned grp = __match_any_sync(0xFFFFFFFFu, valid ? dest : (0x80000000u | (uint32_t) lane));
    const int leader = __ffs(grp) - 1;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(sh.cnt + dest, (uint32_t) __popc(grp));
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    return base + (uint32_t) __popc(grp & ((1u << lane) - 1u));
}

// Where phase 1 puts its edges (GraphCreatorPrefSuf.cpp:397-402 pushes to G[b]; phase 2 needs them by target c):
//   mode 0: straight into the row of the target read in the transposed graph -- `row_cap` entries per target,
//           position = atomicAdd(indeg[c]); entries beyond the capacity go to `list` (then *n_list != 0 and the
//           rows are rebuilt in CSR form before phase 2);
//   mode 1: appended to `list` (one warp-aggregated atomicAdd per warp);
//   mode 2: appended to the segment of the rank that owns the target read (`sh`, Edge1 entries).
// `c_base` is subtracted from the target id wherever it indexes rows (sharded: rows cover the rank's own reads).
struct Phase1Out {
    int mode;
    uint32_t *indeg;
    RevEntry *rows;
    uint32_t row_cap;
    Edge1 *list;
    uint32_t *n_list;
    uint32_t list_cap;
    uint32_t c_base;
    ShardOut sh;
};

// one edge, one thread (generic kernels; the fast kernel aggregates its list appends per warp)
__device__ __forceinline__ void emit_edge1(const Phase1Out &out, uint32_t b, uint32_t c, uint32_t o, uint64_t t) {
    if (out.mode == 2) {
        const uint32_t d = shard_of(out.sh, c);
        const uint32_t i = atomicAdd(out.sh.cnt + d, 1u);
        if (i < out.sh.cap) {
            Edge1 x;
            x.c = (int32_t) c, x.b = (int32_t) b, x.o = (int32_t) o, x.pad = 0, x.t = t;
            reinterpret_cast<Edge1 *>(out.sh.seg)[(uint64_t) d * out.sh.cap + i] = x;
        }
        return;
    }
    if (out.mode == 0) {
        const uint32_t ci = c - out.c_base;
        const uint32_t pos = atomicAdd(out.indeg + ci, 1u);
        if (pos < out.row_cap) {
            RevEntry r;
            r.b = (int32_t) b, r.o = (int32_t) o, r.t = t;
            out.rows[(uint64_t) ci * out.row_cap + pos] = r;
            return;
        }
    }
    const uint32_t i = atomicAdd(out.n_list, 1u);
    if (i < out.list_cap) {
        Edge1 x;
        x.c = (int32_t) (c - out.c_base), x.b = (int32_t) b, x.o = (int32_t) o, x.pad = 0, x.t = t;
        out.list[i] = x;
    }
}

// Rows of the transposed phase-1 graph as phase 2 reads them.  Fixed-capacity form: row i holds its first `cap`
// entries, indeg[i] is its full size, the few entries beyond the capacity sit in the overflow list `over` (targets
// with indeg > cap take the generic kernels, which pick their entries out of that list).  When cap == 0 or the
// overflow list is longer than kOverScanMax the CSR form (rev_off, rev) is authoritative.  Indexed by c - lo.
constexpr uint32_t kOverScanMax = 2048;
// queues shorter than this skip the second fast pass (tpr_kernels.cu) and go straight to the generic kernels
constexpr uint32_t kSecondPassMin = 4096;
struct RowsView {
    const uint32_t *indeg;
    const RevEntry *rows;
    uint32_t cap;
    const uint32_t *n_over;
    const Edge1 *over;
    const uint32_t *rev_off;
    const RevEntry *rev;
};
__device__ __forceinline__ bool rows_are_csr(const RowsView &v) { return v.cap == 0 || *v.n_over > kOverScanMax; }
__device__ __forceinline__ const RevEntry *get_row(const RowsView &v, bool csr, uint32_t i, uint32_t &deg) {
    if (csr) {
        const uint32_t r0 = v.rev_off[i];
        deg = v.rev_off[i + 1] - r0;
        return v.rev + r0;
    }
    deg = v.indeg[i];
    return v.rows + (uint64_t) i * v.cap;
}

// Seed index.  A bucket is one 128-byte L2 line:
//     word 0        number of inserts that chose this bucket (beyond kBucketCap: they went on to the next bucket)
//     word 1        unused
//     words 2..11   20 tags, 16 bits each (tag_of(); 0 = empty slot)
//     words 12..31  20 read ids
// Why a whole line: on B200 a random access that misses L2 costs one REQUEST whatever it carries -- 39 G requests/s for
// 32-byte sectors and 38 G/s for whole 128-byte lines when the lanes of ONE load instruction cover the line
// (scripts/probes/random_coop.cu) -- so the unit worth fetching is the line.  Why so many entries: the bucket of a seed
// window is chosen by the MINIMIZER of the window (its smallest scrambled m-mer), not by the window: equal windows still
// meet in one bucket, and the windows a read probes at consecutive overlap lengths share theirs (5 distinct buckets for
// the 29 lengths of phase 2 with m = 20), but all reads that start within a few nucleotides of one another land in the
// same bucket too -- 5 entries on average where a probe looks, more than 20 in 0.2 % of the probes (then the chain goes on
// in the next bucket).  The 16-bit tag is taken from the hash of the whole window.
struct SeedTable {
    uint32_t *slots;
    uint32_t n_buckets;
    uint32_t slice;  // chains wrap inside slices of this many buckets (= n_buckets unless the build is sharded: rank
                     // r fills the buckets [r * slice, (r + 1) * slice) and the ranks then exchange slices)
    uint32_t min_m;  // m of the minimizer; 0: the bucket is chosen by the hash of the key itself (dictionary use, preprocess.cu)
};

// 32 bits of a packed read starting at bit position `bit`.
__device__ __forceinline__ uint32_t bits32(const uint32_t *__restrict__ p, uint32_t bit) {
    const uint32_t w = bit >> 5;
    return __funnelshift_r(__ldg(p + w), __ldg(p + w + 1), bit & 31u);
}
// 64 bits of a packed read starting at bit position `bit`.
__device__ __forceinline__ uint64_t bits64(const uint32_t *__restrict__ p, uint32_t bit) {
    const uint32_t w = bit >> 5, s = bit & 31u;
    const uint32_t a = __ldg(p + w), b = __ldg(p + w + 1), c = __ldg(p + w + 2);
    return (uint64_t) __funnelshift_r(a, b, s) | ((uint64_t) __funnelshift_r(b, c, s) << 32);
}

// nbits of pa starting at bit offset bita == nbits of pb starting at word boundary 0 ?
// Loads are issued 8 words at a time with no data-dependent branch in between, so that the (mostly
// successful) comparisons cost one memory round trip per 128 nucleotides instead of one per word.
// May read up to 9 words past the compared range (the read buffer is padded, see kReadPadBytes).
__device__ __forceinline__ bool equal_bits_aligned(const uint32_t *__restrict__ pa, uint32_t bita,
                                                   const uint32_t *__restrict__ pb, uint32_t nbits) {
    const uint32_t s = bita & 31u;
    const uint32_t *q = pa + (bita >> 5);
    const uint32_t n_words = (nbits + 31u) >> 5;
    for (uint32_t k0 = 0; k0 < n_words; k0 += 8) {
        uint32_t wa[9], wb[8];
#pragma unroll
        for (int j = 0; j < 9; j++) wa[j] = (k0 + j <= n_words) ? __ldg(q + k0 + j) : 0u;
#pragma unroll
        for (int j = 0; j < 8; j++) wb[j] = (k0 + j < n_words) ? __ldg(pb + k0 + j) : 0u;
        uint32_t diff = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const uint32_t k = k0 + j;
            uint32_t x = __funnelshift_r(wa[j], wa[j + 1], s) ^ wb[j];
            const uint32_t bits_left = k < n_words ? nbits - 32u * k : 0u;
            const uint32_t m = bits_left >= 32u ? 0xFFFFFFFFu : ((1u << bits_left) - 1u);
            diff |= x & m;
        }
        if (diff) return false;
    }
    return true;
}

// Same test on cached 128-bit heads (first 64 nucleotides): bits [bita, bita+nbits) of `a` against bits
// [0, nbits) of `b`; requires bita + nbits <= 128.
__device__ __forceinline__ bool equal_bits_head(const uint32_t (&a)[4], uint32_t bita, const uint32_t (&b)[4],
                                                uint32_t nbits) {
    const uint32_t ws = bita >> 5, s = bita & 31u;
    uint32_t diff = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        // word (ws + j) and (ws + j + 1) of a, zero beyond the head
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for (int t = 0; t < 4; t++) {
            if ((uint32_t) t == ws + j) lo = a[t];
            if ((uint32_t) t == ws + j + 1) hi = a[t];
        }
        const uint32_t x = __funnelshift_r(lo, hi, s) ^ b[j];
        const uint32_t bits_left = nbits > 32u * j ? nbits - 32u * j : 0u;
        const uint32_t m = bits_left >= 32u ? 0xFFFFFFFFu : ((1u << bits_left) - 1u);
        diff |= x & m;
    }
    return diff == 0;
}

// Hash of the 2K-bit seed window: pair-multiply-shift over its two 32-bit halves (four IMADs).  The high word picks
// the bucket, the top bits of the low word are the tag; a false tag hit only costs one exact compare.
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    return (uint64_t) (uint32_t) x * 0x9E3779B97F4A7C15ull + (x >> 32) * 0xC2B2AE3D27D4EB4Full;
}
__device__ __forceinline__ uint32_t bucket_of(uint64_t h, uint32_t n_buckets) {
    return __umulhi((uint32_t) (h >> 32), n_buckets);
}
// Minimizer of a seed window (`win` = its 2 * seed_nt bits): the smallest scrambled m-mer.  Equal windows have equal
// minimizers, so they still meet in one bucket; neighbouring windows of a read mostly share theirs.
__device__ __forceinline__ uint32_t window_minimizer(uint64_t win, uint32_t seed_nt, uint32_t m) {
    const uint64_t mask = (1ull << (2u * m)) - 1ull;  // m < 32
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t j = 0; j + m <= seed_nt; j++) {
        const uint64_t x = (win >> (2u * j)) & mask;
        uint32_t v = ((uint32_t) x ^ ((uint32_t) (x >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
        v ^= v >> 15;
        best = min(best, v);
    }
    return best;
}
// the scrambled m-mer at position j of a window, and the minimum over the window, for host and device (the kernels use
// window_minimizer above; tests/minimizer_check.cu compares the sliding form below with this one on the CPU)
__host__ __device__ __forceinline__ uint32_t scrambled_mmer(uint64_t win, uint32_t j, uint32_t m) {
    const uint64_t x = (win >> (2u * j)) & ((1ull << (2u * m)) - 1ull);  // m < 32
    uint32_t v = ((uint32_t) x ^ ((uint32_t) (x >> 32) * 0x85EBCA6Bu)) * 0x9E3779B1u;
    return v ^ (v >> 15);
}
__host__ __device__ __forceinline__ uint32_t window_minimizer_ref(uint64_t win, uint32_t seed_nt, uint32_t m) {
    uint32_t best = 0xFFFFFFFFu;
    for (uint32_t j = 0; j + m <= seed_nt; j++) {
        const uint32_t v = scrambled_mmer(win, j, m);
        best = v < best ? v : best;
    }
    return best;
}
// The same minimum kept up to date while the window moves by one nucleotide per step (the probe loops of the fast kernels
// walk the overlap lengths that way): one new m-mer per step, a rescan only when the minimum itself leaves the window.
// Not wired into the kernels yet; tests/test_minimizer_cpu.py checks it against window_minimizer_ref.
struct SlidingMinimizer {
    uint32_t best;  // smallest scrambled m-mer of the current window
    int pos;        // position j of (one occurrence of) it, 0 = the m-mer at the low end of the window
    __host__ __device__ void reset(uint64_t win, uint32_t seed_nt, uint32_t m) {
        best = 0xFFFFFFFFu, pos = 0;
        for (uint32_t j = 0; j + m <= seed_nt; j++) {
            const uint32_t v = scrambled_mmer(win, j, m);
            if (v < best) best = v, pos = (int) j;
        }
    }
    // the window now starts one nucleotide further into the read (its low m-mer left, a new one entered at the high end)
    __host__ __device__ uint32_t slide_up(uint64_t win, uint32_t seed_nt, uint32_t m) {
        if (--pos < 0) {
            reset(win, seed_nt, m);
        } else {
            const uint32_t v = scrambled_mmer(win, seed_nt - m, m);
            if (v < best) best = v, pos = (int) (seed_nt - m);
        }
        return best;
    }
    // the window now starts one nucleotide earlier (its high m-mer left, a new one entered at the low end)
    __host__ __device__ uint32_t slide_down(uint64_t win, uint32_t seed_nt, uint32_t m) {
        if (++pos > (int) (seed_nt - m)) {
            reset(win, seed_nt, m);
        } else {
            const uint32_t v = scrambled_mmer(win, 0, m);
            if (v <= best) best = v, pos = 0;
        }
        return best;
    }
};
// bucket of a seed window `win` (seed_nt nucleotides) with hash h
__device__ __forceinline__ uint32_t bucket_index_rt(const SeedTable &t, uint64_t win, uint64_t h, uint32_t seed_nt) {
    if (!t.min_m) return bucket_of(h, t.n_buckets);
    return bucket_of(mix64((uint64_t) window_minimizer(win, seed_nt, t.min_m)), t.n_buckets);
}
// next bucket of a chain (rare path)
__device__ __forceinline__ uint32_t next_bucket(const SeedTable &t, uint32_t bk) {
    const uint32_t nb = bk + 1;
    return nb % t.slice == 0 ? nb - t.slice : nb;
}
// 16-bit tag of a key with hash h: the bit pattern of a positive NORMAL half-precision number (exponent field 1 .. 30), so
// that two tags per word can be compared with one HSET2 (__heq2_mask) -- the integer SIMD compares (__vcmpeq2) are emulated
// with half a dozen instructions on sm_100.  30 720 values; never 0 (0 marks an empty slot).
__device__ __forceinline__ uint32_t tag_of(uint64_t h) {
    const uint32_t x = (uint32_t) (h >> 16) & 0xFFFFu;
    return ((1u + (((x >> 10) * 30u) >> 6)) << 10) | (x & 0x3FFu);
}

// Walk the bucket chain that starts at bucket bk and call f(read_id) for every entry whose tag matches.  General form:
// plain loads, any caller (generic kernels, dictionary of preprocess.cu, chains of the fast kernels).
template <class F>
__device__ __forceinline__ void probe_seed_at(const SeedTable &t, uint64_t h, uint32_t bk, F &&f) {
    const uint32_t tag = tag_of(h);
    while (true) {
        const uint32_t *b = t.slots + (uint64_t) bk * kBucketWords;
        const uint4 q0 = __ldg(reinterpret_cast<const uint4 *>(b));
        const uint32_t cnt = q0.x, n = cnt < (uint32_t) kBucketCap ? cnt : (uint32_t) kBucketCap;
        if (n) {
            const uint4 q1 = __ldg(reinterpret_cast<const uint4 *>(b) + 1), q2 = __ldg(reinterpret_cast<const uint4 *>(b) + 2);
            const uint32_t tw[10] = {q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
            for (int k = 0; k < 10; k++) {
                if ((tw[k] & 0xFFFFu) == tag && (uint32_t) (2 * k) < n) f(__ldg(b + 12 + 2 * k));
                if ((tw[k] >> 16) == tag && (uint32_t) (2 * k + 1) < n) f(__ldg(b + 13 + 2 * k));
            }
        }
        if (cnt <= (uint32_t) kBucketCap) return;
        bk = next_bucket(t, bk);
    }
}
template <class F>
__device__ __forceinline__ void probe_seed(const SeedTable &t, uint64_t h, F &&f) {
    probe_seed_at(t, h, bucket_of(h, t.n_buckets), f);
}
// the same for a seed window `win` of seed_nt nucleotides (honours t.min_m)
template <class F>
__device__ __forceinline__ void probe_seed_window(const SeedTable &t, uint64_t win, uint32_t seed_nt, F &&f) {
    const uint64_t h = mix64(win);
    probe_seed_at(t, h, bucket_index_rt(t, win, h, seed_nt), f);
}

// The table must be zeroed before the first insert.  One atomic (claims a slot) + two stores into the same line.
__device__ __forceinline__ void insert_tag_at(const SeedTable &t, uint32_t tag, uint32_t bk, uint32_t id) {
    while (true) {
        uint32_t *b = t.slots + (uint64_t) bk * kBucketWords;
        const uint32_t pos = atomicAdd(b, 1u);
        if (pos < (uint32_t) kBucketCap) {
            reinterpret_cast<uint16_t *>(b + 2)[pos] = (uint16_t) tag;
            b[12 + pos] = id;
            return;
        }
        bk = next_bucket(t, bk);
    }
}
__device__ __forceinline__ void insert_seed_at(const SeedTable &t, uint64_t h, uint32_t bk, uint32_t id) {
    insert_tag_at(t, tag_of(h), bk, id);
}
__device__ __forceinline__ void insert_seed(const SeedTable &t, uint64_t h, uint32_t id) {
    insert_seed_at(t, h, bucket_of(h, t.n_buckets), id);
}

// Eight consecutive words of a packed read from a 32-byte aligned address: ONE request to the memory system.  On B200 a
// random access that misses L2 costs the same whatever its size up to a sector (39.4 G requests/s, measured:
// scripts/probes/random_requests.cu), and L2 does not merge concurrent misses to one sector -- eight 4-byte loads of a
// sector that is still on its way are eight DRAM reads.  Not allocated in L1 (random data is never reused there).
__device__ __forceinline__ void load8_na(const uint32_t *__restrict__ p, uint32_t (&e)[8]) {
    asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(e[0]), "=r"(e[1]), "=r"(e[2]), "=r"(e[3]), "=r"(e[4]), "=r"(e[5]), "=r"(e[6]), "=r"(e[7])
                 : "l"(p));
}
// words per read slot of the aligned copy the fast kernels work on: a multiple of one sector
__host__ __device__ __forceinline__ uint32_t aligned_stride_words(uint32_t words) { return (words + 7u) & ~7u; }

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kOwnWords = 32;  // staged part of a read in the fast kernels: 512 nucleotides

// Overhang tail of an edge (x -> c, offset o): the last min(o, 32) nucleotides of x[0 .. o), top-aligned in 64
// bits (nucleotide o-1 in bits 62..63).  Phase 2 decides "x[oa-o .. oa) == b[0 .. o)" from two such tails when o <= 32.
__device__ __forceinline__ uint64_t overhang_tail(const uint32_t *__restrict__ p, uint32_t o) {
    if (o == 0) return 0;
    if (o >= 32) return bits64(p, 2u * (o - 32u));
    return bits64(p, 0) << (64u - 2u * o);
}

__device__ __forceinline__ uint64_t sbits64(const uint32_t *own, uint32_t bit) {
    const uint32_t w = bit >> 5, s = bit & 31u;
    const uint32_t a = own[w], b = own[w + 1], c = own[w + 2];
    return (uint64_t) __funnelshift_r(a, b, s) | ((uint64_t) __funnelshift_r(b, c, s) << 32);
}
__device__ __forceinline__ uint64_t overhang_tail_own(const uint32_t *own, uint32_t o) {
    if (o == 0) return 0;
    if (o >= 32) return sbits64(own, 2u * (o - 32u));
    return sbits64(own, 0) << (64u - 2u * o);
}

// Warp-cooperative exact compares: the warp is cut into V groups of gs lanes, group g verifies one candidate, lane
// k of the group compares 32-bit word k.  gs >= number of words of the longest compare of the round.
struct GroupGeom {
    int gs, V, g, k;
};
__device__ __forceinline__ GroupGeom group_geom(int n_words, int lane) {
    GroupGeom q;
    q.gs = n_words <= 8 ? 8 : (n_words <= 10 ? 10 : (n_words <= 16 ? 16 : 32));
    q.V = n_words <= 8 ? 4 : (n_words <= 10 ? 3 : (n_words <= 16 ? 2 : 1));
    q.g = (lane >= q.gs) + (lane >= 2 * q.gs) + (lane >= 3 * q.gs);
    q.k = lane - q.g * q.gs;
    return q;
}
// bit i set <=> group i took part and none of its lanes saw a mismatch
__device__ __forceinline__ unsigned group_ok(const GroupGeom &q, bool active, bool bad) {
    const unsigned badm = __ballot_sync(kFull, active && bad);
    const unsigned actm = __ballot_sync(kFull, active);
    const unsigned gm = q.gs >= 32 ? kFull : ((1u << q.gs) - 1u);
    unsigned ok = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        if (i < q.V) {
            const unsigned m = gm << (i * q.gs);
            if ((actm & m) && !(badm & m)) ok |= 1u << i;
        }
    }
    return ok;
}

}  // namespace alga

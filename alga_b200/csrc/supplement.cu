// Error-rate supplement of the overlap graph (reference: main.cpp:300-355, GraphCreatorLI / GraphCreatorKmerBased /
// GraphCreatorPairwiseKmerBranch, AlignmentControllerHybrid) -- alga_gpu_supplement of include/alga_gpu.h.
//
// The reference walks, per pass, every group of LI k-mers with equal hash and calls canAlign one pair at a time inside
// an order-dependent loop (branch markers + the current neighbours of the source read).  canAlign is a pure function of
// (read a, read b, offset), and so are the filters in front of it, so the work is restructured as
//
//     GPU   LI k-mers of every dead-end read            (Read.cpp:145-226; 70-bit rolling minimum per interval)
//     host  scatter into the 2^20 hash-range buckets, std::sort per bucket (GraphCreatorKmerBased.cpp:94-106,
//           202-259: the tie order inside a bucket is libstdc++'s, exactly as in the reference)
//     GPU   enumeration of every pair that passes the static filters of GraphCreatorPairwiseKmerBranch.cpp:43-62
//           (one thread per k-mer: count, scan, fill) and canAlign of all those pairs in one batch
//           (AlignmentControllerLowErrorRate.cpp:15-49, one pair per warp); the pair list stays on the device
//     host  replay of the ordered loop with the verdicts at hand (:64-84), Graph::addDirectedEdge,
//           retainOnlySmallestOffset
//
// four times, with Read::priorities rotated after each pass (GraphCreatorLI.cpp:18-28).  No CPU fallback: the kernels
// are the only implementation of the k-mer minimum, of the pair filters and of canAlign in this library.
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <utility>
#include <vector>

#include <cub/device/device_radix_sort.cuh>

#include "../../include/alga_gpu.h"
#include "launch.h"

namespace alga {

namespace {

inline int grid_for(uint64_t n_items, int per_block, const LaunchCfg &cfg, int max_blocks_per_sm = 16) {
    uint64_t need = (n_items + per_block - 1) / per_block;
    uint64_t cap = (uint64_t) cfg.sm_count * max_blocks_per_sm;
    if (need < 1) need = 1;
    return (int) (need < cap ? need : cap);
}

typedef unsigned __int128 u128;
constexpr uint64_t kMaxHash = 1000000000000000003ull;  // Params::MAX_HASH_CONSIDERED (Params.cpp:721)

// Read::getLIKmers (Read.cpp:145-226), one warp per read of `ids`.  Window p (0 <= p <= len - K) has the value
// sum_k prio[nt(p + k)] * 4^(K-1-k) (K <= 63: fits 126 bits); interval iv = p / ceil((len - K + 1) / intervals) keeps
// its leftmost minimal window.  Output per read: `intervals` slots of (hash mod 10^18+3, p), p = -1 for intervals
// beyond the last window.
__global__ void __launch_bounds__(256)
li_kmers_kernel(ReadsDev R, const uint32_t *__restrict__ ids, uint32_t n_ids, uint32_t prio_tbl, int K, int intervals,
                uint64_t *__restrict__ hash_out, int32_t *__restrict__ ind_out) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp = (blockIdx.x * (uint64_t) blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t) gridDim.x * blockDim.x) >> 5;
    for (uint64_t q = warp; q < n_ids; q += n_warps) {
        const uint32_t id = ids[q];
        const int len = (int) R.len[id];
        const uint32_t *p = read_ptr(R, id);
        const int n_win = len - K + 1;
        const int ilen = (n_win + intervals - 1) / intervals;
        for (int iv = 0; iv < intervals; iv++) {
            const int w_lo = iv * ilen, w_hi = min(n_win, (iv + 1) * ilen);
            u128 best = ~(u128) 0;
            int best_p = 0x7FFFFFFF;
            for (int w = w_lo + lane; w < w_hi; w += 32) {
                u128 h = 0;
                for (int k = 0; k < K; k++) {
                    const uint32_t j = (uint32_t) (w + k);
                    const uint32_t code = (__ldg(p + (j >> 4)) >> ((j & 15u) * 2u)) & 3u;
                    h = (h << 2) + (u128) ((prio_tbl >> (2u * code)) & 3u);
                }
                if (h < best) {  // strict: within one lane the windows come in ascending order
                    best = h;
                    best_p = w;
                }
            }
            for (int d = 16; d; d >>= 1) {
                const uint64_t oh = __shfl_xor_sync(kFull, (uint64_t) (best >> 64), d);
                const uint64_t ol = __shfl_xor_sync(kFull, (uint64_t) best, d);
                const int op = __shfl_xor_sync(kFull, best_p, d);
                const u128 o = ((u128) oh << 64) | (u128) ol;
                if (o < best || (o == best && op < best_p)) {
                    best = o;
                    best_p = op;
                }
            }
            if (lane == 0) {
                const uint64_t slot = q * (uint64_t) intervals + (uint64_t) iv;
                if (w_lo < w_hi) {
                    hash_out[slot] = (uint64_t) (best % (u128) kMaxHash);
                    ind_out[slot] = best_p;
                } else {
                    hash_out[slot] = 0;
                    ind_out[slot] = -1;
                }
            }
        }
    }
}

thread_local char g_sup_err[512] = "";
int sup_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_sup_err, sizeof(g_sup_err), fmt, ap);
    va_end(ap);
    return code;
}
#define SCK(call)                                                                                              \
    do {                                                                                                       \
        cudaError_t e_ = (call);                                                                               \
        if (e_ != cudaSuccess)                                                                                 \
            return sup_fail(e_ == cudaErrorMemoryAllocation ? ALGA_E_NOMEM : ALGA_E_CUDA, "%s failed: %s (%s:%d)", \
                            #call, cudaGetErrorString(e_), __FILE__, __LINE__);                                \
    } while (0)

struct Buf {  // device buffer freed on scope exit; alloc() only ever grows it (a cudaFree + cudaMalloc per pass cost up to
              // 0.5 s on the B200 box)
    void *p = nullptr;
    size_t cap = 0;
    ~Buf() {
        if (p) cudaFree(p);
    }
    int alloc(size_t bytes) {
        if (p && bytes <= cap) return ALGA_OK;
        if (p) cudaFree(p);
        p = nullptr, cap = 0;
        const size_t want = (bytes ? bytes : 16) + bytes / 8;
        SCK(cudaMalloc(&p, want));
        cap = want;
        return ALGA_OK;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct Kmer {
    uint32_t read;
    uint64_t hash;
    int ind;
    uint32_t read_len;
    bool operator<(const Kmer &o) const {  // Kmer.cpp:58-64
        if (hash != o.hash) return hash < o.hash;
        if (ind != o.ind) return ind > o.ind;
        if (read_len != o.read_len) return read_len < o.read_len;
        return false;
    }
};

// ---- the k-mers of a pass in the reference's final order, on the device --------------------------------------------------
// The reference scatters the k-mers into 2^20 hash-range buckets in read / interval order and runs std::sort on every bucket
// (GraphCreatorKmerBased.cpp:94-106, 139-179, comparator Kmer.cpp:58-64: hash up, position down, read length up).  The bucket
// number is monotone in the hash, so the order of all k-mers is the order by (hash, position desc, length) -- a stable radix
// sort -- EXCEPT among k-mers whose three keys tie: where those end up is a property of libstdc++'s introsort applied to the
// bucket in its fill order.  So: sort on the device, find the buckets that hold a fully tied pair (a few hundred of 10^5
// groups), and re-sort just those on the host exactly as the reference does (fill order = slot order).
__host__ __device__ __forceinline__ uint32_t kmer_bucket(uint64_t hash) {  // GraphCreatorKmerBased.cpp:233
    return (uint32_t) (int) ((1048576ll - 1) * ((double) hash / (double) kMaxHash));
}
__global__ void kmer_flags_kernel(const int32_t *__restrict__ ind, uint32_t n_slots, uint32_t *__restrict__ flag) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t) gridDim.x * blockDim.x)
        flag[i] = ind[i] >= 0 ? 1u : 0u;
}
__global__ void kmer_fill_kernel(const uint32_t *__restrict__ ids, const uint64_t *__restrict__ hash, const int32_t *__restrict__ ind,
                                 const uint32_t *__restrict__ len, uint32_t IV, uint32_t n_slots, const uint32_t *__restrict__ pos,
                                 Kmer *__restrict__ A, uint64_t *__restrict__ key_lo, uint64_t *__restrict__ key_hi,
                                 uint32_t *__restrict__ idx) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n_slots; i += (uint64_t) gridDim.x * blockDim.x) {
        if (ind[i] < 0) continue;
        const uint32_t p = pos[i], id = ids[i / IV];
        const Kmer k{id, hash[i], ind[i], len[id]};
        A[p] = k;
        key_lo[p] = ((uint64_t) (0xFFFFFFFFu - (uint32_t) k.ind) << 32) | k.read_len;  // position down, read length up
        key_hi[p] = k.hash;
        idx[p] = p;
    }
}
__global__ void gather_u64_kernel(const uint64_t *__restrict__ src, const uint32_t *__restrict__ idx, uint32_t n, uint64_t *__restrict__ dst) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) dst[i] = src[idx[i]];
}
// B = A in sorted order; buckets that hold two k-mers with equal (hash, position, length) go to `tied` (with repetitions)
__global__ void kmer_gather_kernel(const Kmer *__restrict__ A, const uint32_t *__restrict__ idx, uint32_t n, Kmer *__restrict__ B,
                                   uint32_t *__restrict__ tied, uint32_t *n_tied, uint32_t tied_cap) {
    for (uint64_t i = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; i < n; i += (uint64_t) gridDim.x * blockDim.x) {
        const Kmer k = A[idx[i]];
        B[i] = k;
        if (i > 0) {
            const Kmer q = A[idx[i - 1]];
            if (q.hash == k.hash && q.ind == k.ind && q.read_len == k.read_len) {
                const uint32_t t = atomicAdd(n_tied, 1u);
                if (t < tied_cap) tied[t] = kmer_bucket(k.hash);
            }
        }
    }
}

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct Pinned {  // page-locked host buffer freed on scope exit
    void *p = nullptr;
    size_t cap = 0;
    ~Pinned() {
        if (p) cudaFreeHost(p);
    }
    int ensure(size_t bytes) {
        if (bytes <= cap && p) return ALGA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr, cap = 0;
        const size_t want = bytes + bytes / 4 + 64;
        SCK(cudaMallocHost(&p, want));
        cap = want;
        return ALGA_OK;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

// The static filters of GraphCreatorPairwiseKmerBranch.cpp:43-62 for the ordered pair (i, j) of one group:
// 0 = candidate, 1 = skip j, 2 = stop this i.
struct PairFilter {
    int min_offset, max_offset_pct, min_overlap_area;
};
__host__ __device__ inline int pair_filter_of(const Kmer &ki, const Kmer &kj, const PairFilter &F, int &offset) {
    if (ki.read == kj.read) return 1;
    offset = ki.ind - kj.ind;
    if (offset < F.min_offset) return 1;
    if (100ll * offset > (long long) F.max_offset_pct * (long long) ki.read_len) return 2;
    const int a = (int) ki.read_len, b = (int) kj.read_len + offset;
    const int overlap = (a < b ? a : b) - offset;
    if (overlap < F.min_overlap_area) return 1;
    if ((int) kj.read_len + offset - (int) ki.read_len < 0) return 1;  // Read::getRightOffset
    return 0;
}

// Pair enumeration on the sorted k-mer array (one thread per k-mer x, which plays i; j runs over the rest of its group,
// i.e. while the hash stays the same): first the number of pairs that pass the filters, then -- at the offsets a scan
// made of the counts -- the pairs themselves as (read_i, read_j, offset) for canAlign plus the position of j.
template <bool FILL>
__global__ void enumerate_pairs_kernel(const Kmer *__restrict__ km, uint32_t nk, PairFilter F, uint32_t *__restrict__ cnt,
                                       const uint64_t *__restrict__ off, int32_t *__restrict__ pairs, uint32_t *__restrict__ pj) {
    for (uint64_t x = blockIdx.x * (uint64_t) blockDim.x + threadIdx.x; x < nk; x += (uint64_t) gridDim.x * blockDim.x) {
        const Kmer ki = km[x];
        uint64_t w = FILL ? off[x] : 0;
        uint32_t c = 0;
        for (uint64_t j = x + 1; j < nk; j++) {
            const Kmer kj = km[j];
            if (kj.hash != ki.hash) break;
            int offset = 0;
            const int f = pair_filter_of(ki, kj, F, offset);
            if (f == 2) break;
            if (f == 1) continue;
            if (FILL) {
                pairs[3 * w] = (int32_t) ki.read;
                pairs[3 * w + 1] = (int32_t) kj.read;
                pairs[3 * w + 2] = offset;
                pj[w] = (uint32_t) j;
                w++;
            }
            c++;
        }
        if (!FILL) cnt[x] = c;
    }
}

}  // namespace

const char *supplement_last_error() { return g_sup_err; }

// Device-side pieces exposed to api.cu (alga_gpu_li_kmers) and used below.
int run_li_kmers(const ReadsDev &R, const uint32_t *d_ids, uint32_t n_ids, const int32_t prio[4], int K, int intervals,
                 uint64_t *d_hash, int32_t *d_ind, cudaStream_t s, const LaunchCfg &cfg) {
    if (!n_ids) return ALGA_OK;
    uint32_t tbl = 0;
    for (int c = 0; c < 4; c++) tbl |= ((uint32_t) prio[c] & 3u) << (2 * c);
    li_kmers_kernel<<<grid_for(n_ids, 8, cfg, 8), 256, 0, s>>>(R, d_ids, n_ids, tbl, K, intervals, d_hash, d_ind);
    if (cfg.launches) (*cfg.launches)++;
    SCK(cudaGetLastError());
    return ALGA_OK;
}

int supplement_impl(const alga_reads *h, const alga_csr *gin, const alga_sup_params *sp, alga_csr *gout, alga_timing *tm) {
    const uint32_t n = h->n_reads;
    if (gin->n_reads != n) return sup_fail(ALGA_E_INVALID, "graph has %u rows, read set %u reads", gin->n_reads, n);
    if (sp->kmer_length < 1 || sp->kmer_length > 63 || sp->intervals < 1 || sp->intervals > 64)
        return sup_fail(ALGA_E_INVALID, "kmer_length must be 1..63 and intervals 1..64");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return sup_fail(ALGA_E_CUDA, "no CUDA device available; libalga_gpu has no CPU fallback");
    }
    if (sp->device < 0 || sp->device >= ndev) return sup_fail(ALGA_E_INVALID, "device %d out of range", sp->device);
    SCK(cudaSetDevice(sp->device));
    LaunchCfg cfg;
    uint64_t launches = 0;
    cfg.launches = &launches;
    SCK(cudaDeviceGetAttribute(&cfg.sm_count, cudaDevAttrMultiProcessorCount, sp->device));
    const double t0 = now_ms();

    // ---- reads to the device (once for all passes)
    const uint64_t n_words = h->word_off ? h->word_off[n] : (uint64_t) n * h->stride_words;
    Buf d_words, d_off, d_len;
    if (d_words.alloc((size_t) n_words * 4 + kReadPadBytes) || d_len.alloc((size_t) (n ? n : 1) * 4)) return ALGA_E_NOMEM;
    SCK(cudaMemcpy(d_words.p, h->words, (size_t) n_words * 4, cudaMemcpyHostToDevice));
    SCK(cudaMemset((char *) d_words.p + (size_t) n_words * 4, 0, kReadPadBytes));
    SCK(cudaMemcpy(d_len.p, h->len_nt, (size_t) n * 4, cudaMemcpyHostToDevice));
    ReadsDev R{};
    R.words = d_words.as<uint32_t>();
    R.len = d_len.as<uint32_t>();
    R.n = n;
    R.stride = h->word_off ? 0 : h->stride_words;
    if (h->word_off) {
        if (d_off.alloc(((size_t) n + 1) * 8)) return ALGA_E_NOMEM;
        SCK(cudaMemcpy(d_off.p, h->word_off, ((size_t) n + 1) * 8, cudaMemcpyHostToDevice));
        R.word_off = d_off.as<uint64_t>();
    }
    const double t1 = now_ms();

    // ---- graph rows + the reads that take part: dead ends only, fixed before the first pass (main.cpp:308-323)
    std::vector<std::vector<std::pair<int, int>>> V(n);
    std::vector<int> indeg(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        V[i].reserve((size_t) (gin->row_off[i + 1] - gin->row_off[i]) + 2);
        for (uint64_t k = gin->row_off[i]; k < gin->row_off[i + 1]; k++) {
            V[i].push_back({gin->nbr[k], gin->off[k]});
            indeg[(size_t) gin->nbr[k]]++;
        }
    }
    std::vector<uint32_t> ids;
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t len = h->len_nt[i];
        const bool dead_end = (indeg[i] == 0 && !V[i].empty()) || (indeg[i] > 0 && V[i].empty());
        // Read::getKmers (Read.cpp:70-72) returns nothing for reads shorter than KMER_LENGTH_BUCKET; getLIKmers needs
        // len >= K (the reference exits otherwise; main.cpp:253-266 removed such reads before)
        if (len == 0 || !dead_end || (int64_t) len < sp->kmer_length_bucket || (int64_t) len < sp->kmer_length) continue;
        ids.push_back(i);
    }
    const uint32_t n_ids = (uint32_t) ids.size();
    const int IV = sp->intervals;
    Buf d_ids, d_hash, d_ind, d_pairs, d_verdict;
    size_t pairs_cap = 0;
    if (d_ids.alloc((size_t) n_ids * 4) || d_hash.alloc((size_t) n_ids * IV * 8) || d_ind.alloc((size_t) n_ids * IV * 4))
        return ALGA_E_NOMEM;
    if (n_ids) SCK(cudaMemcpy(d_ids.p, ids.data(), (size_t) n_ids * 4, cudaMemcpyHostToDevice));
    VerifyDev vd{sp->max_offset_pct, sp->min_offset, sp->min_overlap_area, sp->threshold_pct, sp->same_ends};

    // Host side, flat: all k-mers of a pass in ONE array ordered by bucket (counting sort that keeps the reference's fill
    // order inside a bucket: reads by id, k-mers by interval), std::sort on each bucket's range -- the same sequence,
    // comparator and algorithm as std::sort on the reference's per-bucket vectors, hence the same tie order.
    Pinned km_buf;                                            // k-mers of the pass, bucket-major (page-locked: it is uploaded)
    const int INF = 1000000001;
    const unsigned n_thr = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    // pair enumeration on the device: per k-mer x (as i) the range [pair_off[x], pair_off[x + 1]) of its pairs
    Buf d_km, d_cnt, d_poff, d_pj, scan_ws;
    Buf d_flag, d_pos, d_sort;  // the sort of the k-mers (d_sort: one arena)
    constexpr uint32_t kTiedCap = 1u << 16;                                    // tied pairs listed per pass (more: found on the host)
    uint64_t n_tied_buckets = 0;
    Pinned h_poff, h_pj, h_verdict;
    std::vector<uint64_t> bm;                                  // branch markers of one group: D rows of ceil(D/64) words
    struct Group {
        uint32_t p, q, level;  // k-mers [p, q) of equal hash; see the replay below
    };
    std::vector<Group> grp, by_level;
    std::vector<size_t> level_start;
    std::vector<uint32_t> last_level(n ? n : 1, 0u);           // per read: level of the last group it was a source in
    std::vector<std::vector<uint64_t>> tbm(n_thr);             // branch markers per worker thread
    const bool serial_replay = getenv("ALGA_SUP_SERIAL") != nullptr;  // A/B switch: groups strictly one after the other
    const bool trace = getenv("ALGA_SUP_TRACE") != nullptr;           // per-pass timing of the enumeration on stderr
    uint64_t levels_total = 0;
    double gpu_ms = 0, t_kmers = 0, t_sort = 0, t_enum = 0, t_verify = 0, t_replay = 0;
    uint64_t pairs_total = 0;
    int32_t prio[4] = {0, 1, 2, 3};

    const PairFilter pf_params{sp->min_offset, sp->max_offset_pct, sp->min_overlap_area};
    auto run_threads = [&](auto &&fn) {  // fn(t): bucket range [t, t+1) * kBucketsSort / n_thr
        std::vector<std::thread> th;
        for (unsigned t = 1; t < n_thr; t++) th.emplace_back([&fn, t] { fn(t); });
        fn(0u);
        for (auto &x : th) x.join();
    };

    for (int pass = 0; pass < 4; pass++) {  // GraphCreatorLI.cpp:20-26
        // ---- LI k-mers on the GPU
        const double ta = now_ms();
        if (int rc = run_li_kmers(R, d_ids.as<uint32_t>(), n_ids, prio, sp->kmer_length, IV, d_hash.as<uint64_t>(),
                                  d_ind.as<int32_t>(), 0, cfg))
            return rc;
        SCK(cudaDeviceSynchronize());
        gpu_ms += now_ms() - ta;
        t_kmers += now_ms() - ta;
        // ---- the k-mers in the reference's order: stable radix sort on the device, host re-sort of the buckets with tied keys
        const double tc = now_ms();
        const size_t n_slots = (size_t) n_ids * IV;
        size_t nk = 0;
        Kmer *km = nullptr;
        if (n_slots) {
            if (d_flag.alloc(n_slots * 4) || d_pos.alloc((n_slots + 1) * 4) || scan_ws.alloc(scan_workspace_bytes(n_slots))) return ALGA_E_NOMEM;
            kmer_flags_kernel<<<grid_for(n_slots, 256, cfg, 8), 256>>>(d_ind.as<int32_t>(), (uint32_t) n_slots, d_flag.as<uint32_t>());
            launch_scan_u32(d_flag.as<uint32_t>(), d_pos.as<uint32_t>(), n_slots, scan_ws.p, 0, cfg);
            uint32_t nk32 = 0;
            SCK(cudaMemcpy(&nk32, d_pos.as<uint32_t>() + n_slots, 4, cudaMemcpyDeviceToHost));
            nk = nk32;
            launches += 4;
        }
        if (km_buf.ensure((nk ? nk : 1) * sizeof(Kmer))) return ALGA_E_NOMEM;
        km = km_buf.as<Kmer>();
        if (nk) {
            // one arena for the sort's buffers (every cudaMalloc of this size costs a fraction of a second on the box)
            size_t tmp_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, (const uint64_t *) nullptr, (uint64_t *) nullptr, (const uint32_t *) nullptr,
                                            (uint32_t *) nullptr, (int) nk);
            auto up = [](size_t b) { return (b + 255) & ~(size_t) 255; };
            const size_t o_kmA = 0, o_k0 = o_kmA + up(nk * sizeof(Kmer)), o_k1 = o_k0 + up(nk * 8), o_k2 = o_k1 + up(nk * 8),
                         o_i0 = o_k2 + up(nk * 8), o_i1 = o_i0 + up(nk * 4), o_tied = o_i1 + up(nk * 4),
                         o_cub = o_tied + up(((size_t) kTiedCap + 1) * 4), o_end = o_cub + up(tmp_bytes + 16);
            if (d_sort.alloc(o_end) || d_km.alloc(nk * sizeof(Kmer))) return ALGA_E_NOMEM;
            char *const sb = d_sort.as<char>();
            Kmer *const p_kmA = (Kmer *) (sb + o_kmA);
            uint64_t *const p_k0 = (uint64_t *) (sb + o_k0), *const p_k1 = (uint64_t *) (sb + o_k1), *const p_k2 = (uint64_t *) (sb + o_k2);
            uint32_t *const p_i0 = (uint32_t *) (sb + o_i0), *const p_i1 = (uint32_t *) (sb + o_i1), *const p_tied = (uint32_t *) (sb + o_tied);
            void *const p_cub = sb + o_cub;
            const double ts0 = now_ms();
            kmer_fill_kernel<<<grid_for(n_slots, 256, cfg, 8), 256>>>(d_ids.as<uint32_t>(), d_hash.as<uint64_t>(), d_ind.as<int32_t>(), R.len,
                                                                      (uint32_t) IV, (uint32_t) n_slots, d_pos.as<uint32_t>(), p_kmA, p_k0, p_k1, p_i0);
            // least significant key first: (position desc, length asc), then the hash; both sorts are stable
            SCK(cub::DeviceRadixSort::SortPairs(p_cub, tmp_bytes, (const uint64_t *) p_k0, p_k2, (const uint32_t *) p_i0, p_i1, (int) nk));
            gather_u64_kernel<<<grid_for(nk, 256, cfg, 8), 256>>>(p_k1, p_i1, (uint32_t) nk, p_k0);
            SCK(cub::DeviceRadixSort::SortPairs(p_cub, tmp_bytes, (const uint64_t *) p_k0, p_k2, (const uint32_t *) p_i1, p_i0, (int) nk));
            SCK(cudaMemset(p_tied + kTiedCap, 0, 4));
            kmer_gather_kernel<<<grid_for(nk, 256, cfg, 8), 256>>>(p_kmA, p_i0, (uint32_t) nk, d_km.as<Kmer>(), p_tied, p_tied + kTiedCap, kTiedCap);
            SCK(cudaGetLastError());
            launches += 5;
            if (trace) SCK(cudaDeviceSynchronize());
            const double ts1 = now_ms();
            SCK(cudaMemcpy(km, d_km.p, nk * sizeof(Kmer), cudaMemcpyDeviceToHost));
            const double ts2 = now_ms();
            uint32_t n_tied = 0;
            SCK(cudaMemcpy(&n_tied, p_tied + kTiedCap, 4, cudaMemcpyDeviceToHost));
            std::vector<uint32_t> tied(std::min<uint32_t>(n_tied, kTiedCap));
            if (!tied.empty()) SCK(cudaMemcpy(tied.data(), p_tied, tied.size() * 4, cudaMemcpyDeviceToHost));
            std::vector<uint32_t> rank;  // fill-order rank (= position in the unsorted array) of every k-mer, fetched only if needed
            if (n_tied > kTiedCap) {     // more tied buckets than the list holds: every bucket is suspect
                tied.clear();
                for (size_t x = 1; x < nk; x++)
                    if (km[x].hash == km[x - 1].hash && km[x].ind == km[x - 1].ind && km[x].read_len == km[x - 1].read_len)
                        tied.push_back(kmer_bucket(km[x].hash));
            }
            std::sort(tied.begin(), tied.end());
            tied.erase(std::unique(tied.begin(), tied.end()), tied.end());
            if (!tied.empty()) {
                rank.resize(nk);
                SCK(cudaMemcpy(rank.data(), p_i0, nk * 4, cudaMemcpyDeviceToHost));
                // (with sequencing errors reads that start at the same position are no duplicates any more but still share
                // their k-mers: a quarter of the buckets of config 3 hold a tie -- the buckets are small, the threads share them)
                const Kmer *km_c = km;
                run_threads([&](unsigned t) {
                    std::vector<std::pair<uint32_t, Kmer>> tmp;
                    const size_t z0 = tied.size() * t / n_thr, z1 = tied.size() * (t + 1) / n_thr;
                    for (size_t z = z0; z < z1; z++) {
                        const uint32_t bkt = tied[z];
                        // the bucket's range in the sorted array (the bucket number is monotone in the hash)
                        size_t lo = 0, hi = nk;
                        while (lo < hi) {
                            const size_t mid = (lo + hi) / 2;
                            if (kmer_bucket(km_c[mid].hash) < bkt) lo = mid + 1;
                            else hi = mid;
                        }
                        size_t e = lo;
                        while (e < nk && kmer_bucket(km_c[e].hash) == bkt) e++;
                        tmp.clear();
                        for (size_t x = lo; x < e; x++) tmp.emplace_back(rank[x], km_c[x]);
                        std::sort(tmp.begin(), tmp.end(),
                                  [](const std::pair<uint32_t, Kmer> &a, const std::pair<uint32_t, Kmer> &b) { return a.first < b.first; });
                        for (size_t x = lo; x < e; x++) km[x] = tmp[x - lo].second;  // the bucket as the reference fills it
                        std::sort(km + lo, km + e);                                   // ... and sorts it (GraphCreatorKmerBased.cpp:99)
                    }
                });
                SCK(cudaMemcpy(d_km.p, km, nk * sizeof(Kmer), cudaMemcpyHostToDevice));  // the re-sorted buckets, in one go
            }
            n_tied_buckets += tied.size();
            if (trace)
                fprintf(stderr, "alga_gpu supplement pass %d sort: alloc %.1f ms, device sort %.1f ms, read-back %.1f ms, %u tied (%zu buckets) fixed in %.1f ms\n",
                        pass, ts0 - tc, ts1 - ts0, ts2 - ts1, n_tied, tied.size(), now_ms() - ts2);
        }
        gpu_ms += now_ms() - tc;
        t_sort += now_ms() - tc;
        // ---- every pair that passes the static filters (:43-62), on the device: count per k-mer, scan, fill
        const double td = now_ms();
        uint64_t n_pairs = 0;
        uint64_t *pair_off = nullptr;
        if (nk) {
            if (d_km.alloc(nk * sizeof(Kmer)) || d_cnt.alloc(nk * 4) || d_poff.alloc((nk + 1) * 8) ||
                scan_ws.alloc(scan_workspace_bytes(nk)) || h_poff.ensure((nk + 1) * 8))
                return ALGA_E_NOMEM;
            const double tq0 = now_ms();
            // (the sorted k-mers are on the device already)
            const double tq1 = now_ms();
            const int grid = grid_for(nk, 128, cfg, 16);
            enumerate_pairs_kernel<false><<<grid, 128>>>(d_km.as<Kmer>(), (uint32_t) nk, pf_params, d_cnt.as<uint32_t>(), nullptr,
                                                         nullptr, nullptr);
            if (trace) SCK(cudaDeviceSynchronize());
            const double tq2 = now_ms();
            launch_scan_u64(d_cnt.as<uint32_t>(), d_poff.as<uint64_t>(), nk, scan_ws.p, 0, cfg);
            SCK(cudaGetLastError());
            SCK(cudaMemcpy(h_poff.p, d_poff.p, (nk + 1) * 8, cudaMemcpyDeviceToHost));
            if (trace)
                fprintf(stderr, "alga_gpu supplement pass %d: nk %zu, alloc %.1f ms, upload %.1f ms, count kernel %.1f ms, scan + read-back %.1f ms\n",
                        pass, nk, tq0 - td, tq1 - tq0, tq2 - tq1, now_ms() - tq2);
            pair_off = h_poff.as<uint64_t>();
            n_pairs = pair_off[nk];
            launches += 1;
        }
        pairs_total += n_pairs;
        t_enum += now_ms() - td;
        // ---- the pairs themselves and canAlign of all of them in one batch, still on the device
        const double tb = now_ms();
        if (n_pairs) {
            if (n_pairs > pairs_cap) {
                pairs_cap = (size_t) n_pairs + n_pairs / 4;
                if (d_pairs.alloc(pairs_cap * 12) || d_verdict.alloc(pairs_cap) || d_pj.alloc(pairs_cap * 4)) return ALGA_E_NOMEM;
            }
            if (h_pj.ensure((size_t) n_pairs * 4) || h_verdict.ensure((size_t) n_pairs)) return ALGA_E_NOMEM;
            enumerate_pairs_kernel<true><<<grid_for(nk, 128, cfg, 16), 128>>>(d_km.as<Kmer>(), (uint32_t) nk, pf_params, nullptr,
                                                                              d_poff.as<uint64_t>(), d_pairs.as<int32_t>(),
                                                                              d_pj.as<uint32_t>());
            launches += 1;
            launch_verify_pairs(R, d_pairs.as<int32_t>(), n_pairs, vd, d_verdict.as<uint8_t>(), 0, cfg);
            SCK(cudaGetLastError());
            SCK(cudaMemcpy(h_pj.p, d_pj.p, (size_t) n_pairs * 4, cudaMemcpyDeviceToHost));
            SCK(cudaMemcpy(h_verdict.p, d_verdict.p, (size_t) n_pairs, cudaMemcpyDeviceToHost));
        }
        const uint32_t *pair_j = h_pj.as<uint32_t>();
        const uint8_t *verdict = h_verdict.as<uint8_t>();
        gpu_ms += now_ms() - td;
        t_verify += now_ms() - tb;
        // ---- replay of the ordered loop (:64-84) with the verdicts at hand; sequential: groups share graph rows
        const double te = now_ms();
        // The pairs that passed the static filters were listed per k-mer (as i) by the enumeration above, so a group only
        // visits those; the pair's offset and second read follow from the two k-mers.
        //
        // A group reads and writes nothing but the rows of its own source reads (the k-mers i that have pairs) and its own
        // branch markers, so two groups depend on each other only if they share a source read, and then their order matters.
        // Levels: level(g) = 1 + the highest level of an earlier group that shares a source read with g.  Groups of one
        // level are independent and run in parallel; the levels run one after the other -- every row sees its groups in
        // exactly the order of the sequential loop.
        auto process_group = [&](size_t p, size_t q, std::vector<uint64_t> &marks) {
            const int D = (int) (q - p);
            const int RW = (D + 63) >> 6;  // words per branch-marker row
            marks.assign((size_t) D * RW, 0ull);
            for (int i = D - 2; i >= 0; i--) {
                const size_t x = p + (size_t) i;
                const uint64_t c0 = pair_off[x], c1 = pair_off[x + 1];
                if (c0 == c1) continue;
                const Kmer &ki = km[x];
                const int id1 = (int) ki.read;
                auto &row = V[(size_t) id1];
                // the reference spreads the row into a dense `neighbors` array (:64-66); rows hold one entry per target
                // and stay short, so the entry is looked up in the row itself -- no random access per pair
                uint64_t *bi = marks.data() + (size_t) i * RW;
                for (uint64_t c = c0; c < c1; c++) {
                    const Kmer &kj = km[pair_j[c]];
                    const int j = (int) (pair_j[c] - p);
                    const int id2 = (int) kj.read, offset = ki.ind - kj.ind;
                    const uint8_t can = verdict[c];
                    if (!((bi[j >> 6] >> (j & 63)) & 1ull)) {
                        int *cur = nullptr;  // offset of the edge id1 -> id2, if there is one (= neighbors[id2])
                        for (auto &e : row) {
                            if (e.first == id2) {
                                cur = &e.second;
                                break;
                            }
                        }
                        int cur_off = cur ? *cur : INF;
                        if (cur_off > offset && can) {
                            // Graph::addDirectedEdge (Graph.cpp:53-71): one entry per target, smallest offset
                            if (cur) *cur = offset;
                            else row.push_back({id2, offset});
                            cur_off = offset;
                        }
                        if (cur_off != INF) {
                            bi[j >> 6] |= 1ull << (j & 63);
                            const uint64_t *bj = marks.data() + (size_t) j * RW;
                            for (int t = 0; t < RW; t++) bi[t] |= bj[t];
                        }
                    }
                }
            }
        };
        // groups that have pairs at all, in loop order, with their levels
        grp.clear();
        std::fill(last_level.begin(), last_level.end(), 0u);
        uint32_t n_levels = 0;
        if (n_pairs) {  // groups = runs of equal hash (a run never spans two of the reference's buckets: the bucket follows from the hash)
            size_t p = 0, q = 0;
            const size_t end = nk;
            while (p < end) {
                while (q < end && km[q].hash == km[p].hash) q++;
                if (q - p > 1 && pair_off[q - 1] > pair_off[p]) {  // the last k-mer of a group never plays i
                    uint32_t lvl = 0;
                    for (size_t x = p; x + 1 < q; x++)
                        if (pair_off[x + 1] > pair_off[x]) lvl = std::max(lvl, last_level[(size_t) km[x].read]);
                    lvl++;
                    for (size_t x = p; x + 1 < q; x++)
                        if (pair_off[x + 1] > pair_off[x]) last_level[(size_t) km[x].read] = lvl;
                    grp.push_back(Group{(uint32_t) p, (uint32_t) q, lvl});
                    n_levels = std::max(n_levels, lvl);
                }
                p = q;
            }
        }
        if (serial_replay) {
            for (const Group &g : grp) process_group(g.p, g.q, bm);
        } else {
            // counting sort by level (stable: loop order inside a level, which only matters for locality)
            level_start.assign((size_t) n_levels + 2, 0);
            for (const Group &g : grp) level_start[(size_t) g.level + 1]++;
            for (size_t l = 0; l + 1 < level_start.size(); l++) level_start[l + 1] += level_start[l];
            by_level.resize(grp.size());
            {
                std::vector<size_t> cur(level_start.begin(), level_start.end() - 1);
                for (const Group &g : grp) by_level[cur[g.level]++] = g;
            }
            for (uint32_t l = 1; l <= n_levels; l++) {
                const size_t a = level_start[l], b = level_start[(size_t) l + 1];
                if (b - a < 4096) {  // not worth waking the threads
                    for (size_t z = a; z < b; z++) process_group(by_level[z].p, by_level[z].q, bm);
                    continue;
                }
                run_threads([&](unsigned t) {
                    std::vector<uint64_t> &marks = tbm[t];
                    const size_t z0 = a + (b - a) * t / n_thr, z1 = a + (b - a) * (t + 1) / n_thr;
                    for (size_t z = z0; z < z1; z++) {
                        if (z + 1 < z1)  // rows are scattered over the heap: fetch the next group's row headers early
                            for (size_t x = by_level[z + 1].p; x < by_level[z + 1].q; x++) __builtin_prefetch(&V[(size_t) km[x].read]);
                        process_group(by_level[z].p, by_level[z].q, marks);
                    }
                });
            }
        }
        levels_total += n_levels;
        // ---- retainOnlySmallestOffset (GraphCreatorKmerBased.cpp:87; main.cpp:346 after the last pass)
        run_threads([&](unsigned t) {
            const size_t r0 = (size_t) n * t / n_thr, r1 = (size_t) n * (t + 1) / n_thr;
            for (size_t r = r0; r < r1; r++) {
                auto &row = V[r];
                if (row.size() < 2) continue;
                std::sort(row.begin(), row.end());
                size_t w = 0;
                for (size_t k = 0; k < row.size(); k++)
                    if (w == 0 || row[w - 1].first != row[k].first) row[w++] = row[k];
                row.resize(w);
            }
        });
        std::rotate(prio, prio + 1, prio + 4);
        t_replay += now_ms() - te;
    }

    // ---- result
    uint64_t E = 0;
    for (auto &row : V) E += row.size();
    memset(gout, 0, sizeof(*gout));
    gout->n_reads = n;
    gout->n_edges = E;
    gout->row_off = (uint64_t *) malloc(((size_t) n + 1) * 8);
    gout->nbr = (int32_t *) malloc((size_t) (E ? E : 1) * 4);
    gout->off = (int32_t *) malloc((size_t) (E ? E : 1) * 4);
    if (!gout->row_off || !gout->nbr || !gout->off) {
        free(gout->row_off);
        free(gout->nbr);
        free(gout->off);
        memset(gout, 0, sizeof(*gout));
        return sup_fail(ALGA_E_NOMEM, "out of host memory for %llu edges", (unsigned long long) E);
    }
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; i++) {
        gout->row_off[i] = w;
        for (auto &e : V[i]) {
            gout->nbr[w] = e.first;
            gout->off[w] = e.second;
            w++;
        }
    }
    gout->row_off[n] = w;
    if (tm) {
        memset(tm, 0, sizeof(*tm));
        tm->h2d_ms = t1 - t0;
        tm->device_ms = gpu_ms;  // k-mer + canAlign kernels incl. their transfers
        tm->total_ms = now_ms() - t0;
        tm->kernel_launches = launches;
        tm->stage_ms[0] = t_kmers;   // LI k-mers (GPU + D2H)
        tm->stage_ms[1] = t_sort;    // bucket scatter + std::sort (host)
        tm->stage_ms[2] = t_enum;    // pair enumeration (host)
        tm->stage_ms[3] = t_verify;  // canAlign batch (H2D + GPU + D2H)
        tm->stage_ms[4] = t_replay;  // ordered replay + row dedupe (host)
        tm->stage_ms[5] = (double) n_ids;        // diagnostics: dead-end reads that took part,
        tm->stage_ms[6] = (double) pairs_total;  // pairs verified on the GPU over the four passes
        tm->stage_ms[7] = (double) levels_total; // dependency levels of the replay over the four passes
    }
    return ALGA_OK;
}

}  // namespace alga

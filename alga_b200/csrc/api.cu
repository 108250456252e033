// C ABI of libalga_gpu.so (include/alga_gpu.h): plan / workspace management and the host-side
// orchestration of the overlap-graph pipeline.  No CPU fallback: every compute entry point needs a
// CUDA device and fails with ALGA_E_CUDA otherwise.
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>

#include "../../include/alga_gpu.h"
#include "launch.h"

using namespace alga;

namespace alga {
int supplement_impl(const alga_reads *h, const alga_csr *gin, const alga_sup_params *sp, alga_csr *gout, alga_timing *tm);
}

namespace {

thread_local char g_err[512] = "";

int fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess)                                                                           \
            return fail(e_ == cudaErrorMemoryAllocation ? ALGA_E_NOMEM : ALGA_E_CUDA, "%s failed: %s (%s:%d)", #call, \
                        cudaGetErrorString(e_), __FILE__, __LINE__);                                     \
    } while (0)
#define CKR(expr)              \
    do {                       \
        int r_ = (expr);       \
        if (r_ != ALGA_OK) return r_; \
    } while (0)

struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    // Every buffer keeps kReadPadBytes of readable slack behind the requested size, also when a cached allocation is
    // reused: the compare / window loads of the kernels read up to ~9 words past a packed read (common.cuh kReadPadBytes).
    int ensure(size_t bytes) {
        if (p && bytes + kReadPadBytes <= cap) return ALGA_OK;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        size_t want = bytes + bytes / 8 + 2 * kReadPadBytes;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes + kReadPadBytes;
            e = cudaMalloc(&p, want);
        }
        if (e != cudaSuccess) {
            p = nullptr;
            return fail(ALGA_E_NOMEM, "cudaMalloc of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return ALGA_OK;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const {
        return reinterpret_cast<T *>(p);
    }
};

struct HostBuf {  // page-locked staging buffer
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes) {
        if (bytes <= cap && p) return ALGA_OK;
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        const size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMallocHost(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(ALGA_E_NOMEM, "cudaMallocHost of %zu bytes failed: %s", want, cudaGetErrorString(e));
        }
        cap = want;
        return ALGA_OK;
    }
    void release() {
        if (p) cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

struct Counters {
    unsigned long long n_edges;
    uint32_t n_list;   // phase-1 edges in list form (staged interface)
    uint32_t n_over;   // phase-1 edges that did not fit their fixed-capacity row
    uint32_t n_spill;
    uint32_t n_big;
    uint32_t n_hard1;  // source reads phase 1's fast kernel handed to the generic kernel
    uint32_t rev_overflow;  // the CSR rebuild of the transposed graph did not fit (sharded runs only)
    uint32_t n_hard1b;      // ... of those, the source reads the second fast pass handed to the generic kernel
    uint32_t n_spill2;      // targets the second fast phase-2 pass handed to the generic kernels
};

constexpr uint32_t kRowCapDefault = 16;  // entries per target in the fixed-capacity rows of the transposed phase-1 graph

double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

}  // namespace

struct alga_ps_plan {
    alga_ps_params params{};
    int sm_count = 148;
    LaunchCfg cfg;
    uint64_t launches = 0;
    uint64_t spilled = 0;

    // bound read set
    bool bound = false;
    ReadsDev R{};
    ReadStats stats{};
    PsDev P{};
    bool peek_by_kernel = false;  // control read-backs through launch_peek instead of the DMA engine (set while a bulk
                                  // download shares the device -> host engine, alga_gpu_files_to_graph)
    bool swap_direction = false;  // rs > max_l + 1 corner of the reference (SURVEY.md A.1 note 2)
    bool index_valid = false;

    // owned copies (upload path)
    DevBuf words, word_off, len, from, to;
    DevBuf slots;  // sector-aligned copy of a fixed-stride, equal-length read set (alga_ps_plan_run), what the fast kernels read
    // workspace
    DevBuf stats_d, counters_d, tp, ts, rows, over, list1, hard1, hard1b, spill_queue2, indeg, rev_off, rev, triples, triples1, outdeg, scan_ws,
        spill_queue, caps, spill_off, spill_store, row_off, nbr, off, big_rows, tmp_nbr, tmp_off, sort_ws[3];  // sort_ws: prefix index, suffix index, CSR
    uint32_t over_cap = 0;
    uint32_t row_cap = kRowCapDefault;  // ALGA_PS_ROW_CAP (testing: small rows force the overflow paths)
    SeedTable Tp{}, Ts{};
    Counters *h_counters = nullptr;  // pinned
    ReadStats *h_stats = nullptr;    // pinned
    uint64_t *h_u64 = nullptr;       // pinned
    HostBuf h_row_off, h_nbr, h_off; // pinned result staging (alga_ps_plan_result_host_pinned)

    // result
    uint32_t res_lo = 0, res_hi = 0;
    uint64_t n_edges = 0;
    double last_device_ms = 0;
    double stage_ms[8] = {0};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t side = nullptr;  // the suffix table is built here while phase 1 (prefix table only) runs
    cudaEvent_t ev_stage[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};

    ~alga_ps_plan() {
        DevBuf *all[] = {&words, &word_off, &len, &from, &to, &slots, &stats_d, &counters_d, &tp, &ts, &rows, &over, &list1, &hard1, &hard1b, &spill_queue2, &indeg, &rev_off, &rev,
                         &triples, &triples1, &outdeg, &scan_ws, &spill_queue, &caps, &spill_off, &spill_store, &row_off,
                         &nbr, &off, &big_rows, &tmp_nbr, &tmp_off, &sort_ws[0], &sort_ws[1], &sort_ws[2]};
        for (DevBuf *b : all) b->release();
        h_row_off.release();
        h_nbr.release();
        h_off.release();
        if (h_counters) cudaFreeHost(h_counters);
        if (h_stats) cudaFreeHost(h_stats);
        if (h_u64) cudaFreeHost(h_u64);
        if (ev0) cudaEventDestroy(ev0);
        if (ev1) cudaEventDestroy(ev1);
        if (ev_fork) cudaEventDestroy(ev_fork);
        if (ev_join) cudaEventDestroy(ev_join);
        if (side) cudaStreamDestroy(side);
        for (cudaEvent_t e : ev_stage)
            if (e) cudaEventDestroy(e);
    }
};

namespace {

int use_device(alga_ps_plan *plan) {
    CK(cudaSetDevice(plan->params.device));
    return ALGA_OK;
}

// small device -> host read-back into one of the plan's page-locked scalars
int plan_peek(alga_ps_plan *plan, void *dst, const void *src, uint32_t bytes, cudaStream_t s) {
    if (plan->peek_by_kernel) {
        launch_peek(dst, src, bytes, s, plan->cfg);
        CK(cudaGetLastError());
    } else {
        CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
    }
    return ALGA_OK;
}

int read_counters(alga_ps_plan *plan, cudaStream_t s) {
    CKR(plan_peek(plan, plan->h_counters, plan->counters_d.p, sizeof(Counters), s));
    CK(cudaStreamSynchronize(s));
    return ALGA_OK;
}

int resolve_params(alga_ps_plan *plan) {
    const alga_ps_params &pp = plan->params;
    if (pp.min_overlap < 1) return fail(ALGA_E_INVALID, "min_overlap must be >= 1 (got %d)", pp.min_overlap);
    if (pp.min_offset < 0) return fail(ALGA_E_INVALID, "min_offset must be >= 0 (got %d)", pp.min_offset);
    const int cap = pp.max_len_cap > 0 ? pp.max_len_cap : 500;
    PsDev &P = plan->P;
    P.lmin = pp.min_overlap;
    P.rs = pp.rs_min_overlap;
    P.min_offset = pp.min_offset;
    // The reference loop `while (L <= min(maxReadLength, 500)) { L++; ... }` (GraphCreatorPrefSuf.cpp:92-95)
    // also runs the iteration L = min(maxReadLength, cap) + 1, which is a real overlap length for reads
    // longer than the cap.  max_l is that last iterated length.
    P.max_l = (int32_t) (plan->stats.max_len < (uint32_t) cap ? plan->stats.max_len : (uint32_t) cap) + 1;
    P.seed_nt = P.lmin < 32 ? P.lmin : 32;
    P.seed_mask = P.seed_nt == 32 ? ~0ull : ((1ull << (2 * P.seed_nt)) - 1ull);
    P.uniform_len = (plan->stats.min_len == plan->stats.max_len && plan->R.n) ? plan->stats.max_len : 0u;
    // the reference transposes the phase-1 graph when L reaches rs (GraphCreatorPrefSuf.cpp:288).  If
    // that never happens the single final transpose leaves the phase-1 edges reversed.
    plan->swap_direction = P.rs > P.max_l && P.max_l >= P.lmin;
    return ALGA_OK;
}

int compute_stats(alga_ps_plan *plan, cudaStream_t s, uint32_t max_len_hint) {
    CKR(plan->stats_d.ensure(sizeof(ReadStats)));
    launch_read_stats(plan->R, plan->params.min_overlap, plan->params.min_offset, plan->stats_d.as<ReadStats>(), s,
                      plan->cfg);
    CK(cudaGetLastError());
    CKR(plan_peek(plan, plan->h_stats, plan->stats_d.p, sizeof(ReadStats), s));
    CK(cudaStreamSynchronize(s));
    plan->stats = *plan->h_stats;
    if (max_len_hint && max_len_hint != plan->stats.max_len)
        return fail(ALGA_E_INVALID, "max_len_nt hint %u does not match the read set (%u)", max_len_hint,
                    plan->stats.max_len);
    return resolve_params(plan);
}

int g_bucket_load = 0;  // alga_ps_set_bucket_load(); 0 = default

uint32_t buckets_for(uint32_t entries) {
    // mean occupancy 3 of kBucketCap = 20 slots.  Minimizer buckets fill unevenly -- all reads that start within a few
    // nucleotides of one another share theirs -- so a probe meets 5-6 entries on average and an overflowing bucket in
    // about 0.2 % of the cases (scripts/probes/minimizer_locality.py); 4 would triple that.  Sharded builds, which ship
    // their table slices over NVLink, may prefer denser tables (alga_ps_set_bucket_load).
    static const int env_load = [] {
        const char *e = getenv("ALGA_PS_BUCKET_LOAD");
        const int v = e ? atoi(e) : 0;
        return v >= 1 && v <= 12 ? v : 0;
    }();
    const int load = g_bucket_load ? g_bucket_load : (env_load ? env_load : 3);
    uint64_t nb = ((uint64_t) entries + load - 1) / load;
    if (nb < 64) nb = 64;
    return (uint32_t) nb;
}

void size_table(SeedTable &t, uint32_t entries, int world = 1) {
    t.n_buckets = buckets_for(entries);
    if (world > 1) t.n_buckets = (t.n_buckets + world - 1) / world * world;
    t.slice = t.n_buckets / (world > 1 ? world : 1);
    t.min_m = 0;
}

// m of the minimizer that chooses the bucket of a seed window (common.cuh SeedTable): seed_nt - 12, i.e. 13 m-mers per window
// (the fast kernels keep exactly that many in registers, tpr_kernels.cu kNM): 20 for the usual 32-nucleotide seed.  Seeds
// shorter than 16 nucleotides (min_overlap < 16) are their own minimizer -- every window its own bucket, as in a plain hash
// table -- and are left to the generic kernels (fast_seed()).
uint32_t minimizer_setting(int seed_nt) { return (uint32_t) (seed_nt >= 16 ? seed_nt - (kNM - 1) : seed_nt); }
bool fast_seed(const alga_ps_plan *plan) { return plan->P.seed_nt >= 16; }

// s2 != nullptr: everything that concerns the suffix table goes to s2
int stage_index_size(alga_ps_plan *plan, size_t *bytes_prefix, size_t *bytes_suffix) {  // the two tables, sized and allocated
    if (!plan->bound) return fail(ALGA_E_INVALID, "no read set bound to the plan");
    size_table(plan->Tp, plan->stats.n_prefix);
    size_table(plan->Ts, plan->stats.n_suffix);
    plan->Tp.min_m = plan->Ts.min_m = minimizer_setting(plan->P.seed_nt);
    *bytes_prefix = (size_t) plan->Tp.n_buckets * kBucketWords * 4, *bytes_suffix = (size_t) plan->Ts.n_buckets * kBucketWords * 4;
    CKR(plan->tp.ensure(*bytes_prefix));
    CKR(plan->ts.ensure(*bytes_suffix));
    plan->Tp.slots = plan->tp.as<uint32_t>();
    plan->Ts.slots = plan->ts.as<uint32_t>();
    return ALGA_OK;
}
int stage_index_begin(alga_ps_plan *plan, cudaStream_t s, cudaStream_t s2 = nullptr) {
    size_t bp = 0, bs = 0;
    CKR(stage_index_size(plan, &bp, &bs));
    CK(cudaMemsetAsync(plan->tp.p, 0, bp, s));
    CK(cudaMemsetAsync(plan->ts.p, 0, bs, s2 ? s2 : s));
    return ALGA_OK;
}

// The index build and the CSR assembly by sorting (sorted_stages.cu) instead of by random atomics; ALGA_PS_SORTED=0 keeps the
// scatter kernels (A/B, and the path the sharded build still takes).
bool sorted_stages() {
    static const bool on = !(getenv("ALGA_PS_SORTED") && atoi(getenv("ALGA_PS_SORTED")) == 0);
    return on;
}

int stage_index(alga_ps_plan *plan, cudaStream_t s) {
    CKR(stage_index_begin(plan, s));
    launch_build_index(plan->R, plan->P, plan->Tp, plan->Ts, 0, plan->R.n, 0u, 0xFFFFFFFFu, s, plan->cfg);
    CK(cudaGetLastError());
    plan->index_valid = true;
    return ALGA_OK;
}

// phase 1 for the source reads [lo, hi): first fast pass, second fast pass over what it gave up on, generic kernel
// over what is left.  All queue lengths stay on the device.
int run_phase1(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const Phase1Out &p1, cudaStream_t s) {
    const uint32_t n = hi - lo;
    Counters *dc = plan->counters_d.as<Counters>();
    CKR(plan->hard1.ensure((size_t) (n ? n : 1) * 4));
    CKR(plan->hard1b.ensure((size_t) (n ? n : 1) * 4));
    CK(cudaMemsetAsync(&dc->n_hard1, 0, 4, s));
    CK(cudaMemsetAsync(&dc->n_hard1b, 0, 4, s));
    const int force = (plan->params.flags & ALGA_PS_FORCE_GENERIC) || !fast_seed(plan);
    launch_phase1_tpr(plan->R, plan->Tp, plan->P, plan->stats.max_len, lo, hi, nullptr, nullptr, p1,
                      plan->hard1.as<uint32_t>(), &dc->n_hard1, force, s, plan->cfg);
    const uint32_t *q2 = nullptr, *nq2 = nullptr;
    if (!force) {  // runs only if the queue is long (kSecondPassMin, decided on the device)
        launch_phase1_tpr(plan->R, plan->Tp, plan->P, plan->stats.max_len, lo, hi, plan->hard1.as<uint32_t>(), &dc->n_hard1,
                          p1, plan->hard1b.as<uint32_t>(), &dc->n_hard1b, 0, s, plan->cfg);
        q2 = plan->hard1b.as<uint32_t>();
        nq2 = &dc->n_hard1b;
    }
    launch_phase1_queue(plan->R, plan->Tp, plan->P, force ? n : (n < 4096 ? n : 4096), plan->hard1.as<uint32_t>(),
                        &dc->n_hard1, q2, nq2, p1, s, plan->cfg);
    CK(cudaGetLastError());
    return ALGA_OK;
}

// phase 2 (+ spill path) for targets [lo,hi) given rev rows; leaves triples in plan->triples, count in h_counters
int run_phase2(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const RowsView &rows, uint32_t *outdeg, cudaStream_t s,
               const ShardOut &sh = ShardOut{}) {
    const uint32_t n = hi - lo;
    int list_cap = plan->params.list_cap > 0 ? plan->params.list_cap : 64;
    if (list_cap > 2048) list_cap = 2048;
    uint64_t edge_cap = (uint64_t) n * 3 + (1u << 16);
    CKR(plan->spill_queue.ensure((size_t) (n ? n : 1) * 4));
    CKR(plan->spill_queue2.ensure((size_t) (n ? n : 1) * 4));
    for (int attempt = 0; attempt < 3; attempt++) {
        CKR(plan->triples.ensure((size_t) edge_cap * 12));
        Counters *dc = plan->counters_d.as<Counters>();
        CK(cudaMemsetAsync(&dc->n_edges, 0, 8, s));
        CK(cudaMemsetAsync(&dc->n_spill, 0, 4, s));
        CK(cudaMemsetAsync(&dc->n_spill2, 0, 4, s));
        if (outdeg) CK(cudaMemsetAsync(outdeg, 0, (size_t) plan->R.n * 4, s));
        Phase2Out out{plan->triples.as<int32_t>(), &dc->n_edges, edge_cap, outdeg, plan->spill_queue.as<uint32_t>(),
                      &dc->n_spill, sh};
        if (plan->params.list_cap > 0)  // testing: generic kernel with a tiny on-chip list
            launch_phase2(plan->R, plan->Ts, plan->P, lo, hi, rows, list_cap, out, s, plan->cfg);
        else
            launch_phase2_tpr(plan->R, plan->Ts, plan->P, plan->stats.max_len, lo, hi, nullptr, nullptr, rows, out,
                              (plan->params.flags & ALGA_PS_FORCE_GENERIC) || !fast_seed(plan), s, plan->cfg);
        const uint32_t *spill_q = plan->spill_queue.as<uint32_t>();
        const bool second_pass = false;  // the fast kernel queues every match of a window now: nothing a second fast pass could add
        if (second_pass) {  // what the first pass gave up on, once more with four matches per window
            Phase2Out out2 = out;
            out2.spill_queue = plan->spill_queue2.as<uint32_t>();
            out2.n_spill = &dc->n_spill2;
            launch_phase2_tpr(plan->R, plan->Ts, plan->P, plan->stats.max_len, lo, hi, plan->spill_queue.as<uint32_t>(),
                              &dc->n_spill, rows, out2, 0, s, plan->cfg);
            spill_q = plan->spill_queue2.as<uint32_t>();
        }
        CK(cudaGetLastError());
        CKR(read_counters(plan, s));
        const bool second_ran = second_pass && plan->h_counters->n_spill >= kSecondPassMin;
        if (!second_ran) spill_q = plan->spill_queue.as<uint32_t>();
        const uint32_t n_spill = second_ran ? plan->h_counters->n_spill2 : plan->h_counters->n_spill;
        plan->spilled = n_spill;
        if (n_spill) {
            // targets whose in-neighbour list outgrew shared memory: exact capacity, lists in HBM
            CKR(plan->caps.ensure((size_t) n_spill * 4));
            CKR(plan->spill_off.ensure(((size_t) n_spill + 1) * 8));
            CKR(plan->scan_ws.ensure(scan_workspace_bytes(n_spill)));
            launch_phase2_count(plan->R, plan->Ts, plan->P, lo, rows, spill_q, n_spill,
                                plan->caps.as<uint32_t>(), s, plan->cfg);
            launch_scan_u64(plan->caps.as<uint32_t>(), plan->spill_off.as<uint64_t>(), n_spill, plan->scan_ws.p, s,
                            plan->cfg);
            CK(cudaGetLastError());
            CKR(plan_peek(plan, plan->h_u64, plan->spill_off.as<uint64_t>() + n_spill, 8, s));
            CK(cudaStreamSynchronize(s));
            const uint64_t total = *plan->h_u64;
            CKR(plan->spill_store.ensure((size_t) total * 12 + 16));
            launch_phase2_spill(plan->R, plan->Ts, plan->P, lo, rows, spill_q, n_spill,
                                plan->spill_off.as<uint64_t>(), plan->spill_store.as<uint32_t>(), out, s, plan->cfg);
            CK(cudaGetLastError());
            CKR(read_counters(plan, s));
        }
        if (plan->h_counters->n_edges <= edge_cap) return ALGA_OK;
        edge_cap = plan->h_counters->n_edges + 1024;  // retry with the exact size
    }
    return fail(ALGA_E_CAPACITY, "edge buffer overflow persisted after retries");
}

int build_rev_from_counts(alga_ps_plan *plan, uint32_t n_targets, cudaStream_t s) {
    CKR(plan->rev_off.ensure(((size_t) n_targets + 1) * 4));
    CKR(plan->scan_ws.ensure(scan_workspace_bytes(n_targets)));
    launch_scan_u32(plan->indeg.as<uint32_t>(), plan->rev_off.as<uint32_t>(), n_targets, plan->scan_ws.p, s, plan->cfg);
    CK(cudaGetLastError());
    return ALGA_OK;
}

// rows in place (unsorted), their total on its way into *plan->h_u64: sort every row by (neighbour, offset), publish the result
int finish_csr(alga_ps_plan *plan, uint32_t lo, uint32_t hi, uint64_t n_tr, cudaStream_t s) {
    const uint32_t n = hi - lo;
    Counters *dc = plan->counters_d.as<Counters>();
    CK(cudaMemsetAsync(&dc->n_big, 0, 4, s));
    launch_sort_rows(plan->row_off.as<uint64_t>(), n, plan->nbr.as<int32_t>(), plan->off.as<int32_t>(),
                     plan->big_rows.as<uint32_t>(), &dc->n_big, s, plan->cfg);
    CK(cudaGetLastError());
    CKR(read_counters(plan, s));
    if (plan->h_counters->n_big) {
        CKR(plan->tmp_nbr.ensure((size_t) n_tr * 4));
        CKR(plan->tmp_off.ensure((size_t) n_tr * 4));
        launch_sort_big_rows(plan->row_off.as<uint64_t>(), plan->big_rows.as<uint32_t>(), plan->h_counters->n_big,
                             plan->nbr.as<int32_t>(), plan->off.as<int32_t>(), plan->tmp_nbr.as<int32_t>(),
                             plan->tmp_off.as<int32_t>(), s, plan->cfg);
        CK(cudaGetLastError());
    }
    if (plan->h_counters->n_big) CK(cudaStreamSynchronize(s));
    plan->n_edges = *plan->h_u64;
    plan->res_lo = lo;
    plan->res_hi = hi;
    return ALGA_OK;
}

// outdeg_ready: plan->outdeg already holds the row sizes, indexed by global read id
int stage_csr(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const int32_t *triples, uint64_t n_tr, int swap,
              bool outdeg_ready, cudaStream_t s) {
    const uint32_t n = hi - lo;
    if (sorted_stages() && n_tr < 0x7FFFFFFFull) {
        // rows by sorting the edges on their source: sequential traffic instead of one atomic + one random store per edge
        CKR(plan->row_off.ensure(((size_t) n + 1) * 8));
        CKR(plan->nbr.ensure((size_t) (n_tr ? n_tr : 1) * 4));
        CKR(plan->off.ensure((size_t) (n_tr ? n_tr : 1) * 4));
        CKR(plan->big_rows.ensure((size_t) (n ? n : 1) * 4));
        CKR(plan->sort_ws[2].ensure(sorted_csr_workspace_bytes(n_tr)));
        if (launch_sorted_csr(triples, n_tr, lo, hi, swap, plan->sort_ws[2].p, plan->row_off.as<uint64_t>(), plan->nbr.as<int32_t>(),
                              plan->off.as<int32_t>(), s, plan->cfg))
            return fail(ALGA_E_CUDA, "CSR: radix sort failed");
        CKR(plan_peek(plan, plan->h_u64, plan->row_off.as<uint64_t>() + n, 8, s));  // read with the counters
        return finish_csr(plan, lo, hi, n_tr, s);
    }
    CKR(plan->outdeg.ensure((size_t) (plan->R.n ? plan->R.n : 1) * 4));
    uint32_t *outdeg = plan->outdeg.as<uint32_t>();
    if (!outdeg_ready) {
        CK(cudaMemsetAsync(outdeg, 0, (size_t) (n ? n : 1) * 4, s));
        launch_count_sources(triples, n_tr, lo, hi, swap, outdeg, s, plan->cfg);
    } else {
        outdeg += lo;  // counted by global id
    }
    CKR(plan->row_off.ensure(((size_t) n + 1) * 8));
    CKR(plan->scan_ws.ensure(scan_workspace_bytes(n)));
    launch_scan_u64(outdeg, plan->row_off.as<uint64_t>(), n, plan->scan_ws.p, s, plan->cfg);
    CKR(plan_peek(plan, plan->h_u64, plan->row_off.as<uint64_t>() + n, 8, s));  // read with the counters
    CKR(plan->nbr.ensure((size_t) (n_tr ? n_tr : 1) * 4));
    CKR(plan->off.ensure((size_t) (n_tr ? n_tr : 1) * 4));
    CKR(plan->big_rows.ensure((size_t) (n ? n : 1) * 4));
    if (n_tr < 0xFFFFFFFFull) {
        CKR(plan->tmp_nbr.ensure((size_t) (n_tr ? n_tr : 1) * 8));
        launch_scatter_csr_pairs(triples, n_tr, lo, hi, swap, plan->row_off.as<uint64_t>(), outdeg, plan->tmp_nbr.p,
                                 plan->nbr.as<int32_t>(), plan->off.as<int32_t>(), s, plan->cfg);
    } else {
        launch_scatter_csr(triples, n_tr, lo, hi, swap, plan->row_off.as<uint64_t>(), outdeg, plan->nbr.as<int32_t>(),
                           plan->off.as<int32_t>(), s, plan->cfg);
    }
    return finish_csr(plan, lo, hi, n_tr, s);
}

}  // namespace

namespace alga {
int set_error(int code, const char *msg) { return fail(code, "%s", msg); }  // for the other translation units
}

// ================================================================================================
extern "C" {

const char *alga_gpu_last_error(void) { return g_err; }
const char *alga_gpu_version(void) { return "alga_b200 0.1.0 (sm_100a)"; }

int alga_gpu_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int alga_ps_plan_create(alga_ps_plan **out, const alga_ps_params *params) {
    if (!out || !params) return fail(ALGA_E_INVALID, "null argument");
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ALGA_E_CUDA, "no CUDA device available (%s); libalga_gpu has no CPU fallback",
                    e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    }
    if (params->device < 0 || params->device >= ndev) return fail(ALGA_E_INVALID, "device %d out of range", params->device);
    alga_ps_plan *plan = new (std::nothrow) alga_ps_plan();
    if (!plan) return fail(ALGA_E_NOMEM, "out of host memory");
    plan->params = *params;
    plan->cfg.launches = &plan->launches;
    if (const char *e = getenv("ALGA_PS_ROW_CAP")) {
        const int v = atoi(e);
        if (v >= 1 && v <= 32) plan->row_cap = (uint32_t) v;
    }
    int r = [&]() -> int {
        CK(cudaSetDevice(params->device));
        int sm = 0;
        CK(cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, params->device));
        plan->sm_count = plan->cfg.sm_count = sm;
        CK(cudaMallocHost((void **) &plan->h_counters, sizeof(Counters)));
        CK(cudaMallocHost((void **) &plan->h_stats, sizeof(ReadStats)));
        CK(cudaMallocHost((void **) &plan->h_u64, 8));
        CK(cudaEventCreate(&plan->ev0));
        CK(cudaEventCreate(&plan->ev1));
        CK(cudaEventCreateWithFlags(&plan->ev_fork, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&plan->ev_join, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&plan->side, cudaStreamNonBlocking));
        for (cudaEvent_t &e : plan->ev_stage) CK(cudaEventCreate(&e));
        CKR(plan->counters_d.ensure(sizeof(Counters)));
        return ALGA_OK;
    }();
    if (r != ALGA_OK) {
        delete plan;
        return r;
    }
    *out = plan;
    return ALGA_OK;
}

void alga_ps_plan_destroy(alga_ps_plan *plan) {
    if (!plan) return;
    cudaSetDevice(plan->params.device);
    delete plan;
}

int alga_ps_plan_bind_reads_device(alga_ps_plan *plan, const alga_reads *r, uint32_t max_len_nt) {
    if (!plan || !r) return fail(ALGA_E_INVALID, "null argument");
    if (r->n_reads && (!r->words || !r->len_nt)) return fail(ALGA_E_INVALID, "words / len_nt must not be null");
    if (!r->word_off && r->stride_words == 0 && r->n_reads) return fail(ALGA_E_INVALID, "word_off is null and stride_words is 0");
    if (r->n_reads > 0x7FFFFFFFu) return fail(ALGA_E_INVALID, "too many reads (ids are int32 in the edge arrays)");
    CKR(use_device(plan));
    plan->R.words = r->words;
    plan->R.word_off = r->word_off;
    plan->R.len = r->len_nt;
    plan->R.from = r->align_from;
    plan->R.to = r->align_to;
    plan->R.n = r->n_reads;
    plan->R.stride = r->word_off ? 0 : r->stride_words;
    plan->bound = true;
    plan->index_valid = false;
    return compute_stats(plan, 0, max_len_nt);
}

int alga_ps_plan_upload_reads(alga_ps_plan *plan, const alga_reads *h) {
    if (!plan || !h) return fail(ALGA_E_INVALID, "null argument");
    if (h->n_reads && (!h->words || !h->len_nt)) return fail(ALGA_E_INVALID, "words / len_nt must not be null");
    if (!h->word_off && h->stride_words == 0 && h->n_reads) return fail(ALGA_E_INVALID, "word_off is null and stride_words is 0");
    CKR(use_device(plan));
    const uint32_t n = h->n_reads;
    const uint64_t n_words = h->word_off ? h->word_off[n] : (uint64_t) n * h->stride_words;
    CKR(plan->words.ensure((size_t) n_words * 4 + kReadPadBytes));
    CKR(plan->len.ensure((size_t) (n ? n : 1) * 4));
    cudaStream_t s = 0;
    CK(cudaMemcpyAsync(plan->words.p, h->words, (size_t) n_words * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync((char *) plan->words.p + (size_t) n_words * 4, 0, kReadPadBytes, s));
    CK(cudaMemcpyAsync(plan->len.p, h->len_nt, (size_t) n * 4, cudaMemcpyHostToDevice, s));
    alga_reads d = *h;
    d.words = plan->words.as<uint32_t>();
    d.len_nt = plan->len.as<uint32_t>();
    if (h->word_off) {
        CKR(plan->word_off.ensure(((size_t) n + 1) * 8));
        CK(cudaMemcpyAsync(plan->word_off.p, h->word_off, ((size_t) n + 1) * 8, cudaMemcpyHostToDevice, s));
        d.word_off = plan->word_off.as<uint64_t>();
    }
    if (h->align_from) {
        CKR(plan->from.ensure(n ? n : 1));
        CK(cudaMemcpyAsync(plan->from.p, h->align_from, n, cudaMemcpyHostToDevice, s));
        d.align_from = plan->from.as<uint8_t>();
    }
    if (h->align_to) {
        CKR(plan->to.ensure(n ? n : 1));
        CK(cudaMemcpyAsync(plan->to.p, h->align_to, n, cudaMemcpyHostToDevice, s));
        d.align_to = plan->to.as<uint8_t>();
    }
    return alga_ps_plan_bind_reads_device(plan, &d, 0);
}

int alga_ps_stage_index(alga_ps_plan *plan, void *stream) {
    if (!plan) return fail(ALGA_E_INVALID, "null plan");
    CKR(use_device(plan));
    return stage_index(plan, (cudaStream_t) stream);
}

int alga_ps_stage_phase1(alga_ps_plan *plan, uint32_t lo, uint32_t hi, void *stream, const int32_t **dev_triples,
                         uint64_t *n_triples) {
    if (!plan || !dev_triples || !n_triples) return fail(ALGA_E_INVALID, "null argument");
    if (!plan->index_valid) return fail(ALGA_E_INVALID, "seed index not built (call alga_ps_stage_index)");
    if (lo > hi || hi > plan->R.n) return fail(ALGA_E_INVALID, "bad range [%u,%u)", lo, hi);
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    const uint32_t n = hi - lo;
    const size_t max_edges = (size_t) (n ? n : 1) * kSmallEdgesKept;
    CKR(plan->list1.ensure(max_edges * sizeof(Edge1)));
    CKR(plan->triples1.ensure(max_edges * 12));
    Counters *dc = plan->counters_d.as<Counters>();
    CK(cudaMemsetAsync(&dc->n_list, 0, 4, s));
    Phase1Out p1{1, nullptr, nullptr, 0, plan->list1.as<Edge1>(), &dc->n_list, (uint32_t) max_edges, 0u, ShardOut{}};
    CKR(run_phase1(plan, lo, hi, p1, s));
    launch_edges_to_triples(plan->list1.as<Edge1>(), &dc->n_list, max_edges, plan->triples1.as<int32_t>(), s, plan->cfg);
    CK(cudaGetLastError());
    CKR(read_counters(plan, s));
    *dev_triples = plan->triples1.as<int32_t>();
    *n_triples = plan->h_counters->n_list;
    return ALGA_OK;
}

int alga_ps_stage_phase2(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const int32_t *tin, uint64_t n_in, void *stream,
                         const int32_t **dev_triples_out, uint64_t *n_out) {
    if (!plan || !dev_triples_out || !n_out) return fail(ALGA_E_INVALID, "null argument");
    if (!plan->index_valid) return fail(ALGA_E_INVALID, "seed index not built (call alga_ps_stage_index)");
    if (lo > hi || hi > plan->R.n) return fail(ALGA_E_INVALID, "bad range [%u,%u)", lo, hi);
    if (n_in && !tin) return fail(ALGA_E_INVALID, "null triples");
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    const uint32_t n = hi - lo;
    CKR(plan->indeg.ensure((size_t) (n ? n : 1) * 4));
    CK(cudaMemsetAsync(plan->indeg.p, 0, (size_t) (n ? n : 1) * 4, s));
    launch_count_targets(tin, n_in, lo, hi, plan->indeg.as<uint32_t>(), s, plan->cfg);
    CKR(build_rev_from_counts(plan, n, s));
    CKR(plan->rev.ensure((size_t) (n_in ? n_in : 1) * sizeof(RevEntry)));
    launch_scatter_rev_triples(plan->R, tin, n_in, lo, hi, plan->rev_off.as<uint32_t>(), plan->indeg.as<uint32_t>(),
                               plan->rev.as<RevEntry>(), s, plan->cfg);
    CK(cudaGetLastError());
    const RowsView view{nullptr, nullptr, 0, nullptr, nullptr, plan->rev_off.as<uint32_t>(), plan->rev.as<RevEntry>()};
    CKR(run_phase2(plan, lo, hi, view, nullptr, s));
    *dev_triples_out = plan->triples.as<int32_t>();
    *n_out = plan->h_counters->n_edges;
    return ALGA_OK;
}

int alga_ps_stage_csr(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const int32_t *triples, uint64_t n, int swap,
                      void *stream) {
    if (!plan) return fail(ALGA_E_INVALID, "null plan");
    if (!plan->bound) return fail(ALGA_E_INVALID, "no read set bound to the plan");
    if (lo > hi || hi > plan->R.n) return fail(ALGA_E_INVALID, "bad range [%u,%u)", lo, hi);
    if (n && !triples) return fail(ALGA_E_INVALID, "null triples");
    CKR(use_device(plan));
    return stage_csr(plan, lo, hi, triples, n, swap, false, (cudaStream_t) stream);
}


int alga_ps_plan_bind_reads_uniform(alga_ps_plan *plan, const alga_reads *r, uint32_t len_nt) {
    if (!plan || !r) return fail(ALGA_E_INVALID, "null argument");
    if (!r->words || !r->len_nt || r->word_off || r->stride_words == 0 || len_nt == 0 || r->n_reads == 0)
        return fail(ALGA_E_INVALID, "uniform binding needs words, len_nt, a fixed stride and a length");
    if (r->align_from || r->align_to) return fail(ALGA_E_INVALID, "uniform binding takes no flag arrays");
    if (r->n_reads > 0x7FFFFFFFu) return fail(ALGA_E_INVALID, "too many reads (ids are int32 in the edge arrays)");
    CKR(use_device(plan));
    plan->R.words = r->words;
    plan->R.word_off = nullptr;
    plan->R.len = r->len_nt;
    plan->R.from = plan->R.to = nullptr;
    plan->R.n = r->n_reads;
    plan->R.stride = r->stride_words;
    plan->bound = true;
    plan->index_valid = false;
    // no pass over the reads (they may not be resident yet): every read has this length and takes part
    plan->stats.max_len = plan->stats.min_len = len_nt;
    plan->stats.n_prefix = plan->stats.n_suffix = r->n_reads;
    return resolve_params(plan);
}

int alga_ps_stage_index_range(alga_ps_plan *plan, uint32_t lo, uint32_t hi, int first, void *stream) {
    if (!plan) return fail(ALGA_E_INVALID, "null plan");
    if (lo > hi || hi > plan->R.n) return fail(ALGA_E_INVALID, "bad range [%u,%u)", lo, hi);
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    if (first) {
        CKR(stage_index_begin(plan, s));
        plan->launches = 0;
    } else if (!plan->Tp.slots) {
        return fail(ALGA_E_INVALID, "the first range of a build must be inserted with first != 0");
    }
    launch_build_index(plan->R, plan->P, plan->Tp, plan->Ts, lo, hi, 0u, 0xFFFFFFFFu, s, plan->cfg);
    CK(cudaGetLastError());
    plan->index_valid = true;
    return ALGA_OK;
}

// ---- sharded build: exchange through the ranks' workspaces (peer memory over NVLink) -----------------------
namespace {
struct ShardLayout {
    uint64_t cnt1_off, cnt2_off, seg1_off, seg2_off, total;
    uint32_t cap1, cap2;
};
ShardLayout shard_layout(uint32_t n_shard, int world) {
    ShardLayout L;
    L.cnt1_off = 0;
    L.cnt2_off = 64;
    L.seg1_off = 256;
    // segment sizes are multiples of 4 entries: every segment starts 16-byte aligned (the pull kernels load uint4)
    L.cap1 = (n_shard * (uint32_t) kSmallEdgesKept + 3u) & ~3u;           // worst case: every edge of the rank to one owner
    L.cap2 = (n_shard * (uint32_t) kSmallEdgesKept + 65536u + 3u) & ~3u;  // survivors per owner (checked; ALGA_E_CAPACITY beyond)
    L.seg2_off = (L.seg1_off + (uint64_t) world * L.cap1 * sizeof(Edge1) + 255) & ~255ull;
    L.total = (L.seg2_off + (uint64_t) world * L.cap2 * 12 + 255) & ~255ull;
    return L;
}
int check_shard(alga_ps_plan *plan, const alga_ps_shard *sh, uint32_t *lo, uint32_t *hi) {
    if (!plan || !sh) return fail(ALGA_E_INVALID, "null argument");
    if (sh->world < 1 || sh->world > 8 || sh->rank < 0 || sh->rank >= sh->world)
        return fail(ALGA_E_INVALID, "bad rank %d / world %d (1..8 GPUs of one box)", sh->rank, sh->world);
    if (!plan->bound || sh->n_total != plan->R.n) return fail(ALGA_E_INVALID, "n_total does not match the bound read set");
    if (sh->n_shard == 0 || (uint64_t) sh->n_shard * sh->world < sh->n_total) return fail(ALGA_E_INVALID, "n_shard too small");
    if ((uint64_t) sh->n_shard * kSmallEdgesKept + 65536u > 0xFFFFFFFFull) return fail(ALGA_E_INVALID, "shard too large");
    for (int p = 0; p < sh->world; p++)
        if (!sh->peer_ws[p]) return fail(ALGA_E_INVALID, "null workspace pointer for rank %d", p);
    if (plan->swap_direction) return fail(ALGA_E_INVALID, "rs_min_overlap beyond the longest read is not supported in sharded runs");
    if (sh->table_prefix && sh->table_suffix) {  // the caller's (sliced, exchanged) seed tables are the plan's tables
        size_table(plan->Tp, sh->n_total, sh->world);
        size_table(plan->Ts, sh->n_total, sh->world);
        plan->Tp.min_m = plan->Ts.min_m = minimizer_setting(plan->P.seed_nt);
        plan->Tp.slots = (uint32_t *) sh->table_prefix;
        plan->Ts.slots = (uint32_t *) sh->table_suffix;
        plan->index_valid = true;
    }
    const uint64_t l = (uint64_t) sh->rank * sh->n_shard, h = l + sh->n_shard;
    *lo = (uint32_t) (l < sh->n_total ? l : sh->n_total);
    *hi = (uint32_t) (h < sh->n_total ? h : sh->n_total);
    return ALGA_OK;
}
}  // namespace

uint64_t alga_ps_shard_ws_bytes(uint32_t n_shard, int32_t world) { return shard_layout(n_shard, world).total; }

uint64_t alga_ps_shard_table_bytes(uint32_t n_total, int32_t world) {
    SeedTable t{};
    size_table(t, n_total, world);
    return (uint64_t) t.n_buckets * kBucketWords * 4;
}

int alga_ps_shard_index_range(alga_ps_plan *plan, const alga_ps_shard *sh, uint32_t lo, uint32_t hi, int first, void *stream) {
    uint32_t my_lo, my_hi;
    CKR(check_shard(plan, sh, &my_lo, &my_hi));
    if (!sh->table_prefix || !sh->table_suffix) return fail(ALGA_E_INVALID, "null seed table pointer");
    if (lo > hi || hi > plan->R.n) return fail(ALGA_E_INVALID, "bad range [%u,%u)", lo, hi);
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    size_table(plan->Tp, sh->n_total, sh->world);
    size_table(plan->Ts, sh->n_total, sh->world);
    plan->Tp.min_m = plan->Ts.min_m = minimizer_setting(plan->P.seed_nt);
    plan->Tp.slots = (uint32_t *) sh->table_prefix;
    plan->Ts.slots = (uint32_t *) sh->table_suffix;
    const uint32_t slice = plan->Tp.slice, b_lo = (uint32_t) sh->rank * slice, b_hi = b_lo + slice;
    const int which = ((first >> 1) & 3) ? ((first >> 1) & 3) : 3;  // bit 0: prefix table, bit 1: suffix table
    if (first & 1) {
        if (which & 1) plan->launches = 0;
        const size_t off = (size_t) b_lo * kBucketWords * 4, bytes = (size_t) slice * kBucketWords * 4;
        if (which & 1) CK(cudaMemsetAsync((char *) sh->table_prefix + off, 0, bytes, s));
        if (which & 2) CK(cudaMemsetAsync((char *) sh->table_suffix + off, 0, bytes, s));
    }
    launch_build_index(plan->R, plan->P, plan->Tp, plan->Ts, lo, hi, b_lo, b_hi, s, plan->cfg, which);
    CK(cudaGetLastError());
    plan->index_valid = true;
    return ALGA_OK;
}

int alga_ps_shard_seed_keys(alga_ps_plan *plan, const alga_ps_shard *sh, const uint32_t *shard_words, uint32_t stride_words,
                            uint32_t n_reads, uint32_t *keys, void *stream) {
    uint32_t my_lo, my_hi;
    CKR(check_shard(plan, sh, &my_lo, &my_hi));
    if (!plan->P.uniform_len) return fail(ALGA_E_INVALID, "seed records need equal-length reads");
    if (n_reads && (!shard_words || !keys || stride_words == 0)) return fail(ALGA_E_INVALID, "null argument");
    CKR(use_device(plan));
    size_table(plan->Tp, sh->n_total, sh->world);
    size_table(plan->Ts, sh->n_total, sh->world);
    plan->Tp.min_m = plan->Ts.min_m = minimizer_setting(plan->P.seed_nt);
    launch_seed_keys(shard_words, stride_words, n_reads, plan->P, plan->Tp, plan->Ts, keys, (cudaStream_t) stream, plan->cfg);
    CK(cudaGetLastError());
    return ALGA_OK;
}

int alga_ps_shard_index_keys(alga_ps_plan *plan, const alga_ps_shard *sh, const uint32_t *keys, uint32_t lo, uint32_t hi, int first,
                             void *stream) {
    uint32_t my_lo, my_hi;
    CKR(check_shard(plan, sh, &my_lo, &my_hi));
    if (!sh->table_prefix || !sh->table_suffix) return fail(ALGA_E_INVALID, "null seed table pointer");
    if (lo > hi || hi > plan->R.n) return fail(ALGA_E_INVALID, "bad range [%u,%u)", lo, hi);
    if (hi > lo && !keys) return fail(ALGA_E_INVALID, "null seed records");
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    size_table(plan->Tp, sh->n_total, sh->world);
    size_table(plan->Ts, sh->n_total, sh->world);
    plan->Tp.min_m = plan->Ts.min_m = minimizer_setting(plan->P.seed_nt);
    plan->Tp.slots = (uint32_t *) sh->table_prefix;
    plan->Ts.slots = (uint32_t *) sh->table_suffix;
    const uint32_t slice = plan->Tp.slice, b_lo = (uint32_t) sh->rank * slice, b_hi = b_lo + slice;
    if (first) {
        plan->launches = 0;
        const size_t off = (size_t) b_lo * kBucketWords * 4, bytes = (size_t) slice * kBucketWords * 4;
        CK(cudaMemsetAsync((char *) sh->table_prefix + off, 0, bytes, s));
        CK(cudaMemsetAsync((char *) sh->table_suffix + off, 0, bytes, s));
    }
    launch_index_keys(keys, lo, hi - lo, plan->Tp, plan->Ts, b_lo, b_hi, s, plan->cfg);
    CK(cudaGetLastError());
    plan->index_valid = true;
    return ALGA_OK;
}

void alga_ps_set_bucket_load(int32_t load) { g_bucket_load = load >= 1 && load <= 12 ? load : 0; }

int alga_ps_shard_phase1(alga_ps_plan *plan, const alga_ps_shard *sh, void *stream) {
    uint32_t lo, hi;
    CKR(check_shard(plan, sh, &lo, &hi));
    if (!plan->index_valid) return fail(ALGA_E_INVALID, "seed index not built");
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    const ShardLayout L = shard_layout(sh->n_shard, sh->world);
    char *ws = (char *) sh->peer_ws[sh->rank];
    const uint32_t n = hi - lo;
    Counters *dc = plan->counters_d.as<Counters>();
    CK(cudaMemsetAsync(ws + L.cnt1_off, 0, 32, s));
    Phase1Out p1{2, nullptr, nullptr, 0, nullptr, nullptr, 0, 0u,
                 ShardOut{sh->world, sh->n_shard, (uint32_t *) (ws + L.cnt1_off), ws + L.seg1_off, L.cap1}};
    return run_phase1(plan, lo, hi, p1, s);
}

int alga_ps_shard_phase2(alga_ps_plan *plan, const alga_ps_shard *sh, void *stream) {
    uint32_t lo, hi;
    CKR(check_shard(plan, sh, &lo, &hi));
    if (!plan->index_valid) return fail(ALGA_E_INVALID, "seed index not built");
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    const ShardLayout L = shard_layout(sh->n_shard, sh->world);
    char *ws = (char *) sh->peer_ws[sh->rank];
    const uint32_t n = hi - lo;
    Counters *dc = plan->counters_d.as<Counters>();
    if (plan->over_cap < kOverScanMax) plan->over_cap = sh->n_shard / 8 + 65536;
    const uint64_t rev_cap = (uint64_t) (n ? n : 1) * kSmallEdgesKept * 2 + 65536;
    CKR(plan->rows.ensure((size_t) (n ? n : 1) * plan->row_cap * sizeof(RevEntry)));
    CKR(plan->over.ensure((size_t) plan->over_cap * sizeof(Edge1)));
    CKR(plan->indeg.ensure((size_t) (n ? n : 1) * 4));
    CKR(plan->rev_off.ensure(((size_t) n + 1) * 4));
    CKR(plan->rev.ensure((size_t) rev_cap * sizeof(RevEntry)));
    CKR(plan->scan_ws.ensure(scan_workspace_bytes(n)));
    CK(cudaMemsetAsync(plan->indeg.p, 0, (size_t) (n ? n : 1) * 4, s));
    CK(cudaMemsetAsync(&dc->n_list, 0, 12, s));  // n_list, n_over, n_spill
    CK(cudaMemsetAsync(&dc->rev_overflow, 0, 4, s));
    CK(cudaMemsetAsync(ws + L.cnt2_off, 0, 32, s));
    // rows of the transposed phase-1 graph for the targets this rank owns, pulled out of every rank's workspace
    const void *seg[8];
    const uint32_t *cnt[8];
    for (int p = 0; p < sh->world; p++) {
        const char *pw = (const char *) sh->peer_ws[p];
        seg[p] = pw + L.seg1_off + (uint64_t) sh->rank * L.cap1 * sizeof(Edge1);
        cnt[p] = (const uint32_t *) (pw + L.cnt1_off) + sh->rank;
    }
    Phase1Out rows_out{0, plan->indeg.as<uint32_t>(), plan->rows.as<RevEntry>(), plan->row_cap, plan->over.as<Edge1>(),
                       &dc->n_over, plan->over_cap, lo, ShardOut{}};
    CK(cudaEventRecord(plan->ev_stage[0], s));
    launch_pull_rows(seg, cnt, sh->world, L.cap1, n * (uint32_t) kSmallEdgesKept / (uint32_t) sh->world + 1, rows_out, s,
                     plan->cfg);
    launch_rebuild_rows_csr(&dc->n_over, plan->over_cap, plan->over.as<Edge1>(), plan->indeg.as<uint32_t>(),
                            plan->rows.as<RevEntry>(), plan->row_cap, n, plan->rev_off.as<uint32_t>(), plan->rev.as<RevEntry>(),
                            rev_cap, &dc->rev_overflow, plan->scan_ws.p, s, plan->cfg);
    CK(cudaGetLastError());
    CK(cudaEventRecord(plan->ev_stage[1], s));
    const RowsView view{plan->indeg.as<uint32_t>(), plan->rows.as<RevEntry>(), plan->row_cap, &dc->n_over,
                        plan->over.as<Edge1>(), plan->rev_off.as<uint32_t>(), plan->rev.as<RevEntry>()};
    const ShardOut so{sh->world > 1 ? sh->world : 2, sh->n_shard, (uint32_t *) (ws + L.cnt2_off), ws + L.seg2_off, L.cap2};
    CKR(run_phase2(plan, lo, hi, view, nullptr, s, so));
    CK(cudaEventRecord(plan->ev_stage[2], s));
    CK(cudaEventSynchronize(plan->ev_stage[2]));
    {
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, plan->ev_stage[0], plan->ev_stage[1]));
        plan->stage_ms[2] = ms;  // rows of the transposed graph (pull over NVLink)
        CK(cudaEventElapsedTime(&ms, plan->ev_stage[1], plan->ev_stage[2]));
        plan->stage_ms[3] = ms;  // phase 2
        plan->stage_ms[5] = plan->h_counters->n_over;
        plan->stage_ms[6] = plan->h_counters->n_hard1 >= kSecondPassMin ? plan->h_counters->n_hard1b : plan->h_counters->n_hard1;
    }
    if (plan->h_counters->n_over > plan->over_cap) {
        plan->over_cap = sh->n_shard * (uint32_t) kSmallEdgesKept * 2;
        return fail(ALGA_E_CAPACITY, "phase-1 overflow list too small; the plan has grown it, run the build again");
    }
    if (plan->h_counters->rev_overflow) return fail(ALGA_E_CAPACITY, "transposed phase-1 graph of this shard exceeds its buffer");
    uint32_t h_cnt2[8];
    CK(cudaMemcpyAsync(h_cnt2, ws + L.cnt2_off, 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int p = 0; p < sh->world; p++)
        if (h_cnt2[p] > L.cap2) return fail(ALGA_E_CAPACITY, "%u surviving edges for rank %d exceed the exchange segment (%u)", h_cnt2[p], p, L.cap2);
    return ALGA_OK;
}

int alga_ps_shard_csr(alga_ps_plan *plan, const alga_ps_shard *sh, void *stream) {
    uint32_t lo, hi;
    CKR(check_shard(plan, sh, &lo, &hi));
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    const ShardLayout L = shard_layout(sh->n_shard, sh->world);
    const uint32_t n = hi - lo;
    Counters *dc = plan->counters_d.as<Counters>();
    const void *seg[8];
    const uint32_t *cnt[8];
    for (int p = 0; p < sh->world; p++) {
        const char *pw = (const char *) sh->peer_ws[p];
        seg[p] = pw + L.seg2_off + (uint64_t) sh->rank * L.cap2 * 12;
        cnt[p] = (const uint32_t *) (pw + L.cnt2_off) + sh->rank;
    }
    CKR(plan->outdeg.ensure((size_t) (plan->R.n ? plan->R.n : 1) * 4));
    uint64_t cap_local = (uint64_t) n * 2 + 65536;
    for (int attempt = 0;; attempt++) {
        CKR(plan->triples.ensure((size_t) cap_local * 12));
        CK(cudaMemsetAsync(plan->outdeg.as<uint32_t>() + lo, 0, (size_t) (n ? n : 1) * 4, s));
        launch_pull_triples(seg, cnt, sh->world, L.cap2, n / (uint32_t) sh->world + 1, cap_local, plan->triples.as<int32_t>(),
                            &dc->n_edges, plan->outdeg.as<uint32_t>(), s, plan->cfg);
        CK(cudaGetLastError());
        CKR(read_counters(plan, s));
        if (plan->h_counters->n_edges <= cap_local) break;
        if (attempt) return fail(ALGA_E_CAPACITY, "edge buffer overflow persisted");
        cap_local = plan->h_counters->n_edges + 1024;
    }
    return stage_csr(plan, lo, hi, plan->triples.as<int32_t>(), plan->h_counters->n_edges, 0, true, s);
}

int alga_ps_plan_run(alga_ps_plan *plan, void *stream) {
    if (!plan) return fail(ALGA_E_INVALID, "null plan");
    if (!plan->bound) return fail(ALGA_E_INVALID, "no read set bound to the plan");
    CKR(use_device(plan));
    cudaStream_t s = (cudaStream_t) stream;
    const uint32_t n = plan->R.n;
    if ((uint64_t) n * kSmallEdgesKept > 0xFFFFFFFFull) return fail(ALGA_E_INVALID, "too many reads for one GPU (%u)", n);
    plan->launches = 0;
    plan->spilled = 0;
    const int force = plan->params.flags & ALGA_PS_FORCE_GENERIC;
    Counters *dc = plan->counters_d.as<Counters>();
    if (plan->over_cap < kOverScanMax) plan->over_cap = n / 8 + 65536;
    // Equal-length reads at a fixed stride (the usual case): the kernels work on a copy whose read slots start on 32-byte
    // sector boundaries, so that a candidate read is fetched with one or two 32-byte requests instead of a dozen 4-byte
    // ones (common.cuh load8_na).  The copy is part of the timed pipeline: the caller's layout is the reference's.
    struct RestoreReads {
        alga_ps_plan *p;
        ReadsDev r;
        ~RestoreReads() { p->R = r; }
    } restore_reads{plan, plan->R};
    const bool repack = plan->P.uniform_len && !plan->R.word_off && n &&
                        (plan->R.stride != aligned_stride_words((plan->P.uniform_len + 15u) / 16u) || ((uintptr_t) plan->R.words & 31u));
    for (int attempt = 0;; attempt++) {
        CK(cudaEventRecord(plan->ev0, s));
        const bool sorted_index = sorted_stages() && n < 0x7FFFFFFFu;
        auto do_repack = [&]() -> int {
            if (!repack) return ALGA_OK;
            const uint32_t W = (plan->P.uniform_len + 15u) / 16u, S = aligned_stride_words(W);
            CKR(plan->slots.ensure((size_t) n * S * 4 + kReadPadBytes));
            launch_repack_reads(restore_reads.r.words, restore_reads.r.stride, W, n, plan->slots.as<uint32_t>(), S, s, plan->cfg);
            CK(cudaMemsetAsync(plan->slots.as<char>() + (size_t) n * S * 4, 0, kReadPadBytes, s));
            plan->R.words = plan->slots.as<uint32_t>();
            plan->R.stride = S;
            return ALGA_OK;
        };
        // seed index: the prefix table on this stream, the suffix table -- first needed by phase 2 -- on a side stream,
        // so its build overlaps phase 1
        if (sorted_index) {
            // the records of both tables in one pass over the reads in the CALLER's layout (36 instead of 2 x 64 bytes per read
            // of config 4), so that neither the suffix table nor the repack waits for the other
            size_t bp = 0, bs = 0;
            CKR(stage_index_size(plan, &bp, &bs));
            const size_t wsb = sorted_index_workspace_bytes(n);
            CKR(plan->sort_ws[0].ensure(wsb));
            CKR(plan->sort_ws[1].ensure(wsb));
            launch_seed_records(restore_reads.r, plan->P, plan->Tp, plan->Ts, n, plan->sort_ws[0].p, plan->sort_ws[1].p, s, plan->cfg);
            CK(cudaEventRecord(plan->ev_fork, s));
            CK(cudaStreamWaitEvent(plan->side, plan->ev_fork, 0));
            CK(cudaMemsetAsync(plan->ts.p, 0, bs, plan->side));
            if (launch_sorted_index(plan->Ts, n, plan->sort_ws[1].p, plan->side, plan->cfg)) return fail(ALGA_E_CUDA, "seed index: radix sort failed");
            CK(cudaMemsetAsync(plan->tp.p, 0, bp, s));
            CKR(do_repack());
            if (launch_sorted_index(plan->Tp, n, plan->sort_ws[0].p, s, plan->cfg)) return fail(ALGA_E_CUDA, "seed index: radix sort failed");
        } else {
            CKR(do_repack());
            CK(cudaEventRecord(plan->ev_fork, s));
            CK(cudaStreamWaitEvent(plan->side, plan->ev_fork, 0));
            CKR(stage_index_begin(plan, s, plan->side));
            launch_build_index(plan->R, plan->P, plan->Tp, plan->Ts, 0, n, 0u, 0xFFFFFFFFu, s, plan->cfg, 1);
            launch_build_index(plan->R, plan->P, plan->Tp, plan->Ts, 0, n, 0u, 0xFFFFFFFFu, plan->side, plan->cfg, 2);
        }
        CK(cudaGetLastError());
        CK(cudaEventRecord(plan->ev_join, plan->side));
        plan->index_valid = true;
        CK(cudaEventRecord(plan->ev_stage[0], s));
        // phase 1 writes every edge straight into the row of its target read (transposed graph, plan->row_cap per target)
        CKR(plan->rows.ensure((size_t) (n ? n : 1) * plan->row_cap * sizeof(RevEntry)));
        CKR(plan->over.ensure((size_t) plan->over_cap * sizeof(Edge1)));
        CKR(plan->indeg.ensure((size_t) (n ? n : 1) * 4));
        CKR(plan->rev_off.ensure(((size_t) n + 1) * 4));
        CKR(plan->rev.ensure((size_t) (n ? n : 1) * kSmallEdgesKept * sizeof(RevEntry)));
        CKR(plan->scan_ws.ensure(scan_workspace_bytes(n)));
        CK(cudaMemsetAsync(plan->indeg.p, 0, (size_t) (n ? n : 1) * 4, s));
        CK(cudaMemsetAsync(&dc->n_list, 0, 12, s));  // n_list, n_over, n_spill
        Phase1Out p1{0, plan->indeg.as<uint32_t>(), plan->rows.as<RevEntry>(), plan->row_cap, plan->over.as<Edge1>(), &dc->n_over,
                     plan->over_cap, 0u, ShardOut{}};
        CKR(run_phase1(plan, 0, n, p1, s));
        CK(cudaEventRecord(plan->ev_stage[1], s));
        // only when a row overflowed (decided on the device): CSR form of the transposed graph
        launch_rebuild_rows_csr(&dc->n_over, plan->over_cap, plan->over.as<Edge1>(), plan->indeg.as<uint32_t>(),
                                plan->rows.as<RevEntry>(), plan->row_cap, n, plan->rev_off.as<uint32_t>(), plan->rev.as<RevEntry>(),
                                (uint64_t) (n ? n : 1) * kSmallEdgesKept, &dc->rev_overflow, plan->scan_ws.p, s, plan->cfg);
        CK(cudaGetLastError());
        CK(cudaStreamWaitEvent(s, plan->ev_join, 0));  // the suffix table is complete
        CK(cudaEventRecord(plan->ev_stage[2], s));
        // phase 2 with fused out-degree counting (not in the reversed-result corner)
        CKR(plan->outdeg.ensure((size_t) (n ? n : 1) * 4));
        const bool fuse_outdeg = !plan->swap_direction && !sorted_stages();  // (the sorted CSR needs no row sizes)
        const RowsView view{plan->indeg.as<uint32_t>(), plan->rows.as<RevEntry>(), plan->row_cap, &dc->n_over,
                            plan->over.as<Edge1>(), plan->rev_off.as<uint32_t>(), plan->rev.as<RevEntry>()};
        CKR(run_phase2(plan, 0, n, view, fuse_outdeg ? plan->outdeg.as<uint32_t>() : nullptr, s));
        if (plan->h_counters->n_over > plan->over_cap) {  // the overflow list itself overflowed: size it for the worst case
            if (attempt) return fail(ALGA_E_CAPACITY, "phase-1 overflow list overflow persisted");
            plan->over_cap = n * (uint32_t) kSmallEdgesKept;
            continue;
        }
        CK(cudaEventRecord(plan->ev_stage[3], s));
        CKR(stage_csr(plan, 0, n, plan->triples.as<int32_t>(), plan->h_counters->n_edges, plan->swap_direction ? 1 : 0,
                      fuse_outdeg, s));
        break;
    }
    CK(cudaEventRecord(plan->ev1, s));
    CK(cudaEventSynchronize(plan->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, plan->ev0, plan->ev1));
    plan->last_device_ms = ms;
    cudaEvent_t marks[6] = {plan->ev0, plan->ev_stage[0], plan->ev_stage[1], plan->ev_stage[2], plan->ev_stage[3], plan->ev1};
    for (int k = 0; k < 5; k++) {
        CK(cudaEventElapsedTime(&ms, marks[k], marks[k + 1]));
        plan->stage_ms[k] = ms;
    }
    plan->stage_ms[5] = plan->h_counters->n_over;   // diagnostics: phase-1 edges beyond their row's capacity,
    plan->stage_ms[6] = plan->h_counters->n_hard1 >= kSecondPassMin ? plan->h_counters->n_hard1b : plan->h_counters->n_hard1;  // ... that took the generic phase-1 kernel
    return ALGA_OK;
}

int alga_ps_plan_result_device(alga_ps_plan *plan, const uint64_t **row_off, const int32_t **nbr, const int32_t **off,
                               uint64_t *n_edges) {
    if (!plan) return fail(ALGA_E_INVALID, "null plan");
    if (row_off) *row_off = plan->row_off.as<uint64_t>();
    if (nbr) *nbr = plan->nbr.as<int32_t>();
    if (off) *off = plan->off.as<int32_t>();
    if (n_edges) *n_edges = plan->n_edges;
    return ALGA_OK;
}

uint32_t alga_ps_plan_result_rows(alga_ps_plan *plan) { return plan ? plan->res_hi - plan->res_lo : 0u; }

int alga_ps_plan_result_host(alga_ps_plan *plan, alga_csr *out) {
    if (!plan || !out) return fail(ALGA_E_INVALID, "null argument");
    CKR(use_device(plan));
    const uint32_t n = plan->res_hi - plan->res_lo;
    const uint64_t E = plan->n_edges;
    out->n_reads = n;
    out->n_edges = E;
    out->borrowed = 0;
    out->row_off = (uint64_t *) malloc(((size_t) n + 1) * 8);
    out->nbr = (int32_t *) malloc((size_t) (E ? E : 1) * 4);
    out->off = (int32_t *) malloc((size_t) (E ? E : 1) * 4);
    if (!out->row_off || !out->nbr || !out->off) {
        alga_gpu_free_csr(out);
        return fail(ALGA_E_NOMEM, "out of host memory for %llu edges", (unsigned long long) E);
    }
    CK(cudaMemcpy(out->row_off, plan->row_off.p, ((size_t) n + 1) * 8, cudaMemcpyDeviceToHost));
    if (E) {
        CK(cudaMemcpy(out->nbr, plan->nbr.p, (size_t) E * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(out->off, plan->off.p, (size_t) E * 4, cudaMemcpyDeviceToHost));
    }
    return ALGA_OK;
}

int alga_ps_plan_result_host_pinned(alga_ps_plan *plan, alga_csr *out) {
    if (!plan || !out) return fail(ALGA_E_INVALID, "null argument");
    CKR(use_device(plan));
    const uint32_t n = plan->res_hi - plan->res_lo;
    const uint64_t E = plan->n_edges;
    CKR(plan->h_row_off.ensure(((size_t) n + 1) * 8));
    CKR(plan->h_nbr.ensure((size_t) (E ? E : 1) * 4));
    CKR(plan->h_off.ensure((size_t) (E ? E : 1) * 4));
    out->n_reads = n;
    out->n_edges = E;
    out->row_off = (uint64_t *) plan->h_row_off.p;
    out->nbr = (int32_t *) plan->h_nbr.p;
    out->off = (int32_t *) plan->h_off.p;
    out->borrowed = 1;
    cudaStream_t s = 0;
    CK(cudaMemcpyAsync(out->row_off, plan->row_off.p, ((size_t) n + 1) * 8, cudaMemcpyDeviceToHost, s));
    if (E) {
        CK(cudaMemcpyAsync(out->nbr, plan->nbr.p, (size_t) E * 4, cudaMemcpyDeviceToHost, s));
        CK(cudaMemcpyAsync(out->off, plan->off.p, (size_t) E * 4, cudaMemcpyDeviceToHost, s));
    }
    CK(cudaStreamSynchronize(s));
    return ALGA_OK;
}

void *alga_gpu_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        fail(ALGA_E_NOMEM, "cudaMallocHost of %zu bytes failed", bytes);
        return nullptr;
    }
    return p;
}
void alga_gpu_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int alga_ps_plan_stats(alga_ps_plan *plan, alga_timing *t) {
    if (!plan || !t) return fail(ALGA_E_INVALID, "null argument");
    memset(t, 0, sizeof(*t));
    t->device_ms = plan->last_device_ms;
    t->kernel_launches = plan->launches;
    t->n_spilled_targets = plan->spilled;
    for (int k = 0; k < 8; k++) t->stage_ms[k] = plan->stage_ms[k];
    return ALGA_OK;
}

void alga_gpu_free_csr(alga_csr *csr) {
    if (!csr) return;
    if (!csr->borrowed) {
        free(csr->row_off);
        free(csr->nbr);
        free(csr->off);
    }
    csr->borrowed = 0;
    csr->row_off = nullptr;
    csr->nbr = csr->off = nullptr;
    csr->n_edges = 0;
}

// One cached plan per process for the one-call entry point (the reference's graph creator is not
// re-entrant either: it works on process-wide statics).
static std::mutex g_build_mutex;
static alga_ps_plan *g_build_plan = nullptr;

int alga_gpu_prefsuf_build(const alga_reads *reads, const alga_ps_params *params, alga_csr *out, alga_timing *timing) {
    if (!reads || !params || !out) return fail(ALGA_E_INVALID, "null argument");
    std::lock_guard<std::mutex> lock(g_build_mutex);
    memset(out, 0, sizeof(*out));
    const double t0 = now_ms();
    if (g_build_plan && g_build_plan->params.device != params->device) {
        alga_ps_plan_destroy(g_build_plan);
        g_build_plan = nullptr;
    }
    if (!g_build_plan) CKR(alga_ps_plan_create(&g_build_plan, params));
    g_build_plan->params = *params;
    CKR(alga_ps_plan_upload_reads(g_build_plan, reads));
    const double t1 = now_ms();
    CKR(alga_ps_plan_run(g_build_plan, nullptr));
    const double t2 = now_ms();
    CKR(alga_ps_plan_result_host_pinned(g_build_plan, out));
    const double t3 = now_ms();
    if (timing) {
        alga_ps_plan_stats(g_build_plan, timing);
        timing->h2d_ms = t1 - t0;
        timing->d2h_ms = t3 - t2;
        timing->total_ms = t3 - t0;
    }
    return ALGA_OK;
}

// ---- fingerprints / pack / verify: host buffers in, host buffers out -------------------------------
namespace {
struct TmpReads {
    DevBuf words, word_off, len;
    ReadsDev R{};
    int upload(const alga_reads *h) {
        const uint32_t n = h->n_reads;
        if (n && (!h->words || !h->len_nt)) return fail(ALGA_E_INVALID, "words / len_nt must not be null");
        if (!h->word_off && h->stride_words == 0 && n) return fail(ALGA_E_INVALID, "word_off is null and stride_words is 0");
        const uint64_t n_words = h->word_off ? h->word_off[n] : (uint64_t) n * h->stride_words;
        CKR(words.ensure((size_t) n_words * 4 + kReadPadBytes));
        CKR(len.ensure((size_t) (n ? n : 1) * 4));
        CK(cudaMemcpy(words.p, h->words, (size_t) n_words * 4, cudaMemcpyHostToDevice));
        CK(cudaMemset((char *) words.p + (size_t) n_words * 4, 0, kReadPadBytes));
        CK(cudaMemcpy(len.p, h->len_nt, (size_t) n * 4, cudaMemcpyHostToDevice));
        R.words = words.as<uint32_t>();
        R.len = len.as<uint32_t>();
        R.n = n;
        R.stride = h->word_off ? 0 : h->stride_words;
        R.word_off = nullptr;
        if (h->word_off) {
            CKR(word_off.ensure(((size_t) n + 1) * 8));
            CK(cudaMemcpy(word_off.p, h->word_off, ((size_t) n + 1) * 8, cudaMemcpyHostToDevice));
            R.word_off = word_off.as<uint64_t>();
        }
        return ALGA_OK;
    }
    ~TmpReads() {
        words.release();
        word_off.release();
        len.release();
    }
};

int pick_device(int device, LaunchCfg *cfg) {
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        return fail(ALGA_E_CUDA, "no CUDA device available; libalga_gpu has no CPU fallback");
    }
    if (device < 0 || device >= ndev) return fail(ALGA_E_INVALID, "device %d out of range", device);
    CK(cudaSetDevice(device));
    CK(cudaDeviceGetAttribute(&cfg->sm_count, cudaDevAttrMultiProcessorCount, device));
    return ALGA_OK;
}
}  // namespace

int alga_gpu_fingerprints(const alga_reads *reads, int32_t L, int32_t device, uint64_t *pre64, uint32_t *pre32,
                          uint64_t *suf64, uint32_t *suf32) {
    if (!reads || !pre64 || !pre32 || !suf64 || !suf32) return fail(ALGA_E_INVALID, "null argument");
    LaunchCfg cfg;
    CKR(pick_device(device, &cfg));
    TmpReads t;
    CKR(t.upload(reads));
    const uint32_t n = reads->n_reads;
    DevBuf a, b, c, d;
    int r = [&]() -> int {
        CKR(a.ensure((size_t) (n ? n : 1) * 8));
        CKR(b.ensure((size_t) (n ? n : 1) * 4));
        CKR(c.ensure((size_t) (n ? n : 1) * 8));
        CKR(d.ensure((size_t) (n ? n : 1) * 4));
        CK(cudaMemcpy(a.p, pre64, (size_t) n * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(b.p, pre32, (size_t) n * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c.p, suf64, (size_t) n * 8, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(d.p, suf32, (size_t) n * 4, cudaMemcpyHostToDevice));
        launch_fingerprints(t.R, L, a.as<uint64_t>(), b.as<uint32_t>(), c.as<uint64_t>(), d.as<uint32_t>(), 0, cfg);
        CK(cudaGetLastError());
        CK(cudaMemcpy(pre64, a.p, (size_t) n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(pre32, b.p, (size_t) n * 4, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(suf64, c.p, (size_t) n * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(suf32, d.p, (size_t) n * 4, cudaMemcpyDeviceToHost));
        return ALGA_OK;
    }();
    a.release();
    b.release();
    c.release();
    d.release();
    return r;
}

int alga_gpu_pack_reads(const uint8_t *ascii, uint32_t n_reads, uint32_t len_nt, int32_t device, uint32_t *words) {
    if (n_reads && len_nt && (!ascii || !words)) return fail(ALGA_E_INVALID, "null argument");
    if (len_nt > 16 * 1024) return fail(ALGA_E_INVALID, "len_nt %u too large for the packing kernel", len_nt);
    LaunchCfg cfg;
    CKR(pick_device(device, &cfg));
    const size_t nbytes = (size_t) n_reads * len_nt;
    const size_t nwords = (size_t) n_reads * ((len_nt + 15) / 16);
    DevBuf in, out;
    int r = [&]() -> int {
        CKR(in.ensure(nbytes + 32));
        CKR(out.ensure((nwords ? nwords : 1) * 4));
        CK(cudaMemcpy(in.p, ascii, nbytes, cudaMemcpyHostToDevice));
        launch_pack_reads(in.as<uint8_t>(), n_reads, len_nt, out.as<uint32_t>(), 0, cfg);
        CK(cudaGetLastError());
        CK(cudaMemcpy(words, out.p, nwords * 4, cudaMemcpyDeviceToHost));
        return ALGA_OK;
    }();
    in.release();
    out.release();
    return r;
}

int alga_gpu_verify_pairs(const alga_reads *reads, const int32_t *pairs, uint64_t n_pairs,
                          const alga_verify_params *params, uint8_t *verdict) {
    if (!reads || !params || (n_pairs && (!pairs || !verdict))) return fail(ALGA_E_INVALID, "null argument");
    if (params->lcs_rate_pct > 0 && (params->lcs_band < 0 || params->lcs_band > 8))
        return fail(ALGA_E_INVALID, "lcs_band must be 0 .. 8 (got %d)", params->lcs_band);
    LaunchCfg cfg;
    CKR(pick_device(params->device, &cfg));
    TmpReads t;
    CKR(t.upload(reads));
    DevBuf dp, dv;
    int r = [&]() -> int {
        CKR(dp.ensure((size_t) (n_pairs ? n_pairs : 1) * 12));
        CKR(dv.ensure((size_t) (n_pairs ? n_pairs : 1)));
        CK(cudaMemcpy(dp.p, pairs, (size_t) n_pairs * 12, cudaMemcpyHostToDevice));
        VerifyDev V{params->max_offset_pct, params->min_offset, params->min_overlap_area, params->threshold_pct,
                    params->same_ends, params->lcs_rate_pct, params->lcs_band};
        launch_verify_pairs(t.R, dp.as<int32_t>(), n_pairs, V, dv.as<uint8_t>(), 0, cfg);
        CK(cudaGetLastError());
        CK(cudaMemcpy(verdict, dv.p, (size_t) n_pairs, cudaMemcpyDeviceToHost));
        return ALGA_OK;
    }();
    dp.release();
    dv.release();
    return r;
}

int alga_gpu_prefix_reads(const alga_reads *reads, int32_t remove_type, int32_t device, uint8_t *mask, alga_timing *timing) {
    if (!reads || (reads->n_reads && !mask)) return fail(ALGA_E_INVALID, "null argument");
    if (remove_type != 1 && remove_type != 2) return fail(ALGA_E_INVALID, "remove_type must be 1 (duplicates) or 2 (all prefix reads)");
    const double t0 = now_ms();
    LaunchCfg cfg;
    uint64_t launches = 0;
    cfg.launches = &launches;
    CKR(pick_device(device, &cfg));
    TmpReads t;
    CKR(t.upload(reads));
    const double t1 = now_ms();
    const uint32_t n = reads->n_reads;
    DevBuf table, lenmap, flags, dmask;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    float dev_ms = 0;
    int r = [&]() -> int {
        SeedTable T{};
        size_table(T, n);
        const size_t tb = (size_t) T.n_buckets * kBucketWords * 4;
        CKR(table.ensure(tb));
        CKR(lenmap.ensure(prefix_reads_lenmap_words() * 4));
        CKR(flags.ensure((size_t) (n ? n : 1) * 4));
        CKR(dmask.ensure(n ? n : 1));
        T.slots = table.as<uint32_t>();
        CK(cudaEventCreate(&e0));
        CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0, 0));
        CK(cudaMemsetAsync(table.p, 0, tb, 0));
        launch_prefix_reads(t.R, T, remove_type, lenmap.as<uint32_t>(), flags.as<uint32_t>(), dmask.as<uint8_t>(), 0, cfg);
        CK(cudaGetLastError());
        CK(cudaEventRecord(e1, 0));
        uint32_t too_long = 0;
        CK(cudaMemcpy(&too_long, lenmap.as<uint32_t>() + prefix_reads_lenmap_words() - 1, 4, cudaMemcpyDeviceToHost));
        if (too_long) return fail(ALGA_E_INVALID, "a read is longer than 65535 nucleotides");
        if (n) CK(cudaMemcpy(mask, dmask.p, n, cudaMemcpyDeviceToHost));
        CK(cudaEventElapsedTime(&dev_ms, e0, e1));
        return ALGA_OK;
    }();
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    table.release();
    lenmap.release();
    flags.release();
    dmask.release();
    if (r == ALGA_OK && timing) {
        memset(timing, 0, sizeof(*timing));
        timing->h2d_ms = t1 - t0;
        timing->device_ms = dev_ms;
        timing->total_ms = now_ms() - t0;
        timing->kernel_launches = launches;
    }
    return r;
}

int alga_gpu_supplement(const alga_reads *reads, const alga_csr *graph_in, const alga_sup_params *params,
                        alga_csr *graph_out, alga_timing *timing) {
    if (!reads || !graph_in || !params || !graph_out) return fail(ALGA_E_INVALID, "null argument");
    if (reads->n_reads && (!reads->words || !reads->len_nt)) return fail(ALGA_E_INVALID, "words / len_nt must not be null");
    if (!reads->word_off && reads->stride_words == 0 && reads->n_reads) return fail(ALGA_E_INVALID, "word_off is null and stride_words is 0");
    if (reads->n_reads && (!graph_in->row_off || (graph_in->n_edges && (!graph_in->nbr || !graph_in->off))))
        return fail(ALGA_E_INVALID, "graph_in arrays must not be null");
    const int rc = supplement_impl(reads, graph_in, params, graph_out, timing);
    if (rc != ALGA_OK) return fail(rc, "%s", supplement_last_error());
    return ALGA_OK;
}

int alga_gpu_li_kmers(const alga_reads *reads, const uint32_t *ids, uint32_t n_ids, const int32_t priorities[4],
                      int32_t kmer_length, int32_t intervals, int32_t device, uint64_t *hash_out, int32_t *ind_out) {
    if (!reads || !priorities || (n_ids && (!ids || !hash_out || !ind_out))) return fail(ALGA_E_INVALID, "null argument");
    if (kmer_length < 1 || kmer_length > 63 || intervals < 1 || intervals > 64)
        return fail(ALGA_E_INVALID, "kmer_length must be 1..63 and intervals 1..64");
    for (uint32_t i = 0; i < n_ids; i++)
        if (ids[i] >= reads->n_reads || (int64_t) reads->len_nt[ids[i]] < kmer_length)
            return fail(ALGA_E_INVALID, "read %u is missing or shorter than the k-mer", ids[i]);
    LaunchCfg cfg;
    CKR(pick_device(device, &cfg));
    TmpReads t;
    CKR(t.upload(reads));
    DevBuf di, dh, dn;
    int r = [&]() -> int {
        CKR(di.ensure((size_t) (n_ids ? n_ids : 1) * 4));
        CKR(dh.ensure((size_t) (n_ids ? n_ids : 1) * intervals * 8));
        CKR(dn.ensure((size_t) (n_ids ? n_ids : 1) * intervals * 4));
        CK(cudaMemcpy(di.p, ids, (size_t) n_ids * 4, cudaMemcpyHostToDevice));
        const int rc = run_li_kmers(t.R, di.as<uint32_t>(), n_ids, priorities, kmer_length, intervals, dh.as<uint64_t>(),
                                    dn.as<int32_t>(), 0, cfg);
        if (rc != ALGA_OK) return fail(rc, "%s", supplement_last_error());
        CK(cudaMemcpy(hash_out, dh.p, (size_t) n_ids * intervals * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(ind_out, dn.p, (size_t) n_ids * intervals * 4, cudaMemcpyDeviceToHost));
        return ALGA_OK;
    }();
    di.release();
    dh.release();
    dn.release();
    return r;
}

}  // extern "C"

// ---- InputReader::readInput / renumbering / files-to-graph (input.cu) ------------------------------------------------
namespace {

struct InputScalarsHost {  // layout of input.cu's InputScalars
    uint32_t first_empty, first_bad, max_len, pad;
    unsigned long long n_with_n, n_str, sum_len, n_alive;
};

struct InputFile {
    DevBuf text, block_cnt, block_off, rec_start, rec_end, info, scalars;
    uint64_t n = 0;
    uint32_t n_rec = 0;
    InputScalarsHost sc{};
    void release() { text.release(), block_cnt.release(), block_off.release(), rec_start.release(), rec_end.release(), info.release(), scalars.release(); }
};

// Device-resident working set of the input stage.  One process-wide instance is kept between calls (as the one-call graph
// build keeps its plan): after the first call nothing is allocated any more.
struct FrontEnd {
    int device = -1;
    InputFile f[2];
    HostBuf stage[2];  // chunked upload of text that is not page-locked
    cudaEvent_t stage_ev[2] = {nullptr, nullptr};
    bool stage_used[2] = {false, false};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    // copies run on their own (non-blocking) stream: the second file arrives while the first is scanned, and the read
    // set goes back to the host while the graph is built
    cudaStream_t copy = nullptr;
    cudaEvent_t ev_up0 = nullptr, ev_up[2] = {nullptr, nullptr}, ev_ready = nullptr;
    // reader output: Global::READS after InputReader::readInput
    DevBuf words, len;
    uint32_t n = 0, stride = 1, max_len = 0;
    // ReadPreprocess
    DevBuf table, lenmap, flags, mask;
    // renumbering
    DevBuf uflag, upos, rscal, scan_ws;
    DevBuf words2, len2, old_id, po;
    uint32_t n2 = 0, stride2 = 1, max_len2 = 0;
    // page-locked result staging (borrowed outputs)
    HostBuf h_words, h_len, h_words2, h_len2, h_old, h_po;

    ReadsDev raw() const {
        ReadsDev R{};
        R.words = words.as<uint32_t>(), R.len = len.as<uint32_t>(), R.n = n, R.stride = stride;
        return R;
    }
    int init(int dev) {
        if (device != dev) release();
        device = dev;
        for (cudaEvent_t *e : {&stage_ev[0], &stage_ev[1]})
            if (!*e) CK(cudaEventCreateWithFlags(e, cudaEventDisableTiming));
        for (cudaEvent_t *e : {&ev[0], &ev[1], &ev_up0, &ev_up[0], &ev_up[1]})
            if (!*e) CK(cudaEventCreate(e));
        if (!ev_ready) CK(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
        if (!copy) CK(cudaStreamCreateWithFlags(&copy, cudaStreamNonBlocking));
        return ALGA_OK;
    }
    void release() {
        f[0].release(), f[1].release();
        for (DevBuf *b : {&words, &len, &table, &lenmap, &flags, &mask, &uflag, &upos, &rscal, &scan_ws, &words2, &len2, &old_id, &po})
            b->release();
        for (HostBuf *b : {&stage[0], &stage[1], &h_words, &h_len, &h_words2, &h_len2, &h_old, &h_po}) b->release();
        for (cudaEvent_t *e : {&stage_ev[0], &stage_ev[1], &ev[0], &ev[1], &ev_up0, &ev_up[0], &ev_up[1], &ev_ready}) {
            if (*e) cudaEventDestroy(*e);
            *e = nullptr;
        }
        if (copy) cudaStreamDestroy(copy);
        copy = nullptr;
        stage_used[0] = stage_used[1] = false;
    }
};

std::mutex g_front_mutex;
FrontEnd g_front;

constexpr size_t kStageChunk = 8u << 20;

// host -> device copy of file text: directly if the caller's buffer is page-locked (alga_gpu_host_alloc), otherwise
// through two page-locked chunks so that the CPU copy of one chunk overlaps the DMA of the previous one
int fe_upload_text(FrontEnd &fe, DevBuf &dst, const uint8_t *src, uint64_t n) {
    cudaStream_t s = fe.copy;
    const size_t padded = (size_t) ((n + 15) / 16) * 16 + 16;
    CKR(dst.ensure(padded));
    CK(cudaMemsetAsync((char *) dst.p + n, 0x0A, padded - (size_t) n, s));  // the kernels never look past n; keep the pad defined
    if (!n) return ALGA_OK;
    cudaPointerAttributes attr{};
    const bool pinned = cudaPointerGetAttributes(&attr, src) == cudaSuccess && attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (pinned) {
        CK(cudaMemcpyAsync(dst.p, src, (size_t) n, cudaMemcpyHostToDevice, s));
        return ALGA_OK;
    }
    CKR(fe.stage[0].ensure(kStageChunk));
    CKR(fe.stage[1].ensure(kStageChunk));
    int k = 0;
    for (uint64_t off = 0; off < n; off += kStageChunk, k ^= 1) {
        const size_t len = (size_t) (n - off < kStageChunk ? n - off : kStageChunk);
        if (fe.stage_used[k]) CK(cudaEventSynchronize(fe.stage_ev[k]));  // the DMA out of this chunk must be done
        fe.stage_used[k] = true;
        memcpy(fe.stage[k].p, src + off, len);
        CK(cudaMemcpyAsync((char *) dst.p + off, fe.stage[k].p, len, cudaMemcpyHostToDevice, s));
        CK(cudaEventRecord(fe.stage_ev[k], s));
    }
    return ALGA_OK;
}

// record index + per-record scan of one file whose text is on the device; leaves f.n_rec, f.sc and f.info
int fe_scan_file(FrontEnd &fe, InputFile &f, uint64_t n, const alga_input_params &p, int which, const LaunchCfg &cfg) {
    const bool plain = p.file_type == ALGA_INPUT_PLAIN;
    const uint32_t lpr = p.file_type == ALGA_INPUT_FASTQ ? 4u : 2u;
    f.n = n;
    const uint64_t nb = input_mark_blocks(n);
    if (nb > 0x7FFFFFFFull) return fail(ALGA_E_INVALID, "input file %d is too large (%llu bytes)", which, (unsigned long long) n);
    CKR(f.scalars.ensure(input_scalars_bytes()));
    InputScalarsHost init{};
    init.first_empty = init.first_bad = 0xFFFFFFFFu;
    CK(cudaMemcpyAsync(f.scalars.p, &init, sizeof(init), cudaMemcpyHostToDevice, 0));
    uint64_t n_marks = 0;
    if (nb) {
        CKR(f.block_cnt.ensure((size_t) nb * 4));
        CKR(f.block_off.ensure(((size_t) nb + 1) * 8));
        CKR(fe.scan_ws.ensure(scan_workspace_bytes(nb)));
        launch_count_marks(f.text.as<uint8_t>(), n, plain, f.block_cnt.as<uint32_t>(), 0, cfg);
        launch_scan_u64(f.block_cnt.as<uint32_t>(), f.block_off.as<uint64_t>(), nb, fe.scan_ws.p, 0, cfg);
        CK(cudaGetLastError());
        CK(cudaMemcpy(&n_marks, f.block_off.as<uint64_t>() + nb, 8, cudaMemcpyDeviceToHost));
    }
    const uint64_t n_cand = plain ? n_marks : (n_marks + lpr - 1) / lpr;
    if (n_cand > 0x1FFFFFFFull) return fail(ALGA_E_INVALID, "input file %d holds too many records (%llu)", which, (unsigned long long) n_cand);
    f.n_rec = 0;
    f.sc = init;
    if (!n_cand) return ALGA_OK;
    CKR(f.rec_start.ensure((size_t) n_cand * 8));
    CKR(f.rec_end.ensure((size_t) n_cand * 8));
    CKR(f.info.ensure((size_t) n_cand * input_rec_info_bytes()));
    launch_write_marks(f.text.as<uint8_t>(), n, plain, f.block_off.as<uint64_t>(), lpr, f.rec_start.as<uint64_t>(),
                       f.rec_end.as<uint64_t>(), n_cand, 0, cfg);
    launch_scan_records(f.text.as<uint8_t>(), n, plain, f.rec_start.as<uint64_t>(), f.rec_end.as<uint64_t>(), n_marks, lpr,
                        (uint32_t) n_cand, p.trim_left, p.trim_right, p.rna != 0, p.str_threshold > 0 ? p.str_threshold : 20, f.info.p,
                        f.scalars.p, 0, cfg);
    CK(cudaGetLastError());
    CK(cudaMemcpy(&f.sc, f.scalars.p, sizeof(f.sc), cudaMemcpyDeviceToHost));
    f.n_rec = f.sc.first_empty < n_cand ? f.sc.first_empty : (uint32_t) n_cand;
    if (f.sc.first_bad < f.n_rec)
        return fail(ALGA_E_INVALID, "record %u of input file %d holds a character other than A, C, G, T, N, U", f.sc.first_bad, which);
    launch_record_totals(f.info.p, f.n_rec, f.scalars.p, 0, cfg);
    CK(cudaGetLastError());
    CK(cudaMemcpy(&f.sc, f.scalars.p, sizeof(f.sc), cudaMemcpyDeviceToHost));
    return ALGA_OK;
}

struct InputTimes {
    double h2d_ms = 0, kernel_ms = 0;
};

// InputReader::readInput: file text (host) -> fe.words / fe.len (device); fills the counters of `info`
int fe_read_input_impl(FrontEnd &fe, const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2, const alga_input_params &p,
                       const LaunchCfg &cfg, alga_read_set *info, InputTimes *tm) {
    const bool paired = text2 != nullptr;
    const double t0 = now_ms();
    CK(cudaEventRecord(fe.ev_up0, fe.copy));
    CKR(fe_upload_text(fe, fe.f[0].text, text1, n1));
    CK(cudaEventRecord(fe.ev_up[0], fe.copy));
    if (paired) {
        CKR(fe_upload_text(fe, fe.f[1].text, text2, n2));
        CK(cudaEventRecord(fe.ev_up[1], fe.copy));
    }
    CK(cudaStreamWaitEvent(0, fe.ev_up[0], 0));
    CKR(fe_scan_file(fe, fe.f[0], n1, p, 1, cfg));  // runs while the second file is still arriving
    if (paired) {
        CK(cudaStreamWaitEvent(0, fe.ev_up[1], 0));
        CKR(fe_scan_file(fe, fe.f[1], n2, p, 2, cfg));
        if (fe.f[0].n_rec != fe.f[1].n_rec)
            return fail(ALGA_E_INVALID, "the mate files hold different numbers of records (%u and %u)", fe.f[0].n_rec, fe.f[1].n_rec);
    }
    const uint64_t n_reads = (uint64_t) fe.f[0].n_rec * (paired ? 4 : 2);
    if (n_reads > 0x7FFFFFFFull) return fail(ALGA_E_INVALID, "too many reads (%llu): ids are 31-bit", (unsigned long long) n_reads);
    uint32_t max_len = fe.f[0].sc.max_len;
    if (paired && fe.f[1].sc.max_len > max_len) max_len = fe.f[1].sc.max_len;
    fe.n = (uint32_t) n_reads;
    fe.max_len = max_len;
    fe.stride = max_len ? (max_len + 15) / 16 : 1;
    info->n_reads = fe.n;
    info->stride_words = fe.stride;
    info->max_len_nt = max_len;
    info->n_records[0] = fe.f[0].n_rec;
    info->n_records[1] = paired ? fe.f[1].n_rec : 0;
    info->n_with_n = fe.f[0].sc.n_with_n + (paired ? fe.f[1].sc.n_with_n : 0);
    info->n_str = fe.f[0].sc.n_str + (paired ? fe.f[1].sc.n_str : 0);
    CKR(fe.words.ensure((size_t) (n_reads ? n_reads : 1) * fe.stride * 4));
    CKR(fe.len.ensure((size_t) (n_reads ? n_reads : 1) * 4));
    for (int k = 0; k < (paired ? 2 : 1); k++)
        launch_pack_records(fe.f[k].text.as<uint8_t>(), fe.f[k].info.p, fe.f[k].n_rec, p.rna != 0, paired ? 4u : 2u, 2u * k, fe.stride,
                            fe.words.as<uint32_t>(), fe.len.as<uint32_t>(), 0, cfg);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(0));
    float up_ms = 0;
    CK(cudaEventElapsedTime(&up_ms, fe.ev_up0, fe.ev_up[paired ? 1 : 0]));
    tm->h2d_ms = up_ms;                 // the uploads on the copy stream (for page-locked text: pure DMA time)
    tm->kernel_ms = now_ms() - t0;      // wall: call -> packed reads resident, uploads included (they overlap the scan)
    return ALGA_OK;
}

int fe_read_input(FrontEnd &fe, const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2, const alga_input_params &p,
                  const LaunchCfg &cfg, alga_read_set *info, InputTimes *tm) {
    const int r = fe_read_input_impl(fe, text1, n1, text2, n2, p, cfg, info, tm);
    if (r != ALGA_OK) cudaStreamSynchronize(fe.copy);  // nothing may still be reading the caller's buffers
    return r;
}

// main.cpp:150-232 on a device-resident read set: R + mask (device, may be null) -> fe.words2 / len2 / old_id / po.
// stride_out == 0: ceil(longest surviving read / 16).  Reads shorter than min_keep_len become nullptr (length 0).
int fe_remap(FrontEnd &fe, const ReadsDev &R, const uint8_t *d_mask, uint32_t stride_out, uint32_t min_keep_len, const LaunchCfg &cfg) {
    const uint32_t n_units = R.n / 2;
    fe.n2 = 0, fe.stride2 = stride_out ? stride_out : 1, fe.max_len2 = 0;
    if (!n_units) return ALGA_OK;
    CKR(fe.uflag.ensure((size_t) n_units * 4));
    CKR(fe.upos.ensure(((size_t) n_units + 1) * 4));
    CKR(fe.rscal.ensure(8));
    CKR(fe.scan_ws.ensure(scan_workspace_bytes(n_units)));
    CK(cudaMemsetAsync(fe.rscal.p, 0, 8, 0));
    launch_remap_flags(R, d_mask, n_units, fe.uflag.as<uint32_t>(), fe.rscal.p, 0, cfg);
    launch_scan_u32(fe.uflag.as<uint32_t>(), fe.upos.as<uint32_t>(), n_units, fe.scan_ws.p, 0, cfg);
    CK(cudaGetLastError());
    uint32_t sc[2] = {0, 0}, units_out = 0;
    CK(cudaMemcpy(sc, fe.rscal.p, 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(&units_out, fe.upos.as<uint32_t>() + n_units, 4, cudaMemcpyDeviceToHost));
    if (sc[1]) return fail(ALGA_E_INVALID, "read %u is present without its reverse complement (main.cpp:173 asserts)", sc[1] - 1);
    fe.n2 = 2 * units_out;
    fe.max_len2 = sc[0];
    if (!stride_out) fe.stride2 = sc[0] ? (sc[0] + 15) / 16 : 1;
    if (!fe.n2) return ALGA_OK;
    CKR(fe.words2.ensure((size_t) fe.n2 * fe.stride2 * 4));
    CKR(fe.len2.ensure((size_t) fe.n2 * 4));
    CKR(fe.old_id.ensure((size_t) fe.n2 * 4));
    CKR(fe.po.ensure(fe.n2));
    launch_remap_scatter(R, n_units, fe.uflag.as<uint32_t>(), fe.upos.as<uint32_t>(), fe.stride2, fe.words2.as<uint32_t>(),
                         fe.len2.as<uint32_t>(), fe.old_id.as<uint32_t>(), fe.po.as<uint8_t>(), min_keep_len, 0, cfg);
    CK(cudaGetLastError());
    return ALGA_OK;
}

// ReadPreprocess::getPrefixReads on the reader's output -> fe.mask (device)
int fe_prefix_reads(FrontEnd &fe, int remove_type, const LaunchCfg &cfg) {
    const uint32_t n = fe.n;
    SeedTable T{};
    size_table(T, n);
    const size_t tb = (size_t) T.n_buckets * kBucketWords * 4;
    CKR(fe.table.ensure(tb));
    CKR(fe.lenmap.ensure(prefix_reads_lenmap_words() * 4));
    CKR(fe.flags.ensure((size_t) (n ? n : 1) * 4));
    CKR(fe.mask.ensure(n ? n : 1));
    T.slots = fe.table.as<uint32_t>();
    CK(cudaMemsetAsync(fe.table.p, 0, tb, 0));
    launch_prefix_reads(fe.raw(), T, remove_type, fe.lenmap.as<uint32_t>(), fe.flags.as<uint32_t>(), fe.mask.as<uint8_t>(), 0, cfg);
    CK(cudaGetLastError());
    uint32_t too_long = 0;
    CK(cudaMemcpy(&too_long, fe.lenmap.as<uint32_t>() + prefix_reads_lenmap_words() - 1, 4, cudaMemcpyDeviceToHost));
    if (too_long) return fail(ALGA_E_INVALID, "a read is longer than 65535 nucleotides");
    return ALGA_OK;
}

void clear_read_set(alga_read_set *rs) { memset(rs, 0, sizeof(*rs)); }

// device -> page-locked staging of the library; the arrays stay valid until the next call that produces the same kind of
// read set (out->borrowed = 1)
// s == fe.copy with wait == false: the copies are only queued (behind fe.ev_ready); the caller synchronises fe.copy.
// cfg != nullptr: the bulk goes through a copy kernel instead of the DMA engine (see launch_copy_to_host).
int fe_download(FrontEnd &fe, bool remapped, alga_read_set *out, cudaStream_t s = 0, bool wait = true, const LaunchCfg *cfg = nullptr) {
    const uint32_t n = remapped ? fe.n2 : fe.n, stride = remapped ? fe.stride2 : fe.stride;
    HostBuf &hw = remapped ? fe.h_words2 : fe.h_words, &hl = remapped ? fe.h_len2 : fe.h_len;
    const size_t wb = (size_t) n * stride * 4;
    CKR(hw.ensure(wb ? wb : 4));
    CKR(hl.ensure(n ? (size_t) n * 4 : 4));
    out->n_reads = n;
    out->stride_words = stride;
    out->max_len_nt = remapped ? fe.max_len2 : fe.max_len;
    out->words = (uint32_t *) hw.p;
    out->len_nt = (uint32_t *) hl.p;
    out->borrowed = 1;
    auto copy = [&](void *dst, const void *src, size_t bytes) -> int {
        if (!bytes) return ALGA_OK;
        if (cfg) launch_copy_to_host(dst, src, bytes, s, *cfg);
        else CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
        return ALGA_OK;
    };
    CKR(copy(hw.p, remapped ? fe.words2.p : fe.words.p, wb));
    CKR(copy(hl.p, remapped ? fe.len2.p : fe.len.p, (size_t) n * 4));
    if (remapped) {
        CKR(fe.h_old.ensure(n ? (size_t) n * 4 : 4));
        CKR(fe.h_po.ensure(n ? n : 4));
        out->old_id = (uint32_t *) fe.h_old.p;
        out->paired_offset = (uint8_t *) fe.h_po.p;
        CKR(copy(fe.h_old.p, fe.old_id.p, (size_t) n * 4));
        CKR(copy(fe.h_po.p, fe.po.p, n));
    }
    CK(cudaGetLastError());
    if (wait) CK(cudaStreamSynchronize(s));
    return ALGA_OK;
}

int check_input_params(const alga_input_params *p) {
    if (p->file_type < ALGA_INPUT_PLAIN || p->file_type > ALGA_INPUT_FASTQ)
        return fail(ALGA_E_INVALID, "file_type must be ALGA_INPUT_PLAIN, _FASTA or _FASTQ");
    if (p->trim_left < 0 || p->trim_right < 0) return fail(ALGA_E_INVALID, "trim_left / trim_right must not be negative");
    return ALGA_OK;
}

}  // namespace

void alga_gpu_free_read_set(alga_read_set *rs) {
    if (!rs) return;
    if (!rs->borrowed) {
        free(rs->words);
        free(rs->len_nt);
        free(rs->old_id);
        free(rs->paired_offset);
    }
    clear_read_set(rs);
}

int alga_gpu_read_input(const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2, const alga_input_params *params,
                        alga_read_set *out, alga_timing *timing) {
    if (!params || !out || (n1 && !text1) || (n2 && !text2)) return fail(ALGA_E_INVALID, "null argument");
    CKR(check_input_params(params));
    std::lock_guard<std::mutex> lock(g_front_mutex);
    clear_read_set(out);
    const double t0 = now_ms();
    LaunchCfg cfg;
    uint64_t launches = 0;
    cfg.launches = &launches;
    CKR(pick_device(params->device, &cfg));
    FrontEnd &fe = g_front;
    CKR(fe.init(params->device));
    InputTimes tm;
    int r = fe_read_input(fe, text1, n1, text2, n2, *params, cfg, out, &tm);
    const double t1 = now_ms();
    if (r == ALGA_OK) r = fe_download(fe, false, out);
    if (r != ALGA_OK) {
        clear_read_set(out);
        return r;
    }
    if (timing) {
        memset(timing, 0, sizeof(*timing));
        timing->h2d_ms = tm.h2d_ms;        // upload on the copy stream (overlaps the scan of the first file)
        timing->device_ms = tm.kernel_ms;  // wall: call -> packed reads resident, uploads included
        timing->d2h_ms = now_ms() - t1;
        timing->total_ms = now_ms() - t0;
        timing->kernel_launches = launches;
    }
    return ALGA_OK;
}

int alga_gpu_remap_reads(const alga_reads *reads, const uint8_t *remove_mask, int32_t device, alga_read_set *out,
                         alga_timing *timing) {
    if (!reads || !out) return fail(ALGA_E_INVALID, "null argument");
    if (reads->n_reads & 1u) return fail(ALGA_E_INVALID, "n_reads must be even (both strands of every record)");
    std::lock_guard<std::mutex> lock(g_front_mutex);
    clear_read_set(out);
    const double t0 = now_ms();
    LaunchCfg cfg;
    uint64_t launches = 0;
    cfg.launches = &launches;
    CKR(pick_device(device, &cfg));
    FrontEnd &fe = g_front;
    CKR(fe.init(device));
    TmpReads t;
    CKR(t.upload(reads));
    const uint32_t n = reads->n_reads;
    if (remove_mask && n) {
        CKR(fe.mask.ensure(n));
        CK(cudaMemcpy(fe.mask.p, remove_mask, n, cudaMemcpyHostToDevice));
    }
    const double t1 = now_ms();
    int r = fe_remap(fe, t.R, remove_mask && n ? fe.mask.as<uint8_t>() : nullptr, 0, 0, cfg);
    if (r == ALGA_OK) r = cudaStreamSynchronize(0) == cudaSuccess ? ALGA_OK : fail(ALGA_E_CUDA, "remap kernels failed: %s", cudaGetErrorString(cudaGetLastError()));
    const double t2 = now_ms();
    if (r == ALGA_OK) r = fe_download(fe, true, out);
    if (r != ALGA_OK) {
        clear_read_set(out);
        return r;
    }
    if (timing) {
        memset(timing, 0, sizeof(*timing));
        timing->h2d_ms = t1 - t0;
        timing->device_ms = t2 - t1;
        timing->d2h_ms = now_ms() - t2;
        timing->total_ms = now_ms() - t0;
        timing->kernel_launches = launches;
    }
    return ALGA_OK;
}

// main.cpp:82-291 in one call, device-resident between the stages
int alga_gpu_files_to_graph(const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2, const alga_driver_params *dp,
                            alga_overlap_graph *out, alga_timing *timing) {
    if (!dp || !out || (n1 && !text1) || (n2 && !text2)) return fail(ALGA_E_INVALID, "null argument");
    CKR(check_input_params(&dp->input));
    if (dp->remove_type < 0 || dp->remove_type > 2) return fail(ALGA_E_INVALID, "remove_type must be 0 (none), 1 (duplicates) or 2 (all prefix reads)");
    std::lock_guard<std::mutex> lock(g_front_mutex);
    std::lock_guard<std::mutex> lock2(g_build_mutex);
    memset(out, 0, sizeof(*out));
    const double t0 = now_ms();
    LaunchCfg cfg;
    uint64_t launches = 0;
    cfg.launches = &launches;
    CKR(pick_device(dp->input.device, &cfg));
    FrontEnd &fe = g_front;
    CKR(fe.init(dp->input.device));
    InputTimes tm;
    alga_read_set info{};
    CKR(fe_read_input(fe, text1, n1, text2, n2, dp->input, cfg, &info, &tm));
    const double t1 = now_ms();
    // main.cpp:93-110 -- parameters from the average read length (float arithmetic as in the reference)
    const unsigned long long cnt = fe.f[0].sc.n_alive + (text2 ? fe.f[1].sc.n_alive : 0);
    const unsigned long long sum = fe.f[0].sc.sum_len + (text2 ? fe.f[1].sc.sum_len : 0);
    const double avg = cnt ? (double) sum / (double) cnt : 0.0;
    const float scale = dp->scale > 0 ? dp->scale : 0.55f;
    const int LEN = (int) (avg + dp->input.trim_left + dp->input.trim_right);
    const int L = (int) ((float) LEN * scale);
    const int RS = (int) ((float) LEN * (scale + 1) / 2);
    out->avg_len = avg;
    out->min_overlap = dp->min_overlap > 0 ? dp->min_overlap : L;
    out->rs_min_overlap = dp->rs_min_overlap > 0 ? dp->rs_min_overlap : (dp->min_overlap > 0 ? (dp->min_overlap + LEN) / 2 : RS);
    out->li_kmer_length = dp->min_overlap > 0 ? dp->min_overlap : (2 * L / 3 < 60 ? 2 * L / 3 : 60);  // Params.cpp:488-497 / main.cpp:103
    const uint32_t min_keep = (uint32_t) (3 + out->li_kmer_length);  // LI_KMER_INTERVALS (Params.cpp:706) + LI_KMER_LENGTH
    out->n_records[0] = info.n_records[0], out->n_records[1] = info.n_records[1];
    out->n_with_n = info.n_with_n, out->n_str = info.n_str;
    out->n_reads_in = fe.n;
    // main.cpp:132-140 -- duplicates / prefix reads
    if (dp->remove_type) CKR(fe_prefix_reads(fe, dp->remove_type, cfg));
    CK(cudaStreamSynchronize(0));
    const double t2 = now_ms();
    // main.cpp:150-232 (+ 253-266: reads too short for the graph creators become nullptr)
    CKR(fe_remap(fe, fe.raw(), dp->remove_type ? fe.mask.as<uint8_t>() : nullptr, fe.stride, min_keep, cfg));
    CK(cudaStreamSynchronize(0));
    const double t3 = now_ms();
    if (out->min_overlap < 1) {  // no read survived the reader: nothing to build
        CKR(fe_download(fe, true, &out->reads));
        out->graph.row_off = (uint64_t *) calloc(1, 8);
        return out->graph.row_off ? ALGA_OK : fail(ALGA_E_NOMEM, "host allocation failed");
    }
    // main.cpp:249-291 -- GraphCreatorPrefSuf + retainOnlySmallestOffset on the device-resident reads
    alga_ps_params pp{};
    pp.min_overlap = out->min_overlap, pp.rs_min_overlap = out->rs_min_overlap, pp.min_offset = 0, pp.max_len_cap = 500;
    pp.device = dp->input.device;
    if (g_build_plan && g_build_plan->params.device != pp.device) {
        alga_ps_plan_destroy(g_build_plan);
        g_build_plan = nullptr;
    }
    if (!g_build_plan) CKR(alga_ps_plan_create(&g_build_plan, &pp));
    g_build_plan->params = pp;
    CKR(fe.words2.ensure((size_t) (fe.n2 ? fe.n2 : 1) * fe.stride2 * 4));
    CKR(fe.len2.ensure((size_t) (fe.n2 ? fe.n2 : 1) * 4));
    alga_reads dr{};
    dr.n_reads = fe.n2, dr.words = fe.words2.as<uint32_t>(), dr.stride_words = fe.stride2, dr.len_nt = fe.len2.as<uint32_t>();
    // the renumbered read set goes back to the host on the copy stream while the graph is built from its device copy
    CK(cudaEventRecord(fe.ev_ready, 0));
    CK(cudaStreamWaitEvent(fe.copy, fe.ev_ready, 0));
    // (bulk through the DMA engine; the build's control read-backs go through a kernel meanwhile, or they would queue
    // behind it -- measured: 2.1 ms of a 12.4 ms call.  A/B switch ALGA_FE_COPY_KERNEL: bulk through a copy kernel instead.)
    static const bool copy_kernel = getenv("ALGA_FE_COPY_KERNEL") != nullptr;
    CKR(fe_download(fe, true, &out->reads, fe.copy, false, copy_kernel ? &cfg : nullptr));
    g_build_plan->peek_by_kernel = !copy_kernel;
    int rb = alga_ps_plan_bind_reads_device(g_build_plan, &dr, 0);
    if (rb == ALGA_OK) rb = alga_ps_plan_run(g_build_plan, nullptr);
    g_build_plan->peek_by_kernel = false;
    if (rb != ALGA_OK) {
        cudaStreamSynchronize(fe.copy);
        return rb;
    }
    const double t4 = now_ms();
    CKR(alga_ps_plan_result_host_pinned(g_build_plan, &out->graph));
    CK(cudaStreamSynchronize(fe.copy));
    const double t5 = now_ms();
    if (timing) {
        alga_ps_plan_stats(g_build_plan, timing);
        const double graph_dev_ms = timing->device_ms;
        timing->kernel_launches += launches;
        timing->h2d_ms = tm.h2d_ms;              // upload of the file text (copy stream; overlaps the scan of file 1)
        timing->d2h_ms = t5 - t4;                // what is left of the downloads once the graph is built (CSR + tail of the reads)
        timing->total_ms = t5 - t0;
        timing->device_ms = t4 - t0;             // call -> CSR resident on the device, uploads included
        timing->stage_ms[0] = tm.kernel_ms;      // reader: call -> packed reads resident (wall, uploads included)
        timing->stage_ms[1] = t2 - t1;           // prefix reads
        timing->stage_ms[2] = t3 - t2;           // renumbering
        timing->stage_ms[3] = graph_dev_ms;      // GraphCreatorPrefSuf pipeline (CUDA events)
        timing->stage_ms[4] = (t4 - t3) - graph_dev_ms;  // bind + host side of the graph build
        timing->stage_ms[5] = timing->stage_ms[6] = timing->stage_ms[7] = 0;
    }
    return ALGA_OK;
}

// ---- first simplifier step (simplify.cu) ------------------------------------------------------------------------------
namespace {
struct TriangleWs {  // process-wide cached workspace, as for the graph build and the input stage
    int device = -1;
    DevBuf row, nbr, off, keep, kept, new_row, onbr, ooff, big, nbig, tn, to, scan_ws;
    HostBuf h_row, h_nbr, h_off;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    void release() {
        for (DevBuf *b : {&row, &nbr, &off, &keep, &kept, &new_row, &onbr, &ooff, &big, &nbig, &tn, &to, &scan_ws}) b->release();
        for (HostBuf *b : {&h_row, &h_nbr, &h_off}) b->release();
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        e0 = e1 = nullptr;
    }
};
std::mutex g_tri_mutex;
TriangleWs g_tri;
}  // namespace

int alga_gpu_cut_triangles(const alga_csr *gin, int32_t max_offset, int32_t device, alga_csr *gout, alga_timing *timing) {
    if (!gin || !gout) return fail(ALGA_E_INVALID, "null argument");
    const uint32_t n = gin->n_reads;
    const uint64_t E = gin->n_edges;
    if (n && !gin->row_off) return fail(ALGA_E_INVALID, "row_off must not be null");
    if (E && (!gin->nbr || !gin->off)) return fail(ALGA_E_INVALID, "nbr / off must not be null");
    if (n && gin->row_off[n] != E) return fail(ALGA_E_INVALID, "row_off[n] differs from n_edges");
    std::lock_guard<std::mutex> lock(g_tri_mutex);
    const double t0 = now_ms();
    LaunchCfg cfg;
    uint64_t launches = 0;
    cfg.launches = &launches;
    CKR(pick_device(device, &cfg));
    TriangleWs &w = g_tri;
    if (w.device != device) w.release();
    w.device = device;
    if (!w.e0) CK(cudaEventCreate(&w.e0));
    if (!w.e1) CK(cudaEventCreate(&w.e1));
    // the input may be the borrowed result of an earlier call: copy it to the device before the staging buffers are reused
    CKR(w.row.ensure(((size_t) n + 1) * 8));
    CKR(w.nbr.ensure((size_t) (E ? E : 1) * 4));
    CKR(w.off.ensure((size_t) (E ? E : 1) * 4));
    if (n) CK(cudaMemcpyAsync(w.row.p, gin->row_off, ((size_t) n + 1) * 8, cudaMemcpyHostToDevice, 0));
    if (E) {
        CK(cudaMemcpyAsync(w.nbr.p, gin->nbr, (size_t) E * 4, cudaMemcpyHostToDevice, 0));
        CK(cudaMemcpyAsync(w.off.p, gin->off, (size_t) E * 4, cudaMemcpyHostToDevice, 0));
    }
    CK(cudaStreamSynchronize(0));
    const double t1 = now_ms();
    memset(gout, 0, sizeof(*gout));
    gout->n_reads = n;
    gout->borrowed = 1;
    CKR(w.h_row.ensure(((size_t) n + 1) * 8));
    gout->row_off = (uint64_t *) w.h_row.p;
    memset(gout->row_off, 0, ((size_t) n + 1) * 8);
    float dev_ms = 0;
    uint64_t E2 = 0;
    if (n && E) {
        CKR(w.keep.ensure((size_t) E));
        CKR(w.kept.ensure((size_t) n * 4));
        CKR(w.new_row.ensure(((size_t) n + 1) * 8));
        CKR(w.scan_ws.ensure(scan_workspace_bytes(n)));
        CKR(w.big.ensure((size_t) n * 4));
        CKR(w.nbig.ensure(4));
        CK(cudaEventRecord(w.e0, 0));
        // the marking kernel bisects row a for neighbour b: rows sorted by (neighbour, offset).  A graph built here arrives that
        // way, the OUTPUT of this call -- sorted by (offset, neighbour) -- or a reference graph after sortEdgesByIncreasingOffset
        // does not: sort the device copy first (the decision per entry does not depend on the order inside a row, and the
        // result is sorted by (offset, neighbour) below whatever came in).  No-op pass for sorted rows.
        CK(cudaMemsetAsync(w.nbig.p, 0, 4, 0));
        launch_sort_rows(w.row.as<uint64_t>(), n, w.nbr.as<int32_t>(), w.off.as<int32_t>(), w.big.as<uint32_t>(), w.nbig.as<uint32_t>(), 0, cfg);
        {
            uint32_t n_big_in = 0;
            CK(cudaMemcpy(&n_big_in, w.nbig.p, 4, cudaMemcpyDeviceToHost));
            if (n_big_in) {
                CKR(w.tn.ensure((size_t) E * 4));
                CKR(w.to.ensure((size_t) E * 4));
                launch_sort_big_rows(w.row.as<uint64_t>(), w.big.as<uint32_t>(), n_big_in, w.nbr.as<int32_t>(), w.off.as<int32_t>(),
                                     w.tn.as<int32_t>(), w.to.as<int32_t>(), 0, cfg);
                CK(cudaGetLastError());
            }
        }
        launch_triangle_marks(w.row.as<uint64_t>(), w.nbr.as<int32_t>(), w.off.as<int32_t>(), n, max_offset, w.keep.as<uint8_t>(),
                              w.kept.as<uint32_t>(), 0, cfg);
        launch_scan_u64(w.kept.as<uint32_t>(), w.new_row.as<uint64_t>(), n, w.scan_ws.p, 0, cfg);
        CK(cudaGetLastError());
        CK(cudaMemcpy(&E2, w.new_row.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost));
        CKR(w.onbr.ensure((size_t) (E2 ? E2 : 1) * 4));
        CKR(w.ooff.ensure((size_t) (E2 ? E2 : 1) * 4));
        CK(cudaMemsetAsync(w.nbig.p, 0, 4, 0));
        launch_triangle_compact(w.row.as<uint64_t>(), w.nbr.as<int32_t>(), w.off.as<int32_t>(), w.keep.as<uint8_t>(), n,
                                w.new_row.as<uint64_t>(), w.onbr.as<int32_t>(), w.ooff.as<int32_t>(), 0, cfg);
        // Graph::sortEdgesByIncreasingOffset: rows by (offset, neighbour) -- the (first, second) row sort with the arrays swapped
        launch_sort_rows(w.new_row.as<uint64_t>(), n, w.ooff.as<int32_t>(), w.onbr.as<int32_t>(), w.big.as<uint32_t>(),
                         w.nbig.as<uint32_t>(), 0, cfg);
        CK(cudaGetLastError());
        uint32_t n_big = 0;
        CK(cudaMemcpy(&n_big, w.nbig.p, 4, cudaMemcpyDeviceToHost));
        if (n_big) {
            CKR(w.tn.ensure((size_t) (E2 ? E2 : 1) * 4));
            CKR(w.to.ensure((size_t) (E2 ? E2 : 1) * 4));
            launch_sort_big_rows(w.new_row.as<uint64_t>(), w.big.as<uint32_t>(), n_big, w.ooff.as<int32_t>(), w.onbr.as<int32_t>(),
                                 w.tn.as<int32_t>(), w.to.as<int32_t>(), 0, cfg);
            CK(cudaGetLastError());
        }
        CK(cudaEventRecord(w.e1, 0));
        CK(cudaEventSynchronize(w.e1));
        CK(cudaEventElapsedTime(&dev_ms, w.e0, w.e1));
    }
    const double t2 = now_ms();
    CKR(w.h_nbr.ensure((size_t) (E2 ? E2 : 1) * 4));
    CKR(w.h_off.ensure((size_t) (E2 ? E2 : 1) * 4));
    gout->n_edges = E2;
    gout->nbr = (int32_t *) w.h_nbr.p;
    gout->off = (int32_t *) w.h_off.p;
    if (n && E) {
        CK(cudaMemcpyAsync(gout->row_off, w.new_row.p, ((size_t) n + 1) * 8, cudaMemcpyDeviceToHost, 0));
        if (E2) {
            CK(cudaMemcpyAsync(gout->nbr, w.onbr.p, (size_t) E2 * 4, cudaMemcpyDeviceToHost, 0));
            CK(cudaMemcpyAsync(gout->off, w.ooff.p, (size_t) E2 * 4, cudaMemcpyDeviceToHost, 0));
        }
        CK(cudaStreamSynchronize(0));
    }
    if (timing) {
        memset(timing, 0, sizeof(*timing));
        timing->h2d_ms = t1 - t0;
        timing->device_ms = dev_ms;
        timing->d2h_ms = now_ms() - t2;
        timing->total_ms = now_ms() - t0;
        timing->kernel_launches = launches;
    }
    return ALGA_OK;
}

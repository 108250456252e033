// alga_gpu_prefsuf_build_multi: the sharded overlap-graph build of alga_ps_shard_* driven from ONE process -- one host
// thread per GPU, peer access between the devices (NVLink), host barriers between the stages -- so that the reference's
// single-process driver (main.cpp:249-291 through shim/GraphCreatorPrefSufGpu.cpp, ALGA_GPU_DEVICES=<n>) can use all GPUs
// of the box.  Same stages, same kernels as the one-process-per-GPU harness (alga_b200/distributed.py):
//   H2D of the rank's shard | peers' shards over NVLink + this rank's slice of the seed tables | slices exchanged |
//   phase 1 | phase 2 | CSR rows | D2H of the rows.
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../../include/alga_gpu.h"
#include "launch.h"

using namespace alga;

namespace {

struct HostBarrier {
    std::mutex m;
    std::condition_variable cv;
    int n, waiting = 0, phase = 0;
    explicit HostBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m);
        const int ph = phase;
        if (++waiting == n) {
            waiting = 0;
            phase++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return phase != ph; });
        }
    }
};

struct Rank {
    alga_ps_plan *plan = nullptr;
    uint32_t *shard = nullptr;    // this rank's reads, caller's layout (peers copy it from here)
    uint32_t *compact = nullptr;  // landing zone of the peers' shards
    uint32_t *slots = nullptr;    // all reads, one sector-aligned slot each
    uint32_t *len = nullptr;
    void *ws = nullptr, *tp = nullptr, *ts = nullptr;
    cudaStream_t s = nullptr;
    int status = ALGA_OK;
    char err[512] = "";
    uint64_t n_edges = 0;
    double device_ms = 0;
};

void release(Rank &r) {
    if (r.plan) alga_ps_plan_destroy(r.plan);
    void *bufs[] = {r.shard, r.compact, r.slots, r.len, r.ws, r.tp, r.ts};
    for (void *p : bufs)
        if (p) cudaFree(p);
    if (r.s) cudaStreamDestroy(r.s);
    r = Rank{};
}

#define RCK(call)                                                                                         \
    do {                                                                                                  \
        cudaError_t e_ = (call);                                                                          \
        if (e_ != cudaSuccess && me.status == ALGA_OK) {                                                  \
            me.status = e_ == cudaErrorMemoryAllocation ? ALGA_E_NOMEM : ALGA_E_CUDA;                     \
            snprintf(me.err, sizeof(me.err), "GPU %d: %s failed: %s", d, #call, cudaGetErrorString(e_)); \
        }                                                                                                 \
    } while (0)
#define RST(expr)                                                                 \
    do {                                                                          \
        if (me.status == ALGA_OK) {                                               \
            const int r_ = (expr);                                                \
            if (r_ != ALGA_OK) {                                                  \
                me.status = r_;                                                   \
                snprintf(me.err, sizeof(me.err), "GPU %d: %s", d, alga_gpu_last_error()); \
            }                                                                     \
        }                                                                         \
    } while (0)

}  // namespace

namespace alga {
int set_error(int code, const char *msg);  // api.cu
}

extern "C" int alga_gpu_prefsuf_build_multi(const alga_reads *reads, const alga_ps_params *params, int32_t n_gpus, alga_csr *out,
                                            alga_timing *timing) {
    if (!reads || !params || !out) return set_error(ALGA_E_INVALID, "null argument");
    int n_dev = alga_gpu_device_count();
    if (n_dev == 0) return set_error(ALGA_E_CUDA, "no CUDA device available; libalga_gpu has no CPU fallback");
    if (n_gpus > 8) n_gpus = 8;
    if (n_gpus > n_dev) n_gpus = n_dev;
    // the sharded kernels take equal-length reads at a fixed stride, none removed, no flags cleared; anything else (and
    // small inputs) is built on one GPU
    const uint32_t n = reads->n_reads;
    uint32_t stride = reads->word_off ? 0u : reads->stride_words;
    bool uniform = n_gpus > 1 && n >= 4096 && reads->words && reads->len_nt && reads->len_nt[0] > 0;
    if (uniform && reads->word_off) {  // an offset array that is a fixed stride in disguise (the shim passes one)
        stride = (uint32_t) (reads->word_off[1] - reads->word_off[0]);
        for (uint32_t i = 0; i <= n && uniform; i++) uniform = reads->word_off[i] == (uint64_t) i * stride;
    }
    uniform = uniform && stride > 0;
    if (uniform) {
        const uint32_t l0 = reads->len_nt[0];
        for (uint32_t i = 0; i < n && uniform; i++)
            uniform = reads->len_nt[i] == l0 && (!reads->align_from || reads->align_from[i]) && (!reads->align_to || reads->align_to[i]);
        uniform = uniform && stride >= (l0 + 15u) / 16u;
    }
    if (!uniform) return alga_gpu_prefsuf_build(reads, params, out, timing);
    memset(out, 0, sizeof(*out));
    const auto t0 = std::chrono::steady_clock::now();
    const int world = n_gpus;
    const uint32_t len_nt = reads->len_nt[0], W = stride, S = aligned_stride_words((len_nt + 15u) / 16u);
    const uint32_t n_shard = (uint32_t) ((((uint64_t) n + world - 1) / world + 1) & ~1ull);
    std::vector<Rank> R(world);
    HostBarrier bar(world);
    std::vector<uint64_t> row_cnt(world, 0);
    out->row_off = (uint64_t *) malloc(((size_t) n + 1) * 8);
    if (!out->row_off) return set_error(ALGA_E_NOMEM, "out of host memory");
    std::vector<int32_t *> h_nbr(world, nullptr), h_off(world, nullptr);
    alga_ps_shard base{};
    base.world = world, base.n_shard = n_shard, base.n_total = n;

    auto worker = [&](int d) {
        Rank &me = R[d];
        const uint32_t lo = (uint32_t) std::min<uint64_t>((uint64_t) d * n_shard, n), hi = (uint32_t) std::min<uint64_t>((uint64_t) (d + 1) * n_shard, n);
        RCK(cudaSetDevice(d));
        for (int p = 0; p < world; p++)
            if (p != d) {
                cudaError_t e = cudaDeviceEnablePeerAccess(p, 0);
                if (e == cudaErrorPeerAccessAlreadyEnabled) cudaGetLastError();
                else if (e != cudaSuccess && me.status == ALGA_OK) {
                    me.status = ALGA_E_CUDA;
                    snprintf(me.err, sizeof(me.err), "GPU %d cannot reach GPU %d over peer memory: %s", d, p, cudaGetErrorString(e));
                    cudaGetLastError();
                }
            }
        alga_ps_params pp = *params;
        pp.device = d;
        RST(alga_ps_plan_create(&me.plan, &pp));
        RCK(cudaStreamCreateWithFlags(&me.s, cudaStreamNonBlocking));
        const size_t full_words = (size_t) n_shard * world * S;
        RCK(cudaMalloc(&me.shard, (size_t) n_shard * W * 4 + 256));
        RCK(cudaMalloc(&me.compact, (size_t) n_shard * W * 4 + 256));
        RCK(cudaMalloc(&me.slots, full_words * 4 + kReadPadBytes));
        RCK(cudaMalloc(&me.len, (size_t) n * 4));
        const uint64_t ws_bytes = alga_ps_shard_ws_bytes(n_shard, world), tb = alga_ps_shard_table_bytes(n, world);
        RCK(cudaMalloc(&me.ws, ws_bytes));
        RCK(cudaMalloc(&me.tp, tb));
        RCK(cudaMalloc(&me.ts, tb));
        if (me.status == ALGA_OK) {
            RCK(cudaMemsetAsync(me.ws, 0, ws_bytes, me.s));
            RCK(cudaMemsetAsync(me.slots, 0, full_words * 4 + kReadPadBytes, me.s));
            RCK(cudaMemcpyAsync(me.len, reads->len_nt, (size_t) n * 4, cudaMemcpyHostToDevice, me.s));
            RCK(cudaMemcpyAsync(me.shard, reads->words + (size_t) lo * W, (size_t) (hi - lo) * W * 4, cudaMemcpyHostToDevice, me.s));
            RCK(cudaStreamSynchronize(me.s));
        }
        bar.wait();  // every rank's shard and buffers are in place
        bool all_ok = true;
        for (int p = 0; p < world; p++) all_ok = all_ok && R[p].status == ALGA_OK;
        alga_ps_shard sh = base;
        sh.rank = d;
        for (int p = 0; p < world; p++) sh.peer_ws[p] = R[p].ws;
        sh.table_prefix = me.tp, sh.table_suffix = me.ts;
        cudaEvent_t e0 = nullptr, e1 = nullptr;
        if (all_ok) {
            alga_reads dr{};
            dr.n_reads = n, dr.words = me.slots, dr.stride_words = S, dr.len_nt = me.len;
            RST(alga_ps_plan_bind_reads_uniform(me.plan, &dr, len_nt));
            RCK(cudaEventCreate(&e0));
            RCK(cudaEventCreate(&e1));
            RCK(cudaEventRecord(e0, me.s));
            LaunchCfg cfg;
            cudaDeviceGetAttribute(&cfg.sm_count, cudaDevAttrMultiProcessorCount, d);
            for (int k = 0; k < world && me.status == ALGA_OK; k++) {  // the peers' shards over NVLink, this rank's slice of the tables
                const int p = (d + k) % world;
                const uint32_t plo = (uint32_t) std::min<uint64_t>((uint64_t) p * n_shard, n), phi = (uint32_t) std::min<uint64_t>((uint64_t) (p + 1) * n_shard, n);
                const uint32_t *src = R[p].shard;
                if (p != d) {
                    RCK(cudaMemcpyPeerAsync(me.compact, d, R[p].shard, p, (size_t) (phi - plo) * W * 4, me.s));
                    src = me.compact;
                }
                launch_repack_reads(src, W, std::min(W, S), phi - plo, me.slots + (size_t) plo * S, S, me.s, cfg);
                RST(alga_ps_shard_index_range(me.plan, &sh, plo, phi, k == 0, me.s));
            }
            RCK(cudaStreamSynchronize(me.s));
        }
        bar.wait();  // every rank's slice of the tables is final
        all_ok = true;
        for (int p = 0; p < world; p++) all_ok = all_ok && R[p].status == ALGA_OK;
        if (all_ok) {
            const size_t sb = (size_t) (alga_ps_shard_table_bytes(n, world) / world);
            for (int k = 1; k < world; k++) {
                const int p = (d + k) % world;
                RCK(cudaMemcpyPeerAsync((char *) me.tp + p * sb, d, (char *) R[p].tp + p * sb, p, sb, me.s));
                RCK(cudaMemcpyPeerAsync((char *) me.ts + p * sb, d, (char *) R[p].ts + p * sb, p, sb, me.s));
            }
            RST(alga_ps_shard_phase1(me.plan, &sh, me.s));
            RCK(cudaStreamSynchronize(me.s));
        }
        bar.wait();  // every rank's phase-1 edges are in its workspace
        all_ok = true;
        for (int p = 0; p < world; p++) all_ok = all_ok && R[p].status == ALGA_OK;
        if (all_ok) RST(alga_ps_shard_phase2(me.plan, &sh, me.s));  // synchronises its stream
        bar.wait();  // ... and the survivors
        all_ok = true;
        for (int p = 0; p < world; p++) all_ok = all_ok && R[p].status == ALGA_OK;
        if (all_ok) {
            RST(alga_ps_shard_csr(me.plan, &sh, me.s));
            if (e1) {
                RCK(cudaEventRecord(e1, me.s));
                RCK(cudaEventSynchronize(e1));
                float ms = 0;
                if (cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess) me.device_ms = ms;
            }
            const uint64_t *d_ro = nullptr;
            const int32_t *d_nb = nullptr, *d_of = nullptr;
            uint64_t ne = 0;
            RST(alga_ps_plan_result_device(me.plan, &d_ro, &d_nb, &d_of, &ne));
            if (me.status == ALGA_OK) {
                me.n_edges = ne;
                row_cnt[d] = ne;
                h_nbr[d] = (int32_t *) malloc((size_t) (ne ? ne : 1) * 4);
                h_off[d] = (int32_t *) malloc((size_t) (ne ? ne : 1) * 4);
                if (!h_nbr[d] || !h_off[d]) {
                    me.status = ALGA_E_NOMEM;
                    snprintf(me.err, sizeof(me.err), "out of host memory");
                } else {
                    // rows of [lo, hi): offsets relative to the rank's first edge, shifted by the caller below
                    RCK(cudaMemcpy(out->row_off + lo, d_ro, (size_t) (hi - lo) * 8, cudaMemcpyDeviceToHost));
                    if (ne) {
                        RCK(cudaMemcpy(h_nbr[d], d_nb, (size_t) ne * 4, cudaMemcpyDeviceToHost));
                        RCK(cudaMemcpy(h_off[d], d_of, (size_t) ne * 4, cudaMemcpyDeviceToHost));
                    }
                }
            }
        }
        if (e0) cudaEventDestroy(e0);
        if (e1) cudaEventDestroy(e1);
        bar.wait();  // nobody reads a peer's buffers any more
        cudaSetDevice(d);
        const int st = me.status;
        char err[512];
        memcpy(err, me.err, sizeof(err));
        const uint64_t ne = me.n_edges;
        const double ms = me.device_ms;
        release(me);
        me.status = st, me.n_edges = ne, me.device_ms = ms;
        memcpy(me.err, err, sizeof(err));
    };

    std::vector<std::thread> th;
    for (int d = 0; d < world; d++) th.emplace_back(worker, d);
    for (auto &t : th) t.join();
    int status = ALGA_OK;
    const char *msg = "";
    for (int d = 0; d < world; d++)
        if (R[d].status != ALGA_OK && status == ALGA_OK) status = R[d].status, msg = R[d].err;
    uint64_t total = 0;
    double dev_ms = 0;
    for (int d = 0; d < world; d++) total += row_cnt[d], dev_ms = std::max(dev_ms, R[d].device_ms);
    if (status == ALGA_OK) {
        out->nbr = (int32_t *) malloc((size_t) (total ? total : 1) * 4);
        out->off = (int32_t *) malloc((size_t) (total ? total : 1) * 4);
        if (!out->nbr || !out->off) status = ALGA_E_NOMEM, msg = "out of host memory";
    }
    if (status == ALGA_OK) {
        uint64_t base_e = 0;
        for (int d = 0; d < world; d++) {
            const uint32_t lo = (uint32_t) std::min<uint64_t>((uint64_t) d * n_shard, n), hi = (uint32_t) std::min<uint64_t>((uint64_t) (d + 1) * n_shard, n);
            for (uint32_t i = lo; i < hi; i++) out->row_off[i] += base_e;
            if (row_cnt[d]) {
                memcpy(out->nbr + base_e, h_nbr[d], (size_t) row_cnt[d] * 4);
                memcpy(out->off + base_e, h_off[d], (size_t) row_cnt[d] * 4);
            }
            base_e += row_cnt[d];
        }
        out->row_off[n] = total;
        out->n_reads = n, out->n_edges = total, out->borrowed = 0;
    }
    for (int d = 0; d < world; d++) {
        free(h_nbr[d]);
        free(h_off[d]);
    }
    if (status != ALGA_OK) {
        alga_gpu_free_csr(out);
        return set_error(status, msg);
    }
    if (timing) {
        memset(timing, 0, sizeof(*timing));
        timing->device_ms = dev_ms;
        timing->total_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        timing->stage_ms[7] = world;
    }
    return ALGA_OK;
}

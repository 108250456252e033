// Drop-in replacement for the reference's src/GraphCreators/GraphCreatorPrefSuf.cpp.
//
// Build ALGA with this file INSTEAD of that one and link libalga_gpu.so: the declared class
// (include/GraphCreators/GraphCreatorPrefSuf.h:20-28 -- constructor, virtual destructor, clear(),
// startAlignmentGraphCreation()) keeps its name and signature, so main.cpp:249-291 and the second call site
// main.cpp:649-655 run unchanged; every other source file of the reference is compiled as it is.
// This file contains no algorithm: it gathers vector<Read*> into the packed layout of include/alga_gpu.h, calls
// alga_gpu_prefsuf_build (hand-written CUDA, sm_100a) and copies the returned CSR into Graph::V.
//
// Contract reproduced from the reference (file:line in /root/reference):
//   * reads and G are borrowed; (*reads)[i] may be nullptr; read->getId() == i      (GraphCreator.cpp:9-17, main.cpp:225-231)
//   * parameters come from Params statics                                             (GraphCreatorPrefSuf.cpp:76-95)
//   * net side effect on Params: MIN_OVERLAP_AREA = MIN_OVERLAP_PREF_SUF             (GraphCreatorPrefSuf.cpp:76,120)
//   * TimeMeasurer::GRAPH_CREATOR is started / stopped around the call                (GraphCreatorPrefSuf.cpp:74,125)
//   * Global::removeIsolatedReads() when constructed with remove_isolated_reads      (GraphCreatorPrefSuf.cpp:105)
//   * on return (*G)[b] holds the forward edges (c, offset); rows come back sorted by c with one entry per c,
//     which is what Graph::retainOnlySmallestOffset (main.cpp:291) would leave anyway
//   * errors: message on cerr + exit(1), no exception crosses the boundary            (e.g. InputReader.cpp:324-327)
#include <GraphCreators/GraphCreatorPrefSuf.h>

#include <Global.h>
#include <Params.h>
#include <Utils/TimeMeasurer.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "alga_gpu.h"

namespace {
struct PinnedBuffer {  // page-locked staging memory from the library: the upload runs at full host->device rate
    void *p = nullptr;
    explicit PinnedBuffer(size_t bytes) : p(alga_gpu_host_alloc(bytes)) {
        if (!p) {
            std::cerr << "alga_gpu: " << alga_gpu_last_error() << std::endl;
            exit(1);
        }
    }
    ~PinnedBuffer() { alga_gpu_host_free(p); }
};
}  // namespace

GraphCreatorPrefSuf::GraphCreatorPrefSuf(vector<Read *> *reads, Graph *G, bool remove_isolated_reads)
    : GraphCreator(reads, G), removeIsolatedReadsBeforeReversingGraph(remove_isolated_reads) {
    maxReadLength = 0;
    currentPrefSufLength = 0;
    prefHashFactor = 0;
    prefHashFactorAdditional = 0;
    prefixKmersBuckets = 0;
}

GraphCreatorPrefSuf::~GraphCreatorPrefSuf() { clear(); }

void GraphCreatorPrefSuf::clear() {}  // the device workspace is cached inside libalga_gpu across calls

void GraphCreatorPrefSuf::startAlignmentGraphCreation() {
    TimeMeasurer::startMeasurement(TimeMeasurer::GRAPH_CREATOR);
    Params::MIN_OVERLAP_AREA = Params::MIN_OVERLAP_PREF_SUF;

    const uint32_t n = (uint32_t) reads->size();
    std::vector<uint64_t> word_off((size_t) n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        Read *r = (*reads)[i];
        word_off[i + 1] = word_off[i] + (r ? (uint64_t) r->getSequence().countBlocks() : 0);
    }
    const uint64_t n_words = word_off[n];
    PinnedBuffer words_buf((size_t) (n_words ? n_words : 1) * sizeof(uint32_t));
    uint32_t *words = static_cast<uint32_t *>(words_buf.p);
    std::vector<uint32_t> len(n, 0);
    std::vector<uint8_t> from(n, 0), to(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        Read *r = (*reads)[i];
        if (!r) continue;
        Bitset &seq = r->getSequence();
        const int nb = (int) (word_off[i + 1] - word_off[i]);
        for (int b = 0; b < nb; b++) words[word_off[i] + b] = seq.getBlock(b);
        len[i] = (uint32_t) r->size();
        from[i] = alignFrom[i] ? 1 : 0;
        to[i] = alignTo[i] ? 1 : 0;
    }

    alga_reads in;
    memset(&in, 0, sizeof(in));
    in.n_reads = n;
    in.words = words;
    in.word_off = word_off.data();
    in.len_nt = len.data();
    in.align_from = from.data();
    in.align_to = to.data();
    alga_ps_params p;
    memset(&p, 0, sizeof(p));
    p.min_overlap = Params::MIN_OVERLAP_PREF_SUF;
    p.rs_min_overlap = Params::REMOVE_SMALL_OVERLAP_EDGES_MIN_OVERLAP;
    p.min_offset = Params::MIN_OFFSET_FOR_ALIGNMENT;
    p.max_len_cap = 500;  // GraphCreatorPrefSuf.cpp:92
    const char *dev = getenv("ALGA_GPU_DEVICE");
    p.device = dev ? atoi(dev) : 0;

    // ALGA_GPU_DEVICES=<n>: use up to n GPUs of the box (equal-length read sets; anything else runs on one GPU)
    const char *ndev = getenv("ALGA_GPU_DEVICES");
    const int n_gpus = ndev ? atoi(ndev) : 1;

    alga_csr g;
    alga_timing t;
    const int rc = n_gpus > 1 ? alga_gpu_prefsuf_build_multi(&in, &p, n_gpus, &g, &t) : alga_gpu_prefsuf_build(&in, &p, &g, &t);
    if (rc != ALGA_OK) {
        std::cerr << "alga_gpu_prefsuf_build failed: " << alga_gpu_last_error() << std::endl;
        exit(1);
    }
    for (uint32_t b = 0; b < n; b++) {
        VPII &row = (*G)[(int) b];
        row.clear();
        const uint64_t s = g.row_off[b], e = g.row_off[b + 1];
        row.reserve((size_t) (e - s));
        for (uint64_t k = s; k < e; k++) row.emplace_back(g.nbr[k], g.off[k]);
    }
    std::cerr << "alga_gpu: " << n << " reads, " << g.n_edges << " edges, device " << t.device_ms << " ms, call "
              << t.total_ms << " ms" << std::endl;
    alga_gpu_free_csr(&g);

    if (removeIsolatedReadsBeforeReversingGraph) Global::removeIsolatedReads();
    TimeMeasurer::stopMeasurement(TimeMeasurer::GRAPH_CREATOR);
}

// Drop-in replacement for the reference's src/GraphCreators/GraphCreatorLI.cpp (the error-rate supplement,
// main.cpp:300-355).
//
// Build ALGA with this file INSTEAD of that one and link libalga_gpu.so: the declared class
// (include/GraphCreators/GraphCreatorLI.h:14-38 -- constructor, createAlignmentsForKmers, startAlignmentGraphCreation;
// setAlignFrom / setAlignTo are inline in the header and forward to `graphCreator`) keeps its name and signature, so
// main.cpp:306-350 runs unchanged.  This file contains no algorithm: it gathers vector<Read*> and Graph::V into the
// layouts of include/alga_gpu.h, calls alga_gpu_supplement (LI k-mers, their sort and canAlign on the GPU, ordered
// edge replay on the host) and copies the returned rows into Graph::V.
//
// Contract reproduced from the reference (file:line in /root/reference):
//   * reads and G are borrowed; flags were set by main.cpp:308-323 to "dead ends only" -- alga_gpu_supplement derives
//     the same flags from the graph it is given, the forwarded flag vectors are not read
//   * parameters come from Params statics set in main.cpp:332-340
//   * four passes with Read::priorities rotated after each (GraphCreatorLI.cpp:18-28): the net rotation is the
//     identity, Read::priorities is left as it was
//   * on return (*G)[b] holds rows sorted by target with one entry per target, which is what
//     Graph::retainOnlySmallestOffset (main.cpp:346) would leave anyway
//   * errors: message on cerr + exit(1)
#include <GraphCreators/GraphCreatorLI.h>
#include <GraphCreators/GraphCreatorPairwiseKmerBranch.h>

#include <Global.h>
#include <Params.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "alga_gpu.h"

GraphCreatorLI::GraphCreatorLI(vector<Read *> *reads, Graph *G) : GraphCreatorKmerBased(reads, G) {
    graphCreator = new GraphCreatorPairwiseKmerBranch(reads, G);  // target of the header's inline flag forwarding
}

void GraphCreatorLI::createAlignmentsForKmers(vector<Kmer> &kmers, int p, int q, int thread_id) {
    graphCreator->createAlignmentsForKmers(kmers, p, q);
}

void GraphCreatorLI::startAlignmentGraphCreation() {
    const uint32_t n = (uint32_t) reads->size();
    std::vector<uint64_t> word_off((size_t) n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        Read *r = (*reads)[i];
        word_off[i + 1] = word_off[i] + (r ? (uint64_t) r->getSequence().countBlocks() : 0);
    }
    std::vector<uint32_t> words((size_t) (word_off[n] ? word_off[n] : 1)), len(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        Read *r = (*reads)[i];
        if (!r) continue;
        Bitset &seq = r->getSequence();
        const int nb = (int) (word_off[i + 1] - word_off[i]);
        for (int b = 0; b < nb; b++) words[word_off[i] + b] = seq.getBlock(b);
        len[i] = (uint32_t) r->size();
    }
    std::vector<uint64_t> row_off((size_t) n + 1, 0);
    for (uint32_t i = 0; i < n; i++) row_off[i + 1] = row_off[i] + (*G)[(int) i].size();
    std::vector<int32_t> nbr((size_t) (row_off[n] ? row_off[n] : 1)), off((size_t) (row_off[n] ? row_off[n] : 1));
    for (uint32_t i = 0; i < n; i++) {
        uint64_t k = row_off[i];
        for (auto &e : (*G)[(int) i]) {
            nbr[k] = e.first;
            off[k] = e.second;
            k++;
        }
    }

    alga_reads in;
    memset(&in, 0, sizeof(in));
    in.n_reads = n;
    in.words = words.data();
    in.word_off = word_off.data();
    in.len_nt = len.data();
    alga_csr gin;
    memset(&gin, 0, sizeof(gin));
    gin.n_reads = n;
    gin.n_edges = row_off[n];
    gin.row_off = row_off.data();
    gin.nbr = nbr.data();
    gin.off = off.data();
    gin.borrowed = 1;
    alga_sup_params p;
    memset(&p, 0, sizeof(p));
    p.max_offset_pct = Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT;
    p.min_offset = Params::MIN_OFFSET_FOR_ALIGNMENT;
    p.min_overlap_area = Params::MIN_OVERLAP_AREA;
    p.threshold_pct = Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR;
    p.same_ends = Params::ALIGNMENT_CONTROLLER_SAME_ENDS_LENGTH;
    p.kmer_length = Params::LI_KMER_LENGTH;
    p.intervals = Params::LI_KMER_INTERVALS;
    p.kmer_length_bucket = Params::KMER_LENGTH_BUCKET;
    const char *dev = getenv("ALGA_GPU_DEVICE");
    p.device = dev ? atoi(dev) : 0;

    alga_csr g;
    alga_timing t;
    if (alga_gpu_supplement(&in, &gin, &p, &g, &t) != ALGA_OK) {
        std::cerr << "alga_gpu_supplement failed: " << alga_gpu_last_error() << std::endl;
        exit(1);
    }
    for (uint32_t b = 0; b < n; b++) {
        VPII &row = (*G)[(int) b];
        row.clear();
        const uint64_t s = g.row_off[b], e = g.row_off[b + 1];
        row.reserve((size_t) (e - s));
        for (uint64_t k = s; k < e; k++) row.emplace_back(g.nbr[k], g.off[k]);
    }
    std::cerr << "alga_gpu supplement: " << (uint64_t) t.stage_ms[5] << " dead-end reads, " << (uint64_t) t.stage_ms[6]
              << " pairs verified, " << gin.n_edges << " -> " << g.n_edges << " edges, call " << t.total_ms << " ms" << std::endl;
    alga_gpu_free_csr(&g);
}

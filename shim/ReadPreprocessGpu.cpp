// Drop-in replacement for the reference's src/IO/ReadPreprocess.cpp.
//
// Build ALGA with this file INSTEAD of that one and link libalga_gpu.so: ReadPreprocess::getPrefixReads
// (include/IO/ReadPreprocess.h:11-17, called from main.cpp:132-134) keeps its signature; the helpers getSortedReads /
// getLCP are used by nothing else in the reference and are not provided.  No algorithm here: the reads are gathered into
// the packed layout of include/alga_gpu.h and alga_gpu_prefix_reads (CUDA) returns the removal mask.
//
// Contract reproduced (file:line in /root/reference): Params::REMOVE_PREF_READS_TYPE selects duplicates only (1) or all
// prefix reads (2) (ReadPreprocess.cpp:36-50); the result has Global::READS.size() entries, true = the driver calls
// Global::removeRead on it (main.cpp:135-140); nullptr reads are never marked.
#include <IO/ReadPreprocess.h>

#include <Params.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <vector>

#include "alga_gpu.h"

VB ReadPreprocess::getPrefixReads() {
    vector<Read *> &reads = Global::READS;
    const uint32_t n = (uint32_t) reads.size();
    std::vector<uint64_t> word_off((size_t) n + 1, 0);
    for (uint32_t i = 0; i < n; i++)
        word_off[i + 1] = word_off[i] + (reads[i] ? (uint64_t) reads[i]->getSequence().countBlocks() : 0);
    uint32_t *words = static_cast<uint32_t *>(alga_gpu_host_alloc((size_t) (word_off[n] ? word_off[n] : 1) * sizeof(uint32_t)));
    if (!words) {
        std::cerr << "alga_gpu: " << alga_gpu_last_error() << std::endl;
        exit(1);
    }
    std::vector<uint32_t> len(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        Read *r = reads[i];
        if (!r) continue;
        Bitset &seq = r->getSequence();
        const int nb = (int) (word_off[i + 1] - word_off[i]);
        for (int b = 0; b < nb; b++) words[word_off[i] + b] = seq.getBlock(b);
        len[i] = (uint32_t) r->size();
    }
    alga_reads in;
    memset(&in, 0, sizeof(in));
    in.n_reads = n;
    in.words = words;
    in.word_off = word_off.data();
    in.len_nt = len.data();
    const int type = Params::REMOVE_PREF_READS_TYPE == Params::PREF_READS_ONLY_DUPLICATES ? 1 : 2;
    const char *dev = getenv("ALGA_GPU_DEVICE");
    std::vector<uint8_t> mask(n ? n : 1, 0);
    alga_timing t;
    if (alga_gpu_prefix_reads(&in, type, dev ? atoi(dev) : 0, mask.data(), &t) != ALGA_OK) {
        std::cerr << "alga_gpu_prefix_reads failed: " << alga_gpu_last_error() << std::endl;
        exit(1);
    }
    alga_gpu_host_free(words);
    VB res(n, false);
    uint64_t cnt = 0;
    for (uint32_t i = 0; i < n; i++)
        if (mask[i]) res[i] = true, cnt++;
    std::cerr << "alga_gpu preprocess: " << n << " reads, " << cnt << " marked, device " << t.device_ms << " ms, call " << t.total_ms
              << " ms" << std::endl;
    return res;
}

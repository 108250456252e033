// Drop-in replacement for the reference's src/IO/InputReader.cpp.
//
// Build ALGA with this file INSTEAD of that one and link libalga_gpu.so: InputReader keeps its declared interface
// (include/IO/InputReader.h:25-49; main.cpp:81-82 constructs it and calls readInput()), every other source file of the
// reference is compiled as it is.  This file contains no parsing: it reads the input file(s) into page-locked memory,
// calls alga_gpu_read_input (record splitting, trimming, N / short-period filters, reverse complements, 2-bit packing
// and the final order of Global::READS, all in CUDA) and wraps the packed reads into the reference's Read objects.
//
// Contract reproduced from the reference (file:line in /root/reference):
//   * files: Params::inStreamFilePath1 / inStreamFilePath2; the second one only with Params::ADD_PAIRED_READS and a file
//     type other than PFASTA (InputReader.cpp:54-56, 186-187); parser by Params::INPUT_FILE_TYPE (Params.cpp:332-335)
//   * Params::READ_END_TRIM_LEFT / RIGHT, Params::RNA, the minimal-period threshold 20 (InputReader.cpp:298-303, 343)
//   * on return Global::READS[i] is nullptr for removed reads, otherwise a Read with getId() == i
//     (InputReader.cpp:78-85); the order is the --threads=1 order of the reference
//   * the summary lines on cerr (InputReader.cpp:117-122)
//   * errors: message on cerr + exit(1)                                               (InputReader.cpp:324-327)
// Not reproduced: Params::REMOVE_READS_WITH_N == 0 (N replaced by a pseudo-random nucleotide, thread-count dependent in
// the reference) and Params::ADD_COMP_REV_READS == 0 -- both non-default; the shim refuses them.
#include <IO/InputReader.h>

#include <Global.h>
#include <Params.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "alga_gpu.h"

namespace {

struct PinnedFile {  // the whole file in page-locked memory: the upload runs at the full host->device rate
    uint8_t *p = nullptr;
    uint64_t n = 0;
    bool load(const std::string &path) {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) return false;
        fseek(f, 0, SEEK_END);
        const long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        n = sz > 0 ? (uint64_t) sz : 0;
        p = static_cast<uint8_t *>(alga_gpu_host_alloc(n ? n : 1));
        if (!p) {
            std::cerr << "alga_gpu: " << alga_gpu_last_error() << std::endl;
            exit(1);
        }
        const size_t got = n ? fread(p, 1, n, f) : 0;
        fclose(f);
        return got == n;
    }
    ~PinnedFile() {
        if (p) alga_gpu_host_free(p);
    }
};

}  // namespace

InputReader::InputReader() {
    STRreads = VI(Params::THREADS, 0);
    Nreads = VI(Params::THREADS, 0);
    NsInRead = VVI(Params::THREADS, VI(1000, 0));
}

void InputReader::readInput() {
    if (!Params::REMOVE_READS_WITH_N || !Params::ADD_COMP_REV_READS) {
        std::cerr << "alga_gpu reader: --remove_reads_with_n=0 / reads without reverse complements are not supported" << std::endl;
        exit(1);
    }
    Global::READS.clear();
    cerr << "starting to read" << endl;
    PinnedFile f1, f2;
    if (!f1.load(Params::inStreamFilePath1)) {
        std::cerr << "alga_gpu reader: cannot read " << Params::inStreamFilePath1 << std::endl;
        exit(1);
    }
    const bool paired = Params::ADD_PAIRED_READS && Params::INPUT_FILE_TYPE != Params::PFASTA && Params::inStreamFilePath2 != "";
    if (paired && !f2.load(Params::inStreamFilePath2)) {
        std::cerr << "alga_gpu reader: cannot read " << Params::inStreamFilePath2 << std::endl;
        exit(1);
    }
    if (Params::INPUT_FILE_TYPE == Params::PFASTA && !Params::ADD_PAIRED_READS) {
        std::cerr << "alga_gpu reader: .pfasta without paired reads is not supported" << std::endl;
        exit(1);
    }

    alga_input_params p;
    memset(&p, 0, sizeof(p));
    p.file_type = Params::INPUT_FILE_TYPE == Params::FASTQ ? ALGA_INPUT_FASTQ
                  : (Params::INPUT_FILE_TYPE == Params::FASTA || Params::INPUT_FILE_TYPE == Params::PFASTA) ? ALGA_INPUT_FASTA
                                                                                                                : ALGA_INPUT_PLAIN;
    p.trim_left = Params::READ_END_TRIM_LEFT;
    p.trim_right = Params::READ_END_TRIM_RIGHT;
    p.rna = Params::RNA;
    p.str_threshold = 20;  // InputReader.cpp:343
    const char *dev = getenv("ALGA_GPU_DEVICE");
    p.device = dev ? atoi(dev) : 0;

    alga_read_set rs;
    alga_timing t;
    if (alga_gpu_read_input(f1.p, f1.n, paired ? f2.p : nullptr, paired ? f2.n : 0, &p, &rs, &t) != ALGA_OK) {
        std::cerr << "alga_gpu_read_input failed: " << alga_gpu_last_error() << std::endl;
        exit(1);
    }

    // packed blocks -> Read objects (the rest of the reference works on vector<Read*>); built through the reference's
    // own constructor so that every Bitset field is exactly what Read::createSequence leaves
    const uint32_t n = rs.n_reads;
    Global::READS.assign(n, nullptr);
    const int T = Params::THREADS > 0 ? Params::THREADS : 1;
    auto job = [&](uint32_t lo, uint32_t hi) {
        static const char nt[4] = {'A', 'C', 'G', 'T'};
        std::string s;
        for (uint32_t i = lo; i < hi; i++) {
            const uint32_t len = rs.len_nt[i];
            if (!len) continue;
            const uint32_t *w = rs.words + (uint64_t) i * rs.stride_words;
            s.resize(len);
            for (uint32_t j = 0; j < len; j++) s[j] = nt[(w[j >> 4] >> ((j & 15) * 2)) & 3];
            Global::READS[i] = new Read((int) i, s);
        }
    };
    std::vector<std::thread> workers;
    const uint32_t chunk = (n + T - 1) / T;
    for (int k = 1; k < T; k++) {
        const uint32_t lo = (uint32_t) k * chunk < n ? (uint32_t) k * chunk : n, hi = lo + chunk < n ? lo + chunk : n;
        workers.emplace_back(job, lo, hi);
    }
    job(0, chunk < n ? chunk : n);
    for (auto &th : workers) th.join();
    Global::READS.shrink_to_fit();

    cerr << "alga_gpu reader: " << rs.n_records[0] << (paired ? " + " + to_string(rs.n_records[1]) : string()) << " records, " << n
         << " reads, upload " << t.h2d_ms << " ms, call " << t.total_ms << " ms" << endl;
    cerr << "There were " << 2 * rs.n_with_n << " reads that contained N and were removed from graph creation process" << endl;
    cerr << "There were  " << 2 * rs.n_str << " reads marked as STR and were removed from graph creation process" << endl;
    alga_gpu_free_read_set(&rs);
}

// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Harness around the UNMODIFIED reference sources (compiled where they lie under
// /root/reference by oracle/Makefile; outputs only into oracle/_ref/).  It feeds
// an already-preprocessed packed read set (the exact input of the hot path)
// into the reference's own GraphCreatorPrefSuf and dumps the resulting edge set.
//
// It reproduces the call sequence of the reference driver around the hot path:
//   main.cpp:239      Global::GRAPH = Graph(N)
//   main.cpp:249      new GraphCreatorPrefSuf(READS, G, false)
//   main.cpp:253-280  setAlignFrom / setAlignTo
//   main.cpp:282      startAlignmentGraphCreation()
//   main.cpp:286-291  clear(); delete; G->retainOnlySmallestOffset()
// and, with --verify, AlignmentControllerLowErrorRate::canAlign on a pair list
// (AlignmentControllerHybrid.cpp:46-83 -> AlignmentControllerLowErrorRate.cpp:15-49), and, with
// `supplement`, the error-rate supplement GraphCreatorLI (main.cpp:300-355) on a given graph.
//
// File formats are documented in oracle/README.md (ALGR = reads, ALGE = edges).
#include <Global.h>
#include <Params.h>
#include <DataStructures/Bitset.h>
#include <DataStructures/Read.h>
#include <DataStructures/Graph.h>
#include <GraphCreators/GraphCreatorPrefSuf.h>
#include <GraphCreators/GraphCreatorLI.h>
#include <IO/ReadPreprocess.h>
#include <IO/InputReader.h>
#include <GraphSimplifiers/GraphSimplifier.h>
#include <AlignmentControllers/AlignmentControllerHybrid.h>
#include <AlignmentControllers/AlignmentControllerLowErrorRate.h>

#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

int alga_reference_main(int argc, char **argv);  // the reference's main() (oracle/Makefile: -Dmain=alga_reference_main)

namespace {

struct ReadsFile {
    uint32_t n = 0;
    int32_t min_overlap = 0, rs = 0, min_offset = 0, reserved = 0;
    std::vector<uint32_t> len;
    std::vector<uint8_t> align_from, align_to;
    std::vector<uint64_t> word_off;
    std::vector<uint32_t> words;
};

void die(const char *msg) {
    fprintf(stderr, "ref_harness: %s\n", msg);
    exit(2);
}

template <class T>
void read_vec(FILE *f, std::vector<T> &v, size_t n) {
    v.resize(n);
    if (n && fread(v.data(), sizeof(T), n, f) != n) die("short read");
}

ReadsFile load_reads(const char *path) {
    FILE *f = fopen(path, "rb");
    if (!f) die("cannot open reads file");
    char magic[4];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ALGR", 4) != 0) die("bad magic");
    ReadsFile r;
    if (fread(&r.n, 4, 1, f) != 1) die("short header");
    int32_t p[4];
    if (fread(p, 4, 4, f) != 4) die("short header");
    r.min_overlap = p[0];
    r.rs = p[1];
    r.min_offset = p[2];
    r.reserved = p[3];
    read_vec(f, r.len, r.n);
    read_vec(f, r.align_from, r.n);
    read_vec(f, r.align_to, r.n);
    read_vec(f, r.word_off, (size_t) r.n + 1);
    read_vec(f, r.words, (size_t) r.word_off[r.n]);
    fclose(f);
    return r;
}

std::string decode(const ReadsFile &r, uint32_t i) {
    static const char nt[4] = {'A', 'C', 'G', 'T'};
    std::string s(r.len[i], 'A');
    const uint32_t *w = r.words.data() + r.word_off[i];
    for (uint32_t j = 0; j < r.len[i]; j++) s[j] = nt[(w[j >> 4] >> ((j & 15) * 2)) & 3];
    return s;
}

void build_reads(const ReadsFile &r) {
    Global::READS.assign(r.n, nullptr);
    for (uint32_t i = 0; i < r.n; i++) {
        if (r.len[i] == 0) continue;
        Global::READS[i] = new Read((int) i, decode(r, i));
    }
}

double now_s() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int run_prefsuf(const char *in_path, const char *out_path, int threads) {
    ReadsFile r = load_reads(in_path);
    Params::THREADS = threads;
    Params::MIN_OVERLAP_PREF_SUF = r.min_overlap;
    Params::REMOVE_SMALL_OVERLAP_EDGES_MIN_OVERLAP = r.rs;
    Params::MIN_OFFSET_FOR_ALIGNMENT = r.min_offset;
    build_reads(r);

    Global::GRAPH = Graph((int) r.n);
    Graph *G = &Global::GRAPH;

    double t0 = now_s();
    GraphCreator *gc = new GraphCreatorPrefSuf(&Global::READS, G, false);
    for (uint32_t i = 0; i < r.n; i++) {
        gc->setAlignFrom((int) i, r.align_from[i] != 0 && r.len[i] != 0);
        gc->setAlignTo((int) i, r.align_to[i] != 0 && r.len[i] != 0);
    }
    double t1 = now_s();
    gc->startAlignmentGraphCreation();
    static_cast<GraphCreatorPrefSuf *>(gc)->clear();
    delete gc;
    G->retainOnlySmallestOffset();
    double t2 = now_s();

    uint64_t E = 0;
    for (int i = 0; i < G->size(); i++) E += (*G)[i].size();
    if (out_path && strcmp(out_path, "-") != 0) {
        FILE *f = fopen(out_path, "wb");
        if (!f) die("cannot open output");
        fwrite("ALGE", 1, 4, f);
        uint32_t n = r.n;
        fwrite(&n, 4, 1, f);
        fwrite(&E, 8, 1, f);
        for (int i = 0; i < G->size(); i++) {
            for (auto &e : (*G)[i]) {
                int32_t t[3] = {i, e.first, e.second};
                fwrite(t, 4, 3, f);
            }
        }
        fclose(f);
    }
    printf("{\"n\": %u, \"edges\": %llu, \"threads\": %d, \"setup_s\": %.6f, \"graph_s\": %.6f}\n", r.n,
           (unsigned long long) E, threads, t1 - t0, t2 - t1);
    return 0;
}

// pairs file: "ALGP", u64 n, i32 thr (MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR), i32 max_offset_pct
// (MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT), i32 min_overlap_area (MIN_OVERLAP_AREA), i32 min_offset,
// then n x (i32 a, i32 b, i32 off)
int run_verify(const char *in_path, const char *pairs_path, const char *out_path, int lcs_rate = 0, int lcs_band = 2) {
    ReadsFile r = load_reads(in_path);
    Params::THREADS = 1;
    build_reads(r);
    FILE *f = fopen(pairs_path, "rb");
    if (!f) die("cannot open pairs file");
    char magic[4];
    uint64_t n;
    int32_t hp[4];
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ALGP", 4) != 0) die("bad pairs magic");
    if (fread(&n, 8, 1, f) != 1 || fread(hp, 4, 4, f) != 4) die("short pairs header");
    std::vector<int32_t> p(3 * n);
    if (n && fread(p.data(), 4, 3 * n, f) != 3 * n) die("short pairs");
    fclose(f);
    Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR = hp[0];
    Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT = hp[1];
    Params::MIN_OVERLAP_AREA = hp[2];
    Params::MIN_OFFSET_FOR_ALIGNMENT = hp[3];
    if (lcs_rate > 0) {  // pairs the low-error test rejects go on to AlignmentControllerLCS (AlignmentControllerHybrid.cpp:72-76)
        Params::USE_ACLER_INSTEAD_OF_ACLCS = 0;
        Params::MINIMAL_OVERLAP_RATE_FOR_LCS = lcs_rate;
        Params::MAX_ERROR_RATE_FOR_LCS = lcs_band;
    }
    AlignmentControllerHybrid ac;
    std::vector<uint8_t> verdict(n);
    for (uint64_t i = 0; i < n; i++) {
        verdict[i] = ac.canAlign(Global::READS[p[3 * i]], Global::READS[p[3 * i + 1]], p[3 * i + 2]) ? 1 : 0;
    }
    FILE *o = fopen(out_path, "wb");
    if (!o) die("cannot open verdict output");
    fwrite(verdict.data(), 1, n, o);
    fclose(o);
    printf("{\"pairs\": %llu}\n", (unsigned long long) n);
    return 0;
}

// Error-rate supplement (main.cpp:300-355) on a given read set and graph:
//   main.cpp:306      new GraphCreatorLI(READS, G)
//   main.cpp:308-323  flags from in-/out-degrees (dead ends only)
//   main.cpp:332-340  Params of the supplement (passed in: they derive from the average read length / error rate)
//   main.cpp:343-346  startAlignmentGraphCreation(); G->retainOnlySmallestOffset()
// edges_in: ALGE file (the graph after main.cpp:291); out: ALGE file of the graph after main.cpp:346.
int run_supplement(const char *reads_path, const char *edges_path, const char *out_path, int moa, int max_off, int thr,
                   int kmer_bucket_len, int threads) {
    ReadsFile r = load_reads(reads_path);
    Params::THREADS = threads;
    Params::MIN_OFFSET_FOR_ALIGNMENT = r.min_offset;
    build_reads(r);
    Global::GRAPH = Graph((int) r.n);
    Graph *G = &Global::GRAPH;
    {
        FILE *f = fopen(edges_path, "rb");
        if (!f) die("cannot open edges file");
        char magic[4];
        uint32_t n;
        uint64_t E;
        if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ALGE", 4) != 0) die("bad edges magic");
        if (fread(&n, 4, 1, f) != 1 || fread(&E, 8, 1, f) != 1 || n != r.n) die("bad edges header");
        std::vector<int32_t> t(3 * E);
        if (E && fread(t.data(), 4, 3 * E, f) != 3 * E) die("short edges");
        fclose(f);
        for (uint64_t i = 0; i < E; i++) G->pushDirectedEdge(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
    }
    double t0 = now_s();
    GraphCreator *graphCreator = new GraphCreatorLI(&Global::READS, G);
    VI *inDeg = G->getInDegrees();
    for (int i = 0; i < G->size(); i++) {
        graphCreator->setAlignFrom(i, false);
        graphCreator->setAlignTo(i, false);
        if ((*inDeg)[i] == 0 && (*G)[i].size() > 0) graphCreator->setAlignTo(i, true);
        if ((*inDeg)[i] > 0 && (*G)[i].size() == 0) graphCreator->setAlignFrom(i, true);
    }
    delete inDeg;
    Params::MIN_OVERLAP_AREA = moa;
    Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT = max_off;
    Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR = thr;
    Params::LI_KMER_INTERVALS = 6;
    Params::LI_KMER_LENGTH = 35;
    Params::KMER_LENGTH_BUCKET = kmer_bucket_len;
    graphCreator->startAlignmentGraphCreation();
    G->retainOnlySmallestOffset();
    delete graphCreator;
    double t1 = now_s();
    uint64_t E = 0;
    for (int i = 0; i < G->size(); i++) E += (*G)[i].size();
    FILE *f = fopen(out_path, "wb");
    if (!f) die("cannot open output");
    fwrite("ALGE", 1, 4, f);
    uint32_t n = r.n;
    fwrite(&n, 4, 1, f);
    fwrite(&E, 8, 1, f);
    for (int i = 0; i < G->size(); i++)
        for (auto &e : (*G)[i]) {
            int32_t t[3] = {i, e.first, e.second};
            fwrite(t, 4, 3, f);
        }
    fclose(f);
    printf("{\"n\": %u, \"edges\": %llu, \"threads\": %d, \"supplement_s\": %.6f}\n", r.n, (unsigned long long) E, threads,
           t1 - t0);
    return 0;
}

// ReadPreprocess::getPrefixReads (main.cpp:132-134) on a given read set: one byte per read, 1 = removed.
int run_prefix_reads(const char *reads_path, const char *out_path, int remove_type, int threads) {
    ReadsFile r = load_reads(reads_path);
    Params::THREADS = threads;
    Params::REMOVE_PREF_READS_TYPE = remove_type;
    build_reads(r);
    ReadPreprocess prepr;
    VB marked = prepr.getPrefixReads();
    std::vector<uint8_t> mask(r.n, 0);
    uint64_t removed = 0;
    for (uint32_t i = 0; i < r.n; i++) {
        mask[i] = marked[i] ? 1 : 0;
        removed += mask[i];
    }
    FILE *f = fopen(out_path, "wb");
    if (!f) die("cannot open mask output");
    fwrite(mask.data(), 1, r.n, f);
    fclose(f);
    printf("{\"n\": %u, \"removed\": %llu}\n", r.n, (unsigned long long) removed);
    return 0;
}

// InputReader::readInput (main.cpp:67-82) on one or two files: the reference's own option parser sets the file type from the
// extension (Params.cpp:332-335), its reader produces Global::READS; dumped as an ALGR file (nullptr = length 0).
int run_read_input(const char *file1, const char *file2, const char *out_path, int threads, int n_extra, char **extra) {
    std::vector<std::string> args = {"ALGA", std::string("--file1=") + file1};
    if (strcmp(file2, "-") != 0) args.push_back(std::string("--file2=") + file2);
    args.push_back("--threads=" + std::to_string(threads));
    args.push_back("--output=harness_out.fasta");
    for (int i = 0; i < n_extra; i++) args.push_back(extra[i]);  // e.g. --rna=1, --retl=.. (the reference's own option names)
    std::vector<char *> av;
    for (auto &a : args) av.push_back(const_cast<char *>(a.c_str()));
    av.push_back(nullptr);
    Params::initializeParams((int) args.size(), av.data());
    InputReader reader;
    reader.readInput();
    const uint32_t n = (uint32_t) Global::READS.size();
    std::vector<uint32_t> len(n, 0), words;
    std::vector<uint64_t> off(n + 1, 0);
    for (uint32_t i = 0; i < n; i++) {
        Read *r = Global::READS[i];
        if (r != nullptr) {
            if (r->getId() != (int) i) die("read id differs from its index");
            len[i] = (uint32_t) r->size();
            const uint32_t nb = (len[i] + 15) / 16;
            for (uint32_t b = 0; b < nb; b++) words.push_back(r->getSequence().getBlock((int) b));
        }
        off[i + 1] = words.size();
    }
    std::vector<uint8_t> ones(n, 1);
    FILE *f = fopen(out_path, "wb");
    if (!f) die("cannot open output");
    fwrite("ALGR", 1, 4, f);
    fwrite(&n, 4, 1, f);
    int32_t p[4] = {0, 0, 0, 0};
    fwrite(p, 4, 4, f);
    fwrite(len.data(), 4, n, f);
    fwrite(ones.data(), 1, n, f);
    fwrite(ones.data(), 1, n, f);
    fwrite(off.data(), 8, (size_t) n + 1, f);
    fwrite(words.data(), 4, words.size(), f);
    fclose(f);
    printf("{\"n\": %u}\n", n);
    return 0;
}

// First step of GraphSimplifier::simplifyGraphOld (GraphSimplifier.cpp:110-130) on a given graph:
// G->sortEdgesByIncreasingOffset(); simplifier.cutNonAndWeaklyMetricTriangles() with Params::MAX_OFFSET_PARALLEL_PATHS.
int run_triangles(const char *edges_path, const char *out_path, int max_offset, int threads) {
    Params::THREADS = threads;
    Params::MAX_OFFSET_PARALLEL_PATHS = max_offset;
    FILE *f = fopen(edges_path, "rb");
    if (!f) die("cannot open edges file");
    char magic[4];
    uint32_t n;
    uint64_t E;
    if (fread(magic, 1, 4, f) != 4 || memcmp(magic, "ALGE", 4) != 0) die("bad edges magic");
    if (fread(&n, 4, 1, f) != 1 || fread(&E, 8, 1, f) != 1) die("bad edges header");
    std::vector<int32_t> t(3 * E);
    if (E && fread(t.data(), 4, 3 * E, f) != 3 * E) die("short edges");
    fclose(f);
    Global::READS.assign(n, nullptr);
    Global::GRAPH = Graph((int) n);
    Graph *G = &Global::GRAPH;
    for (uint64_t i = 0; i < E; i++) G->pushDirectedEdge(t[3 * i], t[3 * i + 1], t[3 * i + 2]);
    {
        GraphSimplifier simplifier(Global::GRAPH, Global::READS);
        G->sortEdgesByIncreasingOffset();
        simplifier.cutNonAndWeaklyMetricTriangles();
    }
    uint64_t E2 = 0;
    for (int i = 0; i < G->size(); i++) E2 += (*G)[i].size();
    FILE *o = fopen(out_path, "wb");
    if (!o) die("cannot open output");
    fwrite("ALGE", 1, 4, o);
    fwrite(&n, 4, 1, o);
    fwrite(&E2, 8, 1, o);
    for (int i = 0; i < G->size(); i++)
        for (auto &e : (*G)[i]) {
            int32_t x[3] = {i, e.first, e.second};
            fwrite(x, 4, 3, o);
        }
    fclose(o);
    printf("{\"n\": %u, \"edges_in\": %llu, \"edges_out\": %llu}\n", n, (unsigned long long) E, (unsigned long long) E2);
    return 0;
}

// The whole unmodified driver (main.cpp:57-779) on the given reference options; afterwards Global::pairedReadOffset -- filled
// once by the renumbering at main.cpp:150-232 and only read later (Read.cpp:262) -- is written out: u32 n, n bytes.
int run_driver(const char *out_path, int argc, char **argv) {
    std::vector<char *> av;
    static char name[] = "ALGA";
    av.push_back(name);
    for (int i = 0; i < argc; i++) av.push_back(argv[i]);
    av.push_back(nullptr);
    const int rc = alga_reference_main((int) av.size() - 1, av.data());
    if (rc != 0) return rc;
    FILE *o = fopen(out_path, "wb");
    if (!o) die("cannot open output");
    const uint32_t n = (uint32_t) Global::pairedReadOffset.size();
    fwrite(&n, 4, 1, o);
    if (n) fwrite(Global::pairedReadOffset.data(), 1, n, o);
    fclose(o);
    printf("{\"paired_offsets\": %u}\n", n);
    return 0;
}

}  // namespace

int main(int argc, char **argv) {
    Read::priorities = {0, 1, 2, 3};
    Bitset::initializeStaticBlock();
    if (argc >= 4 && strcmp(argv[1], "prefsuf") == 0) {
        int threads = argc >= 5 ? atoi(argv[4]) : 1;
        return run_prefsuf(argv[2], argv[3], threads);
    }
    if (argc >= 5 && strcmp(argv[1], "verify") == 0)
        return run_verify(argv[2], argv[3], argv[4], argc >= 6 ? atoi(argv[5]) : 0, argc >= 7 ? atoi(argv[6]) : 2);
    if (argc >= 4 && strcmp(argv[1], "prefixreads") == 0)
        return run_prefix_reads(argv[2], argv[3], argc >= 5 ? atoi(argv[4]) : 2, argc >= 6 ? atoi(argv[5]) : 1);
    if (argc >= 9 && strcmp(argv[1], "supplement") == 0)
        return run_supplement(argv[2], argv[3], argv[4], atoi(argv[5]), atoi(argv[6]), atoi(argv[7]), atoi(argv[8]),
                              argc >= 10 ? atoi(argv[9]) : 1);
    if (argc >= 4 && strcmp(argv[1], "driver") == 0) return run_driver(argv[2], argc - 3, argv + 3);
    if (argc >= 5 && strcmp(argv[1], "triangles") == 0)
        return run_triangles(argv[2], argv[3], atoi(argv[4]), argc >= 6 ? atoi(argv[5]) : 1);
    if (argc >= 5 && strcmp(argv[1], "readinput") == 0)
        return run_read_input(argv[2], argv[3], argv[4], argc >= 6 ? atoi(argv[5]) : 1, argc > 6 ? argc - 6 : 0, argv + 6);
    fprintf(stderr,
            "usage: %s prefsuf <reads.algr> <edges.alge|-> [threads]\n"
            "       %s verify  <reads.algr> <pairs.algp> <verdict.bin> [lcs_rate_pct [lcs_band]]\n"
            "       %s supplement <reads.algr> <edges_in.alge> <edges_out.alge> <min_overlap_area> <max_offset_pct> "
            "<threshold_pct> <kmer_length_bucket> [threads]\n"
            "       %s prefixreads <reads.algr> <mask.bin> [remove_type 1|2] [threads]\n"
            "       %s readinput <file1> <file2|-> <reads_out.algr> [threads [reference options, e.g. --rna=1]]\n"
            "       %s triangles <edges_in.alge> <edges_out.alge> <max_offset_parallel_paths> [threads]\n"
            "       %s driver <paired_offsets_out.bin> <reference options: --file1=... --threads=1 --output=...>\n",
            argv[0], argv[0], argv[0], argv[0], argv[0], argv[0], argv[0]);
    return 2;
}

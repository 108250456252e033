/* TEST INFRASTRUCTURE ONLY -- see oracle.h.
 *
 * CPU restatement of the error-rate supplement of swacisko/ALGA (runs when --error_rate > 0.01), in the order of
 * the reference at --threads=1:
 *   main.cpp:300-346                      driver: dead-end reads only, Params of the supplement, final row dedupe
 *   GraphCreatorLI.cpp:18-28              four passes, Read::priorities rotated left after each
 *   GraphCreatorKmerBased.cpp:202-259     LI k-mers of the flagged reads scattered into 2^20 hash-range buckets
 *   Read.cpp:145-226                      getLIKmers: per interval the leftmost minimal k-mer under the priorities
 *   GraphCreatorKmerBased.cpp:94-136      std::sort of every bucket (Kmer.cpp:58-64), walk over equal-hash runs
 *   GraphCreatorPairwiseKmerBranch.cpp:16-97   ordered pair loop with branch markers, Graph::addDirectedEdge
 *   AlignmentControllerHybrid.cpp:46-83   canAlign (verify_oracle.c)
 *   GraphCreatorKmerBased.cpp:87 / Graph.cpp:348-387   retainOnlySmallestOffset after every pass
 *
 * C++ because the tie order inside a bucket is whatever libstdc++'s std::sort leaves: this file sorts
 * the same sequence with the same comparator through the same std::sort, which is what pins it to the reference
 * built by the same toolchain (oracle/_ref).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "oracle.h"

namespace {

typedef unsigned __int128 u128;
const uint64_t kMaxHash = 1000000000000000003ull;  // Params::MAX_HASH_CONSIDERED (Params.cpp:721)
const int kInf = 1000000001;                       // Params::INF
const long long kBucketsSort = 1048576ll;          // GraphCreatorKmerBased.cpp:140

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

inline int nt(const oracle_reads *r, uint32_t i, uint32_t j) {
    return (int) ((r->words[r->word_off[i] + (j >> 4)] >> ((j & 15) * 2)) & 3u);
}

// Read::getLIKmers (Read.cpp:145-226)
void li_kmers(const oracle_reads *r, uint32_t id, const int prio[4], int length, int intervals, std::vector<Kmer> &out) {
    const int size = (int) r->len_nt[id];
    int p = 0, q = 0;
    u128 hash = 0;
    while (q < length) {
        hash <<= 2;
        hash += (u128) prio[nt(r, id, (uint32_t) q)];
        q++;
    }
    u128 factor = 1;
    for (int i = 0; i < length - 1; i++) factor <<= 2;
    std::vector<u128> min_hash((size_t) intervals, factor << 2);
    std::vector<Kmer> min_kmer((size_t) intervals, Kmer{id, 0, 0, 0});  // read_len == 0 <=> "size() == 0"
    min_hash[0] = hash;
    min_kmer[0] = Kmer{id, (uint64_t) (hash % (u128) kMaxHash), p, (uint32_t) size};
    const int interval_len = (int) std::ceil(((double) size - length + 1) / intervals);
    int interv = 0;
    while (q < size) {
        hash -= factor * (u128) prio[nt(r, id, (uint32_t) p)];
        hash <<= 2;
        hash += (u128) prio[nt(r, id, (uint32_t) q)];
        p++;
        q++;
        interv = p / interval_len;
        if (hash < min_hash[(size_t) interv]) {
            min_hash[(size_t) interv] = hash;
            min_kmer[(size_t) interv] = Kmer{id, (uint64_t) (hash % (u128) kMaxHash), p, (uint32_t) size};
        }
    }
    while ((int) min_kmer.size() > interv + 1) min_kmer.pop_back();
    for (int j = (int) min_kmer.size() - 1; j >= 0; j--) {
        if (min_kmer[(size_t) j].read_len == 0) {
            std::swap(min_kmer[(size_t) j], min_kmer.back());
            min_kmer.pop_back();
        }
    }
    for (const Kmer &k : min_kmer) out.push_back(k);
}

struct Graph {
    std::vector<std::vector<std::pair<int, int>>> V;
    void add_directed_edge(int a, int b, int off) {  // Graph.cpp:53-71
        if (a == b) return;
        for (auto &e : V[(size_t) a]) {
            if (e.first == b) {
                if (off < e.second) e.second = off;
                return;
            }
        }
        V[(size_t) a].push_back({b, off});
    }
    void retain_only_smallest_offset() {  // Graph.cpp:348-387
        for (auto &row : V) {
            std::sort(row.begin(), row.end());
            size_t w = 0;
            for (size_t k = 0; k < row.size(); k++)
                if (w == 0 || row[w - 1].first != row[k].first) row[w++] = row[k];
            row.resize(w);
        }
    }
};

}  // namespace

extern "C" void oracle_li_kmers(const oracle_reads *r, const uint32_t *ids, uint32_t n_ids, const int32_t *prio, int32_t K,
                                int32_t intervals, uint64_t *hash_out, int32_t *ind_out) {
    const int pr[4] = {prio[0], prio[1], prio[2], prio[3]};
    std::vector<Kmer> tmp;
    for (uint32_t q = 0; q < n_ids; q++) {
        tmp.clear();
        li_kmers(r, ids[q], pr, K, intervals, tmp);
        for (int iv = 0; iv < intervals; iv++) {
            const size_t slot = (size_t) q * (size_t) intervals + (size_t) iv;
            hash_out[slot] = (size_t) iv < tmp.size() ? tmp[(size_t) iv].hash : 0;
            ind_out[slot] = (size_t) iv < tmp.size() ? tmp[(size_t) iv].ind : -1;
        }
    }
}

extern "C" int32_t *oracle_supplement(const oracle_reads *r, const int32_t *edges_in, uint64_t n_in,
                                      const oracle_sup_params *sp, uint64_t *n_out) {
    const uint32_t n = r->n;
    Graph G;
    G.V.resize(n);
    for (uint64_t i = 0; i < n_in; i++) G.V[(size_t) edges_in[3 * i]].push_back({edges_in[3 * i + 1], edges_in[3 * i + 2]});
    // main.cpp:308-323 -- flags, computed once
    std::vector<int> indeg(n, 0);
    for (uint32_t i = 0; i < n; i++)
        for (auto &e : G.V[i]) indeg[(size_t) e.first]++;
    std::vector<uint8_t> use(n, 0);
    for (uint32_t i = 0; i < n; i++) {
        const bool to = indeg[i] == 0 && !G.V[i].empty(), from = indeg[i] > 0 && G.V[i].empty();
        use[i] = (to || from) ? 1 : 0;
    }
    oracle_verify_params vp;
    vp.max_offset_pct = sp->max_offset_pct;
    vp.min_offset = sp->min_offset;
    vp.min_overlap_area = sp->min_overlap_area;
    vp.threshold_pct = sp->threshold_pct;
    vp.same_ends = sp->same_ends;

    int prio[4] = {0, 1, 2, 3};
    std::vector<int> neighbors(n, kInf);
    for (int pass = 0; pass < 4; pass++) {  // GraphCreatorLI.cpp:20-26
        // GraphCreatorKmerBased.cpp:202-259 (one thread: reads in id order, k-mers in interval order)
        std::vector<std::vector<Kmer>> buckets((size_t) kBucketsSort);
        std::vector<Kmer> tmp;
        const double A = 0.0, B = (double) kMaxHash;  // blockSize * (bucket + 1) - 1 with bucket 0
        for (uint32_t i = 0; i < n; i++) {
            if (r->len_nt[i] == 0 || !use[i]) continue;
            if (sp->kmer_length_bucket > (int32_t) r->len_nt[i]) continue;  // Read::getKmers (Read.cpp:70-72)
            tmp.clear();
            li_kmers(r, i, prio, sp->kmer_length, sp->intervals, tmp);
            for (const Kmer &k : tmp) {
                const int ind = (int) ((kBucketsSort - 1) * (((double) k.hash - A) / (B - A)));
                buckets[(size_t) ind].push_back(k);
            }
        }
        for (auto &b : buckets)
            if (!b.empty()) std::sort(b.begin(), b.end());  // GraphCreatorKmerBased.cpp:94-106
        // GraphCreatorKmerBased.cpp:108-136 + GraphCreatorPairwiseKmerBranch.cpp:16-97
        std::vector<std::vector<uint8_t>> bm;
        for (auto &km : buckets) {
            size_t p = 0, q = 0;
            while (p < km.size()) {
                while (q < km.size() && km[q].hash == km[p].hash) q++;
                const int D = (int) (q - p);
                bm.assign((size_t) D, std::vector<uint8_t>((size_t) D, 0));
                for (int i = D - 2; i >= 0; i--) {
                    const Kmer &ki = km[p + (size_t) i];
                    const int id1 = (int) ki.read;
                    for (auto &x : G.V[(size_t) id1]) neighbors[(size_t) x.first] = x.second;
                    for (int j = i + 1; j < D; j++) {
                        const Kmer &kj = km[p + (size_t) j];
                        const int id2 = (int) kj.read;
                        if (id1 == id2) continue;
                        const int offset = ki.ind - kj.ind;
                        if (offset < sp->min_offset) continue;
                        if (100ll * offset > (long long) sp->max_offset_pct * (long long) ki.read_len) break;
                        const int overlap = std::min((int) ki.read_len, (int) kj.read_len + offset) - offset;
                        if (overlap < sp->min_overlap_area) continue;
                        if ((int) kj.read_len + offset - (int) ki.read_len < 0) continue;  // Read::getRightOffset
                        if (!bm[(size_t) i][(size_t) j]) {
                            if (neighbors[(size_t) id2] > offset) {
                                const int32_t pr[3] = {id1, id2, offset};
                                uint8_t v = 0;
                                oracle_verify_pairs(r, pr, 1, &vp, &v);
                                if (v) {
                                    G.add_directed_edge(id1, id2, offset);
                                    neighbors[(size_t) id2] = offset;
                                }
                            }
                            if (neighbors[(size_t) id2] != kInf) {
                                bm[(size_t) i][(size_t) j] = 1;
                                for (int t = 0; t < D; t++) bm[(size_t) i][(size_t) t] |= bm[(size_t) j][(size_t) t];
                            }
                        }
                    }
                    for (auto &x : G.V[(size_t) id1]) neighbors[(size_t) x.first] = kInf;
                }
                p = q;
            }
        }
        G.retain_only_smallest_offset();  // GraphCreatorKmerBased.cpp:87
        std::rotate(prio, prio + 1, prio + 4);
    }
    G.retain_only_smallest_offset();  // main.cpp:346
    uint64_t E = 0;
    for (auto &row : G.V) E += row.size();
    int32_t *out = (int32_t *) malloc((size_t) (E ? E : 1) * 12);
    uint64_t w = 0;
    for (uint32_t i = 0; i < n; i++)
        for (auto &e : G.V[i]) {
            out[3 * w] = (int32_t) i;
            out[3 * w + 1] = e.first;
            out[3 * w + 2] = e.second;
            w++;
        }
    *n_out = E;
    return out;
}

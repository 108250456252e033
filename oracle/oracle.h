/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's hot path.
 *
 * Plain-C restatement of swacisko/ALGA's GraphCreatorPrefSuf (exact prefix/suffix overlaps with
 * on-the-fly transitive reduction) and of AlignmentControllerHybrid::canAlign, in the canonical
 * single-threaded order of the reference.  Pinned against the unmodified reference itself
 * (oracle/_ref/alga_ref_harness; fixtures under tests/golden/).  Only tests/, smoke() and the CPU
 * legs of bench.py may link or load this; the product library never does.
 */
#ifndef ALGA_ORACLE_H
#define ALGA_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    uint32_t n;               /* number of graph nodes (strand-reads) */
    const uint32_t *words;    /* packed 2-bit blocks, reference layout (Read.cpp:40-68) */
    const uint64_t *word_off; /* n+1 offsets into words */
    const uint32_t *len_nt;   /* n lengths, 0 = removed read */
    const uint8_t *align_from, *align_to; /* n flags (GraphCreator.h:46-62) */
} oracle_reads;

/* GraphCreatorPrefSuf.cpp:73-126 followed by Graph::retainOnlySmallestOffset (main.cpp:291).
 * Returns a malloc'ed array of (src, dst, offset) int32 triples sorted by (src, dst); *n_edges = count.
 * Free with oracle_free. */
int32_t *oracle_prefsuf(const oracle_reads *r, int32_t min_overlap, int32_t rs_min_overlap, int32_t min_offset,
                        int32_t max_len_cap, uint64_t *n_edges);

/* Fingerprints of the length-L prefix and suffix of every read (GraphCreatorPrefSuf.cpp:213-236):
 * h64 = sum s_j 4^j mod 10^18+3, h32 = sum s_j 4^j mod 10^9+7.  Entries of reads shorter than L are
 * left untouched.  Arrays have n entries each. */
void oracle_fingerprints(const oracle_reads *r, int32_t L, uint64_t *pre64, uint32_t *pre32, uint64_t *suf64,
                         uint32_t *suf32);

/* AlignmentControllerHybrid.cpp:46-83 -> AlignmentControllerLowErrorRate.cpp:15-49 for n_pairs
 * (a, b, offset) triples. verdict[i] in {0,1}. */
typedef struct {
    int32_t max_offset_pct;   /* Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT */
    int32_t min_offset;       /* Params::MIN_OFFSET_FOR_ALIGNMENT */
    int32_t min_overlap_area; /* Params::MIN_OVERLAP_AREA */
    int32_t threshold_pct;    /* Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR */
    int32_t same_ends;        /* Params::ALIGNMENT_CONTROLLER_SAME_ENDS_LENGTH (3) */
    int32_t lcs_rate_pct;     /* 0: USE_ACLER_INSTEAD_OF_ACLCS = 1 (the default); > 0: USE_ACLER_INSTEAD_OF_ACLCS = 0 and this is
                                 Params::MINIMAL_OVERLAP_RATE_FOR_LCS -- pairs the low-error test rejects go on to the banded LCS */
    int32_t lcs_band;         /* Params::MAX_ERROR_RATE_FOR_LCS (2) */
} oracle_verify_params;

void oracle_verify_pairs(const oracle_reads *r, const int32_t *pairs, uint64_t n_pairs,
                         const oracle_verify_params *p, uint8_t *verdict);

/* Error-rate supplement (main.cpp:300-355: GraphCreatorLI over the dead-end reads of the graph after main.cpp:291,
 * followed by retainOnlySmallestOffset).  edges_in / result: (src, dst, offset) triples, rows in id order.
 * Implemented in supplement_oracle.cpp (C++: bucket tie order = libstdc++ std::sort, as in the reference). */
typedef struct {
    int32_t max_offset_pct;     /* Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT = (1 - SCALE) * avg_len / 2 (main.cpp:335) */
    int32_t min_offset;         /* Params::MIN_OFFSET_FOR_ALIGNMENT */
    int32_t min_overlap_area;   /* Params::MIN_OVERLAP_AREA = (1 + SCALE) * avg_len / 2 (main.cpp:333) */
    int32_t threshold_pct;      /* Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR = 99 - ERROR_RATE (main.cpp:336) */
    int32_t same_ends;          /* Params::ALIGNMENT_CONTROLLER_SAME_ENDS_LENGTH (3) */
    int32_t kmer_length;        /* Params::LI_KMER_LENGTH = 35 (main.cpp:340) */
    int32_t intervals;          /* Params::LI_KMER_INTERVALS = 6 (main.cpp:339) */
    int32_t kmer_length_bucket; /* Params::KMER_LENGTH_BUCKET (main.cpp:104): shorter reads contribute no k-mers */
} oracle_sup_params;
/* Read::getLIKmers (Read.cpp:145-226) of the reads ids[0 .. n_ids): `intervals` slots each, ind = -1 where absent. */
void oracle_li_kmers(const oracle_reads *r, const uint32_t *ids, uint32_t n_ids, const int32_t *prio, int32_t K,
                     int32_t intervals, uint64_t *hash_out, int32_t *ind_out);
int32_t *oracle_supplement(const oracle_reads *r, const int32_t *edges_in, uint64_t n_in, const oracle_sup_params *p,
                           uint64_t *n_out);

/* ReadPreprocess::getPrefixReads (ReadPreprocess.cpp:13-77): mask[i] = 1 for reads that are removed -- duplicates except
 * the one with the greatest id and, with remove_type 2 (the default), reads that are a proper prefix of another read
 * together with their reverse complements (id ^ 1).  remove_type 1 = duplicates only. */
void oracle_prefix_reads(const oracle_reads *r, int32_t remove_type, uint8_t *mask);

/* InputReader::readInput (InputReader.cpp:44-139, --threads=1 order) on one or two in-memory files: both strands of every
 * record in the reference's id order (reverse complement at the even id, mates interleaved), removed reads (N, minimal
 * period <= str_threshold) as length 0.  Outputs are malloc'ed (oracle_free).  See input_oracle.c. */
#define ORACLE_INPUT_PLAIN 0 /* Params::MY_INPUT: whitespace-separated sequences */
#define ORACLE_INPUT_FASTA 1 /* Params::FASTA (and PFASTA with paired reads): header line + sequence line */
#define ORACLE_INPUT_FASTQ 2 /* Params::FASTQ: four lines per record */
#define ORACLE_E_BADCHAR (-1) /* a character other than A C G T N U (the reference exits, InputReader.cpp:324-327) */
#define ORACLE_E_PAIRING (-2) /* mate files with different record counts / a read without its reverse complement */
typedef struct {
    int32_t file_type;
    int32_t trim_left, trim_right; /* Params::READ_END_TRIM_LEFT / RIGHT (3, 3) */
    int32_t rna;                   /* Params::RNA: U -> T */
    int32_t str_threshold;         /* 20, InputReader.cpp:343; <= 0 means 20 */
} oracle_input_params;
int oracle_read_input(const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2, const oracle_input_params *p,
                      uint32_t *n_reads, uint32_t **len_nt, uint64_t **word_off, uint32_t **words, uint64_t *n_with_n,
                      uint64_t *n_str);
/* main.cpp:133-140 + 150-232: compaction of the surviving reads.  old_id[k] = previous id of new read k, paired_offset[k]
 * = Global::pairedReadOffset[k]; both need room for n entries. */
int oracle_remap(const uint32_t *len_nt, const uint8_t *mask, uint32_t n, uint32_t *old_id, uint8_t *paired_offset,
                 uint32_t *n_out);

/* Graph::sortEdgesByIncreasingOffset + GraphSimplifier::cutNonAndWeaklyMetricTriangles (GraphSimplifier.cpp:228-349) on
 * (src, dst, offset) triples grouped by ascending src; result sorted by (src, offset, dst).  See simplify_oracle.c. */
int32_t *oracle_cut_triangles(const int32_t *edges_in, uint64_t n_in, uint32_t n_nodes, int32_t max_offset, uint64_t *n_out);

void oracle_free(void *p);

#ifdef __cplusplus
}
#endif
#endif

/* alga_gpu.h -- C ABI of libalga_gpu.so: the B200 (sm_100a) overlap-graph hot path of ALGA.
 *
 * Every entry point replaces one piece of the reference's CPU hot path and is what a binding of
 * that path (the C++ shim in shim/, cgo, JNI, ctypes ...) would call.  Plain pointers and sizes
 * only; no CUDA or torch types.  All functions return 0 on success and a negative ALGA_E_* code
 * on failure; alga_gpu_last_error() returns a thread-local message.  There is no CPU fallback:
 * without a CUDA device every compute entry point fails with ALGA_E_CUDA.
 *
 * Packed read layout = the reference's Bitset/Read layout (include/DataStructures/Bitset.h:38-45,
 * src/DataStructures/Read.cpp:40-68): A=0 C=1 G=2 T=3, nucleotide j in bits [2(j%16), 2(j%16)+1]
 * of 32-bit block j/16, tail bits zero.  len_nt[i] == 0 marks a removed (nullptr) read.
 */
#ifndef ALGA_GPU_H
#define ALGA_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ALGA_OK 0
#define ALGA_E_INVALID (-1)  /* bad argument */
#define ALGA_E_CUDA (-2)     /* CUDA runtime error / no device */
#define ALGA_E_NOMEM (-3)    /* host or device allocation failed */
#define ALGA_E_CAPACITY (-4) /* internal buffer overflow that a retry could not fix */

/* ---- read set --------------------------------------------------------------------------- */
/* Replaces the vector<Read*>* argument of GraphCreator::GraphCreator (GraphCreator.cpp:9-17) plus the
 * alignFrom/alignTo bit vectors (GraphCreator.h:46-62).  Host or device pointers depending on the
 * entry point.  word_off may be NULL when every read occupies exactly stride_words blocks
 * (read i starts at block i*stride_words). */
typedef struct {
    uint32_t n_reads;           /* N = G->size() */
    const uint32_t *words;      /* concatenated Bitset blocks (Bitset.h:175 getBlock) */
    const uint64_t *word_off;   /* N+1 block offsets, or NULL with stride_words > 0 */
    uint32_t stride_words;      /* blocks per read when word_off == NULL */
    const uint32_t *len_nt;     /* N read lengths in nucleotides (Read::size), 0 = nullptr read */
    const uint8_t *align_from;  /* N flags, may be NULL (= all true) -- GraphCreator::setAlignFrom */
    const uint8_t *align_to;    /* N flags, may be NULL (= all true) -- GraphCreator::setAlignTo */
} alga_reads;

/* ---- GraphCreatorPrefSuf ---------------------------------------------------------------- */
/* The Params statics GraphCreatorPrefSuf reads: */
typedef struct {
    int32_t min_overlap;      /* Params::MIN_OVERLAP_PREF_SUF  (main.cpp:103-107) */
    int32_t rs_min_overlap;   /* Params::REMOVE_SMALL_OVERLAP_EDGES_MIN_OVERLAP (main.cpp:108) */
    int32_t min_offset;       /* Params::MIN_OFFSET_FOR_ALIGNMENT (Params.cpp:709), default 0 */
    int32_t max_len_cap;      /* overlap length cap, 500 in GraphCreatorPrefSuf.cpp:92; <=0 means 500 */
    int32_t device;           /* CUDA device ordinal */
    int32_t list_cap;         /* testing: > 0 runs phase 2 on the generic kernel with this on-chip in-neighbour list
                                 capacity per target read (targets that exceed it take the global-memory path);
                                 <= 0 = the fast kernel */
    int32_t flags;            /* ALGA_PS_* bits */
} alga_ps_params;
#define ALGA_PS_FORCE_GENERIC 1 /* testing: every read takes the generic (fallback) kernels */

/* Forward adjacency in CSR form: row b lists (nbr[k], off[k]) for row_off[b] <= k < row_off[b+1],
 * meaning "read nbr starts at position off of read b" -- exactly what Graph::V[b] holds after
 * main.cpp:282-291 (startAlignmentGraphCreation + retainOnlySmallestOffset): rows sorted by nbr,
 * one entry per nbr.  Library-allocated (host); release with alga_gpu_free_csr. */
typedef struct {
    uint32_t n_reads;
    uint64_t n_edges;
    uint64_t *row_off; /* N+1 */
    int32_t *nbr;      /* E */
    int32_t *off;      /* E */
    int32_t borrowed;  /* != 0: the arrays are page-locked staging buffers owned by the library, valid until the
                          next build on the same plan (alga_gpu_prefsuf_build: the next call); alga_gpu_free_csr
                          then only clears the pointers */
} alga_csr;

typedef struct {
    double h2d_ms;     /* host -> device copies */
    double device_ms;  /* CUDA-event time: packed reads resident -> CSR resident */
    double d2h_ms;     /* device -> host copies */
    double total_ms;   /* wall time of the call */
    uint64_t kernel_launches;
    uint64_t n_spilled_targets; /* targets that took the global-memory list path */
    /* CUDA-event time of each stage of alga_ps_plan_run on its stream (sums to ~device_ms):
     * [0] seed index build, [1] phase 1 (L < rs), [2] transpose of the phase-1 graph,
     * [3] phase 2 (L >= rs, transitive reduction), [4] CSR assembly; diagnostics (counts, not times): [5] phase-1
     * edges that overflowed their fixed-capacity row, [6] source reads that took the generic phase-1 kernel;
     * [7] reserved (0). */
    double stage_ms[8];
} alga_timing;
#define ALGA_STAGE_INDEX 0
#define ALGA_STAGE_PHASE1 1
#define ALGA_STAGE_TRANSPOSE 2
#define ALGA_STAGE_PHASE2 3
#define ALGA_STAGE_CSR 4

/* One-call drop-in for GraphCreatorPrefSuf::startAlignmentGraphCreation (GraphCreatorPrefSuf.cpp:73-126)
 * followed by Graph::retainOnlySmallestOffset (main.cpp:291).  Host buffers in, host CSR out (borrowed page-locked
 * buffers of a process-wide cached plan: valid until the next call; release with alga_gpu_free_csr). */
int alga_gpu_prefsuf_build(const alga_reads *reads, const alga_ps_params *params, alga_csr *out,
                           alga_timing *timing /* may be NULL */);
void alga_gpu_free_csr(alga_csr *csr);

/* The same build on up to n_gpus GPUs of the box, from ONE process (one host thread per GPU, peer memory over NVLink
 * between them): what the reference's single-process driver binds to use more than one GPU (shim: ALGA_GPU_DEVICES=<n>).
 * Rank r owns the reads [r * n_shard, (r + 1) * n_shard); the stages are those of alga_ps_shard_* below.  The sharded kernels
 * take equal-length reads at a fixed stride, none removed, no flags cleared -- what main.cpp:249-282 hands over for
 * equal-length input; any other read set (and n_gpus <= 1, and fewer than 4096 reads) is built on params->device alone.
 * The result arrays are malloc'ed (not borrowed); timing->stage_ms[7] = number of GPUs used. */
int alga_gpu_prefsuf_build_multi(const alga_reads *reads, const alga_ps_params *params, int32_t n_gpus, alga_csr *out,
                                 alga_timing *timing /* may be NULL */);

/* ---- staged, device-resident interface (bench, multi-GPU harness) ------------------------- */
/* A plan owns the device workspace of one GPU.  Device pointers passed in are borrowed. */
typedef struct alga_ps_plan alga_ps_plan;

int alga_ps_plan_create(alga_ps_plan **plan, const alga_ps_params *params);
void alga_ps_plan_destroy(alga_ps_plan *plan);
/* Bind a read set that is already resident on the plan's device (all pointers are device pointers;
 * `words` must be followed by at least 16 readable bytes).  max_len_nt = longest read (calculateMaxReadLength,
 * GraphCreatorPrefSuf.cpp:52-56); pass 0 to have it computed on the device. */
int alga_ps_plan_bind_reads_device(alga_ps_plan *plan, const alga_reads *dev_reads, uint32_t max_len_nt);
/* Copy a host read set to the device and bind it. */
int alga_ps_plan_upload_reads(alga_ps_plan *plan, const alga_reads *host_reads);

/* Whole single-GPU pipeline on `stream` (a cudaStream_t passed as void*, NULL = default stream):
 * seed index build, phase 1 (L < rs_min_overlap), transpose, phase 2 (L >= rs_min_overlap) with
 * transitive reduction, CSR assembly.  Asynchronous except for a few scalar read-backs. */
int alga_ps_plan_run(alga_ps_plan *plan, void *stream);

/* Stages of the same pipeline for key-range / id-range sharding across GPUs (one process per GPU;
 * the exchange between stages is done by the caller, e.g. with torch.distributed all_to_all over NCCL).
 * Ranges are read-id ranges [lo, hi) owned by this rank. */
int alga_ps_stage_index(alga_ps_plan *plan, void *stream);
/* phase 1 for suffix reads b in [lo,hi): emits (b, c, offset) int32 triples, device-resident.  */
int alga_ps_stage_phase1(alga_ps_plan *plan, uint32_t lo, uint32_t hi, void *stream,
                         const int32_t **dev_triples, uint64_t *n_triples);
/* phase 2 for target reads c in [lo,hi) given the phase-1 triples whose target lies in [lo,hi)
 * (device pointer, any order); emits the surviving (b, c, offset) triples, device-resident. */
int alga_ps_stage_phase2(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const int32_t *dev_triples_in,
                         uint64_t n_in, void *stream, const int32_t **dev_triples_out, uint64_t *n_out);
/* CSR assembly for source reads b in [lo,hi) from (b, c, offset) triples (device pointer, any order).
 * With swap_direction != 0 the triples are inserted as (c -> b) instead (the RS > maxL+1 corner of the
 * reference, SURVEY.md A.1 note 2). */
int alga_ps_stage_csr(alga_ps_plan *plan, uint32_t lo, uint32_t hi, const int32_t *dev_triples, uint64_t n,
                      int swap_direction, void *stream);

/* ---- sharded build with the exchange in the kernels (peer memory over NVLink / NVSwitch) ----------------------
 * One process per GPU, up to 8 GPUs of one box.  Rank r owns the reads [r * n_shard, min((r+1) * n_shard, n_total)).
 * Every rank holds all packed reads (bound with alga_ps_plan_bind_reads_*) and the whole seed index, and owns one
 * exchange workspace of alga_ps_shard_ws_bytes() bytes that every other rank can address (peer-mapped memory, e.g.
 * torch.distributed._symmetric_memory); peer_ws[p] is this process's pointer to rank p's workspace.
 *
 *   alga_ps_shard_phase1   phase 1 for the rank's own source reads; every edge (b, c, offset, overhang tail) is
 *                          appended to the segment of the rank that owns c, in the rank's OWN workspace
 *   -- barrier across ranks (the caller's, stream-ordered) --
 *   alga_ps_shard_phase2   reads the segments addressed to this rank out of all workspaces (NVLink loads inside the
 *                          kernel, no staging copy), builds the rows of the transposed graph, runs phase 2 for the
 *                          rank's own target reads; survivors (a, c, offset) go to the segment of the owner of a
 *   -- barrier across ranks --
 *   alga_ps_shard_csr      reads the survivors addressed to this rank and assembles its CSR rows
 *                          (alga_ps_plan_result_*: rows of [lo, hi), row_off relative to lo)
 *
 * The next build may start right away: its first barrier also orders it after the peers' reads of this one. */
typedef struct {
    int32_t rank, world;
    uint32_t n_shard;  /* reads owned per rank (the last rank may own fewer) */
    uint32_t n_total;  /* = n_reads of the bound read set */
    void *peer_ws[8];  /* [world] device pointers, peer_ws[rank] = this rank's own workspace */
    /* this rank's copies of the two seed tables (alga_ps_shard_table_bytes() each, caller-owned and readable by the
     * peers) -- only for alga_ps_shard_index_range, NULL otherwise */
    void *table_prefix, *table_suffix;
} alga_ps_shard;
uint64_t alga_ps_shard_ws_bytes(uint32_t n_shard, int32_t world);
/* Sharded index build: the bucket space of each table is cut into `world` slices; rank r inserts -- out of the reads
 * [lo, hi) of the (replicated) read set -- the seeds that fall into slice r of ITS copy of the tables.  `first`: bit 0 set =
 * clear the slice(s) first; bits 1-2 = which table: 0 both, 1 the prefix table only, 2 the suffix table only (a driver that
 * finishes the prefix slices first can exchange them while the suffix slices are still being filled: phase 1 needs the
 * prefix table only).  When every rank has inserted all reads, slice r of rank r's tables is final and the
 * ranks copy each other's slices (bytes [r, r+1) * table_bytes / world of the table buffers); the plan then probes
 * the caller's tables.  Replaces alga_ps_stage_index_range for sharded runs: no rank inserts more than its share. */
uint64_t alga_ps_shard_table_bytes(uint32_t n_total, int32_t world);
int alga_ps_shard_index_range(alga_ps_plan *plan, const alga_ps_shard *shard, uint32_t lo, uint32_t hi, int first,
                              void *stream);
/* The same build with the seeds computed once per read instead of once per read and rank (equal-length reads):
 * alga_ps_shard_seed_keys writes one 12-byte record per read of `shard_words` (n_reads reads at stride_words 32-bit words,
 * the caller's layout, e.g. this rank's own shard) to `keys` -- {bucket on the prefix side, bucket on the suffix side, the two
 * 16-bit tags} -- which the caller makes readable by the peers; alga_ps_shard_index_keys inserts, out of the records `keys`
 * of the reads [lo, hi), the seeds that fall into this rank's slice (first != 0 clears the slice first). */
int alga_ps_shard_seed_keys(alga_ps_plan *plan, const alga_ps_shard *shard, const uint32_t *shard_words, uint32_t stride_words,
                            uint32_t n_reads, uint32_t *keys, void *stream);
int alga_ps_shard_index_keys(alga_ps_plan *plan, const alga_ps_shard *shard, const uint32_t *keys, uint32_t lo, uint32_t hi,
                             int first, void *stream);
/* Mean entries per 20-slot bucket of the seed tables built from now on (1 .. 12; 0 = the default, 3).  Sharded builds ship
 * their table slices over NVLink and may prefer denser tables; call it on every rank before alga_ps_shard_table_bytes. */
void alga_ps_set_bucket_load(int32_t load);
int alga_ps_shard_phase1(alga_ps_plan *plan, const alga_ps_shard *shard, void *stream);
int alga_ps_shard_phase2(alga_ps_plan *plan, const alga_ps_shard *shard, void *stream);
int alga_ps_shard_csr(alga_ps_plan *plan, const alga_ps_shard *shard, void *stream);
/* Bind device-resident equal-length reads (fixed stride, no flags, none removed) WITHOUT a pass over them: they need
 * not be resident yet, e.g. while the other ranks' shards are still arriving. */
int alga_ps_plan_bind_reads_uniform(alga_ps_plan *plan, const alga_reads *dev_reads, uint32_t len_nt);
/* Seed index in pieces: inserts the reads [lo, hi); first != 0 clears the tables before (start of a build).  Lets the
 * caller overlap the arrival of the other ranks' reads with the insertion of those already there. */
int alga_ps_stage_index_range(alga_ps_plan *plan, uint32_t lo, uint32_t hi, int first, void *stream);

/* Result of the last run / stage_csr: device pointers (row_off has hi-lo+1 entries, relative to lo). */
int alga_ps_plan_result_device(alga_ps_plan *plan, const uint64_t **row_off, const int32_t **nbr,
                               const int32_t **off, uint64_t *n_edges);
/* Number of rows (hi - lo) of the last result; 0 before any run. */
uint32_t alga_ps_plan_result_rows(alga_ps_plan *plan);
/* Copy the result to freshly allocated host arrays. */
int alga_ps_plan_result_host(alga_ps_plan *plan, alga_csr *out);
/* Copy the result into the plan's own page-locked staging buffers (full PCIe/C2C rate, no allocation after the
 * first call); out->borrowed is set. */
int alga_ps_plan_result_host_pinned(alga_ps_plan *plan, alga_csr *out);
/* Counters of the last run. */
int alga_ps_plan_stats(alga_ps_plan *plan, alga_timing *timing);

/* ---- fingerprints (GraphCreatorPrefSuf::updatePrefixHash / updateSuffixHash, :213-236) ---- */
/* For every read i with len >= L: pre64/pre32 = fingerprints of its length-L prefix, suf64/suf32 of its
 * length-L suffix, h64 = sum s_j 4^j mod 10^18+3, h32 = sum s_j 4^j mod 10^9+7 (Params.cpp:721,
 * GraphCreatorPrefSuf.h:42).  Host buffers; entries of shorter reads are left untouched. */
int alga_gpu_fingerprints(const alga_reads *reads, int32_t L, int32_t device, uint64_t *pre64, uint32_t *pre32,
                          uint64_t *suf64, uint32_t *suf32);

/* ---- 2-bit packing (Read::createSequence, Read.cpp:40-68) ----------------------------------- */
/* ascii: n_reads sequences of fixed length len_nt stored back to back (no separators); words: n_reads *
 * ceil(len_nt/16) blocks.  Host buffers. */
int alga_gpu_pack_reads(const uint8_t *ascii, uint32_t n_reads, uint32_t len_nt, int32_t device, uint32_t *words);

/* ---- candidate verification (AlignmentControllerHybrid::canAlign, AlignmentControllerHybrid.cpp:46-83
 *      -> AlignmentControllerLowErrorRate::canAlign, AlignmentControllerLowErrorRate.cpp:15-49) ---- */
typedef struct {
    int32_t max_offset_pct;   /* Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT */
    int32_t min_offset;       /* Params::MIN_OFFSET_FOR_ALIGNMENT */
    int32_t min_overlap_area; /* Params::MIN_OVERLAP_AREA */
    int32_t threshold_pct;    /* Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR */
    int32_t same_ends;        /* Params::ALIGNMENT_CONTROLLER_SAME_ENDS_LENGTH */
    int32_t device;
    int32_t lcs_rate_pct;     /* 0: Params::USE_ACLER_INSTEAD_OF_ACLCS = 1, the reference's default (Params.cpp:703): the verdict of
                                 the low-error test stands.  > 0: USE_ACLER_INSTEAD_OF_ACLCS = 0 and this is
                                 Params::MINIMAL_OVERLAP_RATE_FOR_LCS: pairs the low-error test rejects go on to the banded LCS of
                                 AlignmentControllerLCS::canAlign (AlignmentControllerLCS.cpp:30-59, 61-150) */
    int32_t lcs_band;         /* Params::MAX_ERROR_RATE_FOR_LCS (2): |q - (p - offset)| <= lcs_band; at most 8 */
} alga_verify_params;

/* pairs: n_pairs x (a, b, offset) int32 triples; verdict[i] = canAlign(reads[a], reads[b], offset).
 * One candidate pair per warp.  Host buffers. */
int alga_gpu_verify_pairs(const alga_reads *reads, const int32_t *pairs, uint64_t n_pairs,
                          const alga_verify_params *params, uint8_t *verdict);

/* ---- error-rate supplement -------------------------------------------------------------------
 * Drop-in for main.cpp:300-355 (runs when --error_rate > 0.01): GraphCreatorLI / GraphCreatorPairwiseKmerBranch over the
 * dead-end reads of the graph produced by alga_gpu_prefsuf_build, four passes with rotated nucleotide priorities,
 * followed by Graph::retainOnlySmallestOffset (main.cpp:346).  LI k-mer extraction (Read.cpp:145-226) and all canAlign
 * calls (AlignmentControllerHybrid.cpp:46-83) run on the GPU, and so does the enumeration of the candidate pairs with the
 * static filters of GraphCreatorPairwiseKmerBranch.cpp:43-62.  The sort of the k-mers (GraphCreatorKmerBased.cpp:94-101)
 * is a stable radix sort on the GPU; only the hash-range buckets that hold k-mers tied in every key are put back into the
 * reference's fill order and sorted with libstdc++'s std::sort on the host, so ties fall exactly as in a reference built
 * with the same toolchain.  The order-dependent edge replay (GraphCreatorPairwiseKmerBranch.cpp:64-97) runs on the host. */
typedef struct {
    int32_t max_offset_pct;     /* Params::MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT = (1 - SCALE) * avg_len / 2 (main.cpp:335) */
    int32_t min_offset;         /* Params::MIN_OFFSET_FOR_ALIGNMENT */
    int32_t min_overlap_area;   /* Params::MIN_OVERLAP_AREA = (1 + SCALE) * avg_len / 2 (main.cpp:333) */
    int32_t threshold_pct;      /* Params::MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR = 99 - ERROR_RATE (main.cpp:336) */
    int32_t same_ends;          /* Params::ALIGNMENT_CONTROLLER_SAME_ENDS_LENGTH (3) */
    int32_t kmer_length;        /* Params::LI_KMER_LENGTH = 35 (main.cpp:340) */
    int32_t intervals;          /* Params::LI_KMER_INTERVALS = 6 (main.cpp:339) */
    int32_t kmer_length_bucket; /* Params::KMER_LENGTH_BUCKET (main.cpp:104): shorter reads contribute no k-mers */
    int32_t device;
} alga_sup_params;
/* graph_in: Graph::V after main.cpp:291 (host CSR); graph_out: Graph::V after main.cpp:346 (malloc'ed host CSR, release
 * with alga_gpu_free_csr).  timing (may be NULL): h2d_ms, device_ms (kernels + their transfers), total_ms,
 * kernel_launches; stage_ms[0..4] = LI k-mers, sort of the k-mers (device radix sort + host re-sort of the tied buckets),
 * pair enumeration, canAlign batch, ordered replay (host) (summed over the four passes); stage_ms[5] = dead-end reads that took part, stage_ms[6] = pairs verified, stage_ms[7] =
 * dependency levels of the replay (groups of k-mers that share no source read run in parallel, level after level). */
int alga_gpu_supplement(const alga_reads *reads, const alga_csr *graph_in, const alga_sup_params *params,
                        alga_csr *graph_out, alga_timing *timing);
/* Read::getLIKmers for the reads ids[0 .. n_ids) (host buffers): hash_out / ind_out have n_ids * intervals entries,
 * ind = -1 for intervals beyond the last window; priorities[c] = rank of nucleotide code c (Read::priorities). */
int alga_gpu_li_kmers(const alga_reads *reads, const uint32_t *ids, uint32_t n_ids, const int32_t priorities[4],
                      int32_t kmer_length, int32_t intervals, int32_t device, uint64_t *hash_out, int32_t *ind_out);

/* ---- read preprocessing (the step before the graph build) ----------------------------
 * ReadPreprocess::getPrefixReads (src/IO/ReadPreprocess.cpp:13-77, called from main.cpp:132-134): mask[i] = 1 for every
 * read the reference removes -- identical reads except the one with the greatest id and, with remove_type 2
 * (Params::PREF_READS_ALL_PREFIX_READS, the default), reads that are a proper prefix of another read together with
 * their reverse complements (id ^ 1); remove_type 1 = Params::PREF_READS_ONLY_DUPLICATES.  Host buffers; align flags
 * are ignored.  Reads of up to 65535 nucleotides. */
int alga_gpu_prefix_reads(const alga_reads *reads, int32_t remove_type, int32_t device, uint8_t *mask,
                          alga_timing *timing /* may be NULL */);

/* ---- reading the input (the producer of the read set) -------------------------------------
 * InputReader::readInput (src/IO/InputReader.cpp:44-139, called from main.cpp:82) on the contents of the input file(s):
 * record splitting (readOneRead1, :142-180), space stripping and end trimming (:286-303), the character check (:316-334),
 * removal of reads with an N (:341-347) and of reads whose minimal period is <= 20 (:343-353, MyUtils::MinPeriod), the
 * reverse-complement twin (:362-377), 2-bit packing (Read::createSequence, Read.cpp:40-68) and the final order of
 * Global::READS (:54-85): reverse complement at the even id, forward strand at the odd id, mates of the two files
 * interleaved (file-1 record i -> ids 4i, 4i+1; file-2 record i -> 4i+2, 4i+3).  Results follow the reference's
 * --threads=1 order.  Where the reference exits (a character other than A C G T N U) or indexes out of range (mate files
 * with different record counts) the call fails with ALGA_E_INVALID. */
#define ALGA_INPUT_PLAIN 0 /* Params::MY_INPUT: whitespace-separated sequences */
#define ALGA_INPUT_FASTA 1 /* Params::FASTA, and PFASTA with paired reads: header line + sequence line */
#define ALGA_INPUT_FASTQ 2 /* Params::FASTQ: four lines per record */
typedef struct {
    int32_t file_type;     /* ALGA_INPUT_*: Params::INPUT_FILE_TYPE (from the file extension, Params.cpp:332-335) */
    int32_t trim_left;     /* Params::READ_END_TRIM_LEFT (3) */
    int32_t trim_right;    /* Params::READ_END_TRIM_RIGHT (3) */
    int32_t rna;           /* Params::RNA: U is read as T */
    int32_t str_threshold; /* 20 (InputReader.cpp:343); <= 0 means 20 */
    int32_t device;
} alga_input_params;

/* A read set produced by the library: fixed stride, read i = words[i * stride_words ...], len_nt[i] == 0 = nullptr.
 * Host arrays owned by the library (page-locked staging, see `borrowed`); release with alga_gpu_free_read_set. */
typedef struct {
    uint32_t n_reads;
    uint32_t stride_words;  /* ceil(max_len_nt / 16), at least 1 */
    uint32_t max_len_nt;
    uint32_t *words;        /* n_reads * stride_words */
    uint32_t *len_nt;       /* n_reads */
    uint32_t *old_id;       /* alga_gpu_remap_reads: id of read i before the renumbering; NULL otherwise */
    uint8_t *paired_offset; /* alga_gpu_remap_reads: Global::pairedReadOffset[i] (main.cpp:176-197); NULL otherwise */
    uint64_t n_records[2];  /* alga_gpu_read_input: records read from file 1 / file 2 */
    uint64_t n_with_n;      /* ... records dropped because of an N ("Nreads", InputReader.cpp:346) */
    uint64_t n_str;         /* ... records dropped as short-period repeats ("STRreads", InputReader.cpp:352) */
    int32_t borrowed;       /* != 0: the arrays are page-locked staging buffers owned by the library, valid until the next
                               call of the function that produced them; alga_gpu_free_read_set then only clears them */
} alga_read_set;

/* text1 / text2: the bytes of --file1 / --file2 (text2 may be NULL: single-end).  Text in a buffer from
 * alga_gpu_host_alloc is uploaded at the full host->device rate; any other buffer goes through page-locked chunks.
 * timing (may be NULL): h2d_ms, device_ms, d2h_ms, total_ms, kernel_launches. */
int alga_gpu_read_input(const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2,
                        const alga_input_params *params, alga_read_set *out, alga_timing *timing);

/* The compaction and renumbering of main.cpp:150-232 after Global::removeRead of the reads ReadPreprocess marked
 * (main.cpp:133-140): strand pairs (2u, 2u+1) whose first read is present (len_nt != 0 and remove_mask == 0) keep their
 * order and get consecutive ids; out->old_id and out->paired_offset are filled.  remove_mask may be NULL.  A read that
 * survives without its reverse complement is the reference's assert (main.cpp:173): ALGA_E_INVALID. */
int alga_gpu_remap_reads(const alga_reads *reads, const uint8_t *remove_mask, int32_t device, alga_read_set *out,
                         alga_timing *timing);
void alga_gpu_free_read_set(alga_read_set *rs);

/* ---- from the input files to the overlap graph in one call ------------------------------------------------------
 * The reference driver between main.cpp:82 and main.cpp:291 with every stage on the GPU and nothing but scalars coming
 * back in between: InputReader::readInput (:82), the parameters derived from the average read length (:93-110),
 * ReadPreprocess::getPrefixReads + Global::removeRead (:132-140), the renumbering (:150-232), the removal of reads too
 * short for the graph creators (:253-266), GraphCreatorPrefSuf::startAlignmentGraphCreation and
 * Graph::retainOnlySmallestOffset (:249-291).  What the rest of the driver needs comes back: Global::READS with
 * Global::pairedReadOffset, and Global::GRAPH. */
typedef struct {
    alga_input_params input;
    int32_t remove_type;    /* Params::REMOVE_PREF_READS_TYPE: 0 none, 1 duplicates only, 2 all prefix reads (the default) */
    float scale;            /* Params::SCALE (0.55); <= 0 means 0.55 */
    int32_t min_overlap;    /* Params::MOST_FREQUENTLY_USED_PARAMETER (-l): > 0 overrides the derived minimum overlap */
    int32_t rs_min_overlap; /* > 0 overrides Params::REMOVE_SMALL_OVERLAP_EDGES_MIN_OVERLAP */
} alga_driver_params;
typedef struct {
    alga_read_set reads;    /* Global::READS at main.cpp:282 (old_id: id after the reader; paired_offset filled) */
    alga_csr graph;         /* Global::GRAPH after main.cpp:291 */
    double avg_len;         /* Global::calculateAvgReadLength() at main.cpp:93 */
    int32_t min_overlap, rs_min_overlap, li_kmer_length; /* the Params the driver derived (main.cpp:99-108) */
    uint32_t n_reads_in;    /* reads (both strands, nullptr included) the reader produced */
    uint64_t n_records[2], n_with_n, n_str; /* as in alga_read_set after alga_gpu_read_input */
} alga_overlap_graph;
/* timing (may be NULL): h2d_ms (file text), device_ms (text resident -> CSR resident), d2h_ms (graph + reads), total_ms,
 * stage_ms[0..4] = reader, prefix reads, renumbering, graph build on the device (CUDA events), its host side.
 * out->reads and out->graph are borrowed page-locked buffers (valid until the next call); release both with
 * alga_gpu_free_read_set / alga_gpu_free_csr. */
int alga_gpu_files_to_graph(const uint8_t *text1, uint64_t n1, const uint8_t *text2, uint64_t n2,
                            const alga_driver_params *params, alga_overlap_graph *out, alga_timing *timing);

/* ---- first step of the graph simplifier (the step after the graph build) -------------------
 * Graph::sortEdgesByIncreasingOffset (Graph.cpp:583-614) + GraphSimplifier::cutNonAndWeaklyMetricTriangles
 * (GraphSimplifier.cpp:228-349) as simplifyGraphOld runs them first (GraphSimplifier.cpp:110-130): the edge i -> b of
 * offset w <= max_offset (Params::MAX_OFFSET_PARALLEL_PATHS, main.cpp:95) is removed iff the shortest two-hop path
 * i -> a -> b has length exactly w; all decisions are taken on the input graph.  graph_in: host CSR with rows sorted by
 * neighbour (what alga_gpu_prefsuf_build / alga_gpu_supplement return).  graph_out: host CSR in borrowed page-locked
 * buffers (valid until the next call of this function; release with alga_gpu_free_csr), rows sorted by (offset, neighbour) -- the order sortEdgesByIncreasingOffset leaves; the reference's
 * swap-and-pop removal leaves the surviving entries of a row in another order, the edge set is the same. */
int alga_gpu_cut_triangles(const alga_csr *graph_in, int32_t max_offset, int32_t device, alga_csr *graph_out,
                           alga_timing *timing /* may be NULL */);

/* ---- misc ---------------------------------------------------------------------------------- */
/* Page-locked host memory for callers that stage the packed reads themselves (the shim gathers the blocks of
 * vector<Read*> straight into such a buffer, so the upload runs at full host->device rate).  NULL on failure. */
void *alga_gpu_host_alloc(size_t bytes);
void alga_gpu_host_free(void *p);
int alga_gpu_device_count(void);
const char *alga_gpu_last_error(void);
const char *alga_gpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ALGA_GPU_H */

// Host-callable launchers of the sm_100a kernels (implemented in the .cu files of this directory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace alga {

struct LaunchCfg {
    int sm_count = 148;
    uint64_t *launches = nullptr;  // incremented once per kernel launch
};

// --- read-set statistics: max length, eligible prefix/suffix counts ----------------------------
struct ReadStats {
    uint32_t max_len;
    uint32_t n_prefix;  // reads that can be the target of an edge (alignTo && len >= lmin)
    uint32_t n_suffix;  // reads that can be the source of an edge (alignFrom && len - min_offset >= lmin)
    uint32_t min_len;   // shortest read, removed reads (length 0) included
};
void launch_read_stats(const ReadsDev &R, int lmin, int min_offset, ReadStats *d_stats, cudaStream_t s,
                       const LaunchCfg &cfg);

// --- seed index -------------------------------------------------------------------------------
// inserts the seeds of the reads [lo, hi) whose bucket lies in [b_lo, b_hi) (the tables must have been cleared before
// the first range; prefix and suffix table have the same number of buckets when b_lo / b_hi restrict anything)
// which: 1 = prefix table only, 2 = suffix table only, 3 = both
void launch_build_index(const ReadsDev &R, const PsDev &P, SeedTable prefix, SeedTable suffix, uint32_t lo, uint32_t hi,
                        uint32_t b_lo, uint32_t b_hi, cudaStream_t s, const LaunchCfg &cfg, int which = 3);

// sharded build of equal-length reads: seed records (bucket prefix side, bucket suffix side, the two 16-bit tags: 12 bytes) of
// `n` reads in the caller's fixed-stride layout, and the inserts out of such records into the bucket range [b_lo, b_hi)
void launch_seed_keys(const uint32_t *words, uint32_t stride, uint32_t n, const PsDev &P, SeedTable prefix, SeedTable suffix,
                      uint32_t *keys, cudaStream_t s, const LaunchCfg &cfg);
void launch_index_keys(const uint32_t *keys, uint32_t first_id, uint32_t n, SeedTable prefix, SeedTable suffix, uint32_t b_lo,
                       uint32_t b_hi, cudaStream_t s, const LaunchCfg &cfg);

// --- sharded runs: read what the peers produced for this rank out of their exchange workspaces (NVLink) ----
// seg[p] / cnt[p]: peer p's segment for this rank and its entry count (device pointers valid in this process)
void launch_pull_rows(const void *const *seg, const uint32_t *const *cnt, int world, uint32_t cap, uint32_t n_expected,
                      const Phase1Out &out, cudaStream_t s, const LaunchCfg &cfg);
void launch_pull_triples(const void *const *seg, const uint32_t *const *cnt, int world, uint32_t cap, uint32_t n_expected,
                         uint64_t cap_local, int32_t *triples, unsigned long long *n_total, uint32_t *outdeg,
                         cudaStream_t s, const LaunchCfg &cfg);

// --- phase 1: L in [lmin, min(rs-1, max_l)], keeps the 3 largest (L, c) per source read ----------
// Edges go where `out` says (common.cuh: rows of the transposed graph, or an edge list), each with its overhang tail
// (the last min(offset, 32) nucleotides of b[0 .. offset), top-aligned).
// Thread-per-read fast kernel (tpr_kernels.cu); reads it cannot take are appended to hard_queue (*n_hard must be 0).
// id_list == nullptr: first pass over the reads [lo, hi) (up to two tag matches per window).  Otherwise second pass
// over id_list[0 .. *n_list) -- the queue the first pass left -- with up to four matches per window (reads that start
// at the same position and differ by sequencing errors share their seed); what it gives up on goes to its own queue.
void launch_phase1_tpr(const ReadsDev &R, const SeedTable &prefix, const PsDev &P, uint32_t max_len_nt, uint32_t lo,
                       uint32_t hi, const uint32_t *id_list, const uint32_t *n_list, const Phase1Out &out,
                       uint32_t *hard_queue, uint32_t *n_hard, int force_hard, cudaStream_t s, const LaunchCfg &cfg);
// generic kernel over the queue (n_max = upper bound of the queue length, used for the grid only).  hard_queue2
// (nullable): if the first queue was long enough for the second fast pass (kSecondPassMin), the queue IT left.
void launch_phase1_queue(const ReadsDev &R, const SeedTable &prefix, const PsDev &P, uint32_t n_max,
                         const uint32_t *hard_queue, const uint32_t *n_hard, const uint32_t *hard_queue2,
                         const uint32_t *n_hard2, const Phase1Out &out, cudaStream_t s, const LaunchCfg &cfg);
// edge list -> (b, c, o) triples; the list length is read on the device, n_max bounds the grid
void launch_edges_to_triples(const Edge1 *list, const uint32_t *n_list, uint64_t n_max, int32_t *triples, cudaStream_t s,
                             const LaunchCfg &cfg);
// only if *n_over > kOverScanMax (tested on the device): CSR form of the transposed graph from fixed rows + overflow list
void launch_rebuild_rows_csr(const uint32_t *n_over, uint32_t over_cap, const Edge1 *over, uint32_t *indeg,
                             const RevEntry *rows, uint32_t cap, uint32_t n_targets, uint32_t *rev_off, RevEntry *rev,
                             uint64_t rev_cap, uint32_t *rev_overflow, void *scan_ws, cudaStream_t s,
                             const LaunchCfg &cfg);

// --- reversed phase-1 adjacency from exchanged triples (rows by target c in [lo,hi)) ---------------
// counts -> offsets is done with launch_scan; scatter consumes `cursor` (a copy of the counts).
void launch_count_targets(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, uint32_t *indeg,
                          cudaStream_t s, const LaunchCfg &cfg);
void launch_scatter_rev_triples(const ReadsDev &R, const int32_t *triples, uint64_t n, uint32_t c_lo, uint32_t c_hi,
                                const uint32_t *rev_off, uint32_t *cursor, RevEntry *rev, cudaStream_t s,
                                const LaunchCfg &cfg);

// --- phase 2: L in [max(rs,lmin), max_l] with per-target transitive reduction ----------------------
struct Phase2Out {
    int32_t *triples;             // (b, c, offset)
    unsigned long long *n_edges;  // device counter (may exceed edge_cap -> caller retries)
    uint64_t edge_cap;
    uint32_t *outdeg;             // nullable, indexed by b (global id)
    uint32_t *spill_queue;        // targets whose list outgrew the on-chip capacity
    uint32_t *n_spill;
    ShardOut sh;                  // sharded runs (sh.world > 1): triples go to the segment of the owner of b instead
};
// one surviving edge, one thread (generic kernels)
__device__ __forceinline__ void emit_triple_sharded(const ShardOut &sh, int32_t a, int32_t c, int32_t o) {
    const uint32_t d = shard_of(sh, (uint32_t) a);
    const uint32_t i = atomicAdd(sh.cnt + d, 1u);
    if (i < sh.cap) {
        int32_t *t = reinterpret_cast<int32_t *>(sh.seg) + ((uint64_t) d * sh.cap + i) * 3;
        t[0] = a, t[1] = c, t[2] = o;
    }
}
// thread-per-target fast kernel (tpr_kernels.cu); everything it cannot take goes to out.spill_queue
void launch_phase2_tpr(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t max_len_nt, uint32_t lo,
                       uint32_t hi, const uint32_t *id_list, const uint32_t *n_list, const RowsView &rows,
                       const Phase2Out &out, int force_hard, cudaStream_t s, const LaunchCfg &cfg);
// generic path: sequential replay on a shared-memory list of list_cap entries per target (spills beyond it)
void launch_phase2(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t lo, uint32_t hi,
                   const RowsView &rows, int list_cap, const Phase2Out &out, cudaStream_t s, const LaunchCfg &cfg);
// spill path: per queued target count row size + hits -> caps (u32), then replay with global lists
void launch_phase2_count(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t lo, const RowsView &rows,
                         const uint32_t *queue, uint32_t n_queue, uint32_t *caps, cudaStream_t s,
                         const LaunchCfg &cfg);
void launch_phase2_spill(const ReadsDev &R, const SeedTable &suffix, const PsDev &P, uint32_t lo, const RowsView &rows,
                         const uint32_t *queue, uint32_t n_queue, const uint64_t *spill_off, uint32_t *spill_store,
                         const Phase2Out &out, cudaStream_t s, const LaunchCfg &cfg);

// --- the scatter-shaped stages done by sorting (sorted_stages.cu): seed records of both tables for the reads [0, n) (one pass, R in
// any layout), then one table out of its records into a cleared table (the two tables may be built on two streams); CSR rows
// [lo, hi) (+ row_off, unsorted inside a row) out of n triples.  n < 2^31 each; `ws` = *_workspace_bytes() of scratch per table.
// Return 0 or a cudaError_t of the sort.
size_t sorted_index_workspace_bytes(uint32_t n_reads);
void launch_seed_records(const ReadsDev &R, const PsDev &P, const SeedTable &Tp, const SeedTable &Ts, uint32_t n, void *ws_prefix,
                         void *ws_suffix, cudaStream_t s, const LaunchCfg &cfg);
int launch_sorted_index(const SeedTable &T, uint32_t n, void *ws, cudaStream_t s, const LaunchCfg &cfg);
size_t sorted_csr_workspace_bytes(uint64_t n_edges);
int launch_sorted_csr(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, void *ws, uint64_t *row_off,
                      int32_t *nbr, int32_t *off, cudaStream_t s, const LaunchCfg &cfg);

// --- CSR assembly -----------------------------------------------------------------------------
void launch_count_sources(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, uint32_t *outdeg,
                          cudaStream_t s, const LaunchCfg &cfg);
void launch_scatter_csr(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap,
                        const uint64_t *row_off, uint32_t *cursor, int32_t *nbr, int32_t *off, cudaStream_t s,
                        const LaunchCfg &cfg);
// the same with fewer random accesses (n < 2^32): `cursor` (row sizes on entry) and `pairs` (n x 8 bytes) are scratch
void launch_scatter_csr_pairs(const int32_t *triples, uint64_t n, uint32_t lo, uint32_t hi, int swap, const uint64_t *row_off,
                              uint32_t *cursor, void *pairs, int32_t *nbr, int32_t *off, cudaStream_t s, const LaunchCfg &cfg);
// sort every row by (nbr, off); rows longer than 32 go through `big_rows` (queue of row ids) and tmp buffers
void launch_sort_rows(const uint64_t *row_off, uint32_t n_rows, int32_t *nbr, int32_t *off, uint32_t *big_rows,
                      uint32_t *n_big, cudaStream_t s, const LaunchCfg &cfg);
void launch_sort_big_rows(const uint64_t *row_off, const uint32_t *big_rows, uint32_t n_big, int32_t *nbr,
                          int32_t *off, int32_t *tmp_nbr, int32_t *tmp_off, cudaStream_t s, const LaunchCfg &cfg);

// --- exclusive scan of u32 counts into u32 / u64 offsets (n+1 outputs, out[n] = total) ---------------
size_t scan_workspace_bytes(uint64_t n);
void launch_scan_u32(const uint32_t *in, uint32_t *out, uint64_t n, void *workspace, cudaStream_t s,
                     const LaunchCfg &cfg, const uint32_t *run_if = nullptr);  // run_if: device counter, <= kOverScanMax = skip
void launch_scan_u64(const uint32_t *in, uint64_t *out, uint64_t n, void *workspace, cudaStream_t s,
                     const LaunchCfg &cfg);

// --- misc kernels -----------------------------------------------------------------------------
void launch_fill_u64(uint64_t *p, uint64_t v, uint64_t n, cudaStream_t s, const LaunchCfg &cfg);
void launch_pack_reads(const uint8_t *ascii, uint32_t n_reads, uint32_t len_nt, uint32_t *words, cudaStream_t s,
                       const LaunchCfg &cfg);
// copy of a fixed-stride read set at another stride (`words` words per read are copied, the rest of the slot is zero)
void launch_repack_reads(const uint32_t *in, uint32_t stride_in, uint32_t words, uint64_t n_reads, uint32_t *out,
                         uint32_t stride_out, cudaStream_t s, const LaunchCfg &cfg);
void launch_fingerprints(const ReadsDev &R, int L, uint64_t *pre64, uint32_t *pre32, uint64_t *suf64,
                         uint32_t *suf32, cudaStream_t s, const LaunchCfg &cfg);
struct VerifyDev {
    int32_t max_offset_pct, min_offset, min_overlap_area, threshold_pct, same_ends;
    int32_t lcs_rate_pct, lcs_band;  // lcs_rate_pct > 0: pairs the low-error test rejects go on to the banded LCS
};
void launch_verify_pairs(const ReadsDev &R, const int32_t *pairs, uint64_t n_pairs, const VerifyDev &V,
                         uint8_t *verdict, cudaStream_t s, const LaunchCfg &cfg);

// --- ReadPreprocess::getPrefixReads (preprocess.cu): mask[i] = 1 for duplicates / prefix reads (+ reverse complements)
// t: an empty (zeroed) table sized for R.n entries; lenmap: prefix_reads_lenmap_words() words; flags: R.n words.
// lenmap[last] != 0 afterwards: a read is longer than the 65 535 nucleotides the length map covers.
void launch_prefix_reads(const ReadsDev &R, const SeedTable &t, int remove_type, uint32_t *lenmap, uint32_t *flags,
                         uint8_t *mask, cudaStream_t s, const LaunchCfg &cfg);
size_t prefix_reads_lenmap_words();

// --- InputReader::readInput and the renumbering of main.cpp:150-232 (input.cu) -------------------------------------
// text: the file as it is, in a buffer padded to a multiple of 16 bytes.  Marks = line ends (plain input: token starts);
// mark k with k % lines_per_record == 0 precedes the sequence line of record k / lines_per_record.
uint64_t input_mark_blocks(uint64_t n_bytes);
size_t input_rec_info_bytes();
size_t input_scalars_bytes();
void launch_count_marks(const uint8_t *text, uint64_t n, bool plain, uint32_t *block_cnt, cudaStream_t s, const LaunchCfg &cfg);
void launch_write_marks(const uint8_t *text, uint64_t n, bool plain, const uint64_t *block_off, uint32_t lines_per_record,
                        uint64_t *rec_start, uint64_t *rec_end, uint64_t n_cand, cudaStream_t s, const LaunchCfg &cfg);
// scalars (input_scalars_bytes()): u32 first_empty = first_bad = 0xFFFFFFFF, u32 max_len = 0, u32 pad, u64 n_with_n = n_str = sum_len = n_alive = 0
void launch_scan_records(const uint8_t *text, uint64_t n, bool plain, const uint64_t *rec_start, const uint64_t *rec_end,
                         uint64_t n_marks, uint32_t lines_per_record, uint32_t n_cand, int trim_left, int trim_right, int rna,
                         int str_threshold, void *info, void *scalars, cudaStream_t s, const LaunchCfg &cfg);
void launch_record_totals(const void *info, uint32_t n_rec, void *scalars, cudaStream_t s, const LaunchCfg &cfg);
void launch_pack_records(const uint8_t *text, const void *info, uint32_t n_rec, int rna, uint32_t id_step, uint32_t id_first,
                         uint32_t stride, uint32_t *words, uint32_t *len_out, cudaStream_t s, const LaunchCfg &cfg);
// a few bytes device -> page-locked host memory by a kernel (control read-backs that must not queue behind a bulk DMA)
void launch_peek(void *dst_host, const void *src_dev, uint32_t bytes, cudaStream_t s, const LaunchCfg &cfg);
// bulk copy device -> page-locked host memory by a kernel (not the DMA engine); 16-byte aligned pointers
void launch_copy_to_host(void *dst_host, const void *src_dev, uint64_t bytes, cudaStream_t s, const LaunchCfg &cfg);
// scalars: u32 max_len = 0, u32 err = 0 (1 + id of a read that survives without its reverse complement)
void launch_remap_flags(const ReadsDev &R, const uint8_t *mask, uint32_t n_units, uint32_t *flag, void *scalars, cudaStream_t s,
                        const LaunchCfg &cfg);
void launch_remap_scatter(const ReadsDev &R, uint32_t n_units, const uint32_t *flag, const uint32_t *pos, uint32_t stride,
                          uint32_t *words, uint32_t *len_out, uint32_t *old_id, uint8_t *paired_offset, uint32_t min_keep_len,
                          cudaStream_t s, const LaunchCfg &cfg);

// --- first simplifier step (simplify.cu): GraphSimplifier::cutNonAndWeaklyMetricTriangles on a CSR with rows sorted by neighbour
// keep[e] = 0 for removed entries, kept[i] = surviving entries of row i
void launch_triangle_marks(const uint64_t *row_off, const int32_t *nbr, const int32_t *off, uint32_t n, int32_t max_offset,
                           uint8_t *keep, uint32_t *kept, cudaStream_t s, const LaunchCfg &cfg);
void launch_triangle_compact(const uint64_t *row_off, const int32_t *nbr, const int32_t *off, const uint8_t *keep, uint32_t n,
                             const uint64_t *new_off, int32_t *out_nbr, int32_t *out_off, cudaStream_t s, const LaunchCfg &cfg);

// --- error-rate supplement (supplement.cu) -------------------------------------------------------
// LI k-mers (Read.cpp:145-226) of the reads d_ids[0 .. n_ids): `intervals` slots per read, ind = -1 where absent
int run_li_kmers(const ReadsDev &R, const uint32_t *d_ids, uint32_t n_ids, const int32_t prio[4], int K, int intervals,
                 uint64_t *d_hash, int32_t *d_ind, cudaStream_t s, const LaunchCfg &cfg);
const char *supplement_last_error();

}  // namespace alga

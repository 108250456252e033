"""Seeded parity cases shared by the CPU (oracle vs golden) and GPU (CUDA vs oracle) tests."""
from __future__ import annotations

import numpy as np

from alga_b200 import readset, synth


def _periodic(seed, n_reads, dedupe):
    rng = np.random.default_rng(seed)
    unit = rng.integers(0, 4, size=37, dtype=np.uint8)
    g = np.tile(unit, 400)
    mut = rng.random(g.shape) < 0.02
    g = np.where(mut, (g + 1) & 3, g).astype(np.uint8)
    reads = []
    for _ in range(n_reads):
        ln = int(rng.integers(80, 121))
        s = int(rng.integers(0, len(g) - ln))
        x = g[s:s + ln].copy()
        reads.append(readset.revcomp_codes(x))
        reads.append(x)
    if dedupe:
        reads = synth.remove_prefix_reads_general(reads)
    return readset.from_code_list(reads)


def _with_flags(rs, seed, p=0.7):
    rng = np.random.default_rng(seed)
    rs.align_from[:] = rng.random(rs.n) < p
    rs.align_to[:] = rng.random(rs.n) < p
    return rs


def _with_nulls(rs, seed, p=0.1):
    rng = np.random.default_rng(seed)
    ln = rs.len_nt.copy()
    ln[rng.random(rs.n) < p] = 0
    return readset.ReadSet(rs.words, rs.word_off, ln)


def build_case(name):
    """-> (ReadSet, min_overlap, rs_min_overlap, min_offset)"""
    if name == "cfg1_small":          # BASELINE config 1 shape: 100 bp SE 30x, error-free
        w = synth.make_config("cfg1", scale=0.03)
        return w.reads, w.params.min_overlap, w.params.rs_min_overlap, 0
    if name == "cfg2_small":          # config 2 shape: 2x150 bp 50x, error-free
        w = synth.make_config("cfg2", scale=0.008)
        return w.reads, w.params.min_overlap, w.params.rs_min_overlap, 0
    if name == "cfg3_small":          # config 3 shape: 1 % substitutions
        w = synth.make_config("cfg3", scale=0.008)
        return w.reads, w.params.min_overlap, w.params.rs_min_overlap, 0
    if name == "cfg5_small":          # config 5 shape: 2x100 bp 40x with repeats
        w = synth.make_workload("cfg5s", 60_000, 100, 40, True, 5, repeats=6)
        return w.reads, w.params.min_overlap, w.params.rs_min_overlap, 0
    if name == "varlen":              # ragged lengths, errors, repeat, deduped
        return synth.make_variable_length(30000, 3000, 90, 150, seed=11, error=0.005, repeats=3), 55, 77, 0
    if name == "varlen_dups":         # duplicates and contained reads left in (order-dependent replay)
        return synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False), 40, 60, 0
    if name == "rs_eq_lmin":          # contig-trimming call: min overlap == rs == 25 (main.cpp:651-655)
        return synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False), 25, 25, 0
    if name == "rs_below_lmin":
        return synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False), 40, 10, 0
    if name == "rs_eq_maxl_plus1":    # everything is phase 1, transposed at the last no-op iteration
        return synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False), 40, 151, 0
    if name == "rs_above_maxl":       # reference leaves phase-1 edges reversed (SURVEY A.1 note 2)
        return synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False), 40, 400, 0
    if name == "min_offset":
        return synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False), 40, 60, 5
    if name == "flags":               # alignFrom / alignTo partially cleared
        rs = synth.make_variable_length(20000, 3000, 60, 150, seed=12, repeats=3, dedupe=False)
        return _with_flags(rs, 5), 40, 60, 0
    if name == "nulls":               # removed (nullptr) reads
        w = synth.make_config("cfg1", scale=0.03)
        return _with_nulls(w.reads, 5), 55, 77, 0
    if name == "long_reads":          # contig-like, longer than the 500 cap
        return synth.make_variable_length(50000, 300, 300, 3000, seed=13, repeats=2, dedupe=False), 25, 25, 0
    if name == "long_reads_rs":
        return synth.make_variable_length(50000, 300, 300, 3000, seed=13, repeats=2, dedupe=False), 100, 300, 0
    if name == "periodic_dups":       # low-complexity: many sources per (target, L), long in-neighbour lists
        return _periodic(7, 1500, dedupe=False), 45, 65, 0
    if name == "periodic":
        return _periodic(7, 1500, dedupe=True), 45, 65, 0
    if name == "short_lmin":          # seed shorter than 32 nt
        return synth.make_variable_length(8000, 1500, 30, 80, seed=21, repeats=2, dedupe=False), 12, 20, 0
    if name == "tiny":
        return synth.make_variable_length(400, 6, 60, 90, seed=3, dedupe=False), 20, 30, 0
    if name == "empty":
        return readset.from_code_list([]), 20, 30, 0
    if name == "all_null":
        return readset.from_code_list([None, None, None, None]), 20, 30, 0
    raise KeyError(name)


CASES = ["cfg1_small", "cfg2_small", "cfg3_small", "cfg5_small", "varlen", "varlen_dups", "rs_eq_lmin",
         "rs_below_lmin", "rs_eq_maxl_plus1", "rs_above_maxl", "min_offset", "flags", "nulls", "long_reads",
         "long_reads_rs", "periodic_dups", "periodic", "short_lmin", "tiny", "empty", "all_null"]


def verify_case(seed=31, n_reads=4000, genome=20000, read_len=144, error=0.01):
    """Candidate pairs for AlignmentControllerHybrid::canAlign: overlapping same-strand reads with 1 %
    substitutions (true offsets, offsets off by one, random pairs), supplement parameters of config 3
    (main.cpp:332-340: threshold 97 %, max offset 32 % of the read, min overlap area 111)."""
    rng = np.random.default_rng(seed)
    g = rng.integers(0, 4, size=genome, dtype=np.uint8)
    lens = np.where(rng.random(n_reads) < 0.8, read_len, rng.integers(120, read_len + 1, size=n_reads))
    pos = rng.integers(0, genome - read_len, size=n_reads)
    order = np.argsort(pos, kind="stable")
    pos, lens = pos[order], lens[order]
    reads = []
    for p, ln in zip(pos, lens):
        r = g[p:p + ln].copy()
        hit = rng.random(ln) < error
        r = np.where(hit, (r + rng.integers(1, 4, size=ln)) & 3, r).astype(np.uint8)
        reads.append(r)
    rs = readset.from_code_list(reads)
    pairs = []
    for i in range(n_reads):
        for j in range(i + 1, min(i + 12, n_reads)):
            off = int(pos[j] - pos[i])
            if off > 60:
                break
            pairs.append((i, j, off))
            if rng.random() < 0.2:
                pairs.append((i, j, off + int(rng.integers(-1, 2))))
            if rng.random() < 0.1:
                pairs.append((j, i, off))
    for _ in range(2000):
        pairs.append((int(rng.integers(0, n_reads)), int(rng.integers(0, n_reads)), int(rng.integers(0, 50))))
    pairs = np.array([p for p in pairs if p[2] >= 0], dtype=np.int32)
    vp = dict(threshold_pct=97, max_offset_pct=32, min_overlap_area=111, min_offset=0)
    return rs, pairs, vp


def supplement_params(avg_len: float, error_rate_pct: int = 2, scale: float = 0.55):
    """Params of the supplement as the reference driver derives them (main.cpp:93-115, 332-340), float arithmetic as there."""
    LEN = int(avg_len) + 6
    L = int(np.float32(LEN) * np.float32(scale))
    return dict(threshold_pct=99 - error_rate_pct,
                max_offset_pct=int((np.float32(1.0) - np.float32(scale)) * np.float32(avg_len) / 2),
                min_overlap_area=int((np.float32(1.0) + np.float32(scale)) * np.float32(avg_len) / 2),
                kmer_length_bucket=min(2 * L // 3, 60))


SUPPLEMENT_CASES = ["sup_cfg3", "sup_varlen"]


def supplement_case(name):
    """-> (ReadSet, min_overlap, rs_min_overlap, supplement params): the graph to supplement is the GraphCreatorPrefSuf
    result on the same reads (main.cpp:282-291)."""
    if name == "sup_cfg3":      # BASELINE config 3 shape: 2x150 bp, 1 % substitutions, --error_rate=0.02
        w = synth.make_config("cfg3", scale=0.01)
        return w.reads, w.params.min_overlap, w.params.rs_min_overlap, supplement_params(float(w.reads.len_nt.mean()))
    if name == "sup_varlen":    # ragged lengths 90..150 (some shorter than KMER_LENGTH_BUCKET never occur here), errors, repeats
        rs = synth.make_variable_length(30000, 3000, 90, 150, seed=17, error=0.01, repeats=3)
        return rs, 66, 94, supplement_params(float(rs.len_nt[rs.len_nt > 0].mean()))
    raise KeyError(name)


PREPROCESS_CASES = ["pre_equal", "pre_varlen", "pre_periodic"]


def preprocess_case(name):
    """Read sets BEFORE ReadPreprocess (main.cpp:132-232): duplicates and contained reads still in."""
    if name == "pre_equal":     # equal lengths, 2x150 bp at 50x over 30 kbp: many exact duplicates
        rng = np.random.default_rng(41)
        g = synth.make_genome(30_000, rng)
        m1, m2 = synth.sample_paired_end(g, 150, 50, rng, 0.0)
        return readset.from_code_matrix(synth.strand_nodes(m1, m2))
    if name == "pre_varlen":    # ragged lengths: proper prefixes, their reverse complements, duplicates
        return synth.make_variable_length(15000, 4000, 40, 150, seed=42, repeats=2, dedupe=False)
    if name == "pre_periodic":  # low complexity: long runs of reads that are prefixes of each other
        return _periodic(43, 1500, dedupe=False)
    raise KeyError(name)

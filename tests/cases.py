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

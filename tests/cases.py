"""Seeded parity cases shared by the CPU (oracle vs golden) and GPU (CUDA vs oracle) tests."""
from __future__ import annotations

import numpy as np

from alga_b200 import readset, synth
from alga_b200.graph_creator import supplement_params  # noqa: F401  (re-exported for the tests)


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


# (MINIMAL_OVERLAP_RATE_FOR_LCS, MAX_ERROR_RATE_FOR_LCS) settings of the banded-LCS fixtures (the reference's defaults: 95, 2)
LCS_SETTINGS = [(95, 2), (90, 1), (99, 3), (95, 0)]


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


# ---- input stage (InputReader::readInput, main.cpp:82; renumbering, main.cpp:150-232) -------------------------------
_NT = np.frombuffer(b"ACGT", np.uint8)


def _seq(codes) -> bytes:
    return _NT[np.asarray(codes, dtype=np.uint8)].tobytes()


def _fasta(seqs, prefix=b"r") -> bytes:
    return b"".join(b">" + prefix + str(i).encode() + b"\n" + s + b"\n" for i, s in enumerate(seqs))


def _fastq(seqs, rng) -> bytes:
    out = []
    for i, s in enumerate(seqs):
        q = bytes(rng.integers(33, 74, size=len(s), dtype=np.uint8).tolist())  # quality lines may hold '>', '@', ' ' ...
        out.append(b"@q" + str(i).encode() + b" 1:N:0\n" + s + b"\n+\n" + q + b"\n")
    return b"".join(out)


def _spice(seqs, rng, p_n=0.03, p_str=0.03, p_short=0.03, p_space=0.03):
    """Sprinkle the special cases of InputReader.cpp:286-353 over a list of sequences."""
    out = []
    for s in seqs:
        u = rng.random()
        if u < p_n:                      # an N somewhere -> both strands dropped
            k = int(rng.integers(0, len(s)))
            s = s[:k] + b"N" + s[k + 1:]
        elif u < p_n + p_str:            # short-period repeat (period 1..20, maybe imperfect at the ends that get trimmed)
            per = int(rng.integers(1, 24))
            unit = _seq(rng.integers(0, 4, size=per))
            s = (unit * (len(s) // per + 2))[: len(s)]
            if rng.random() < 0.3:
                s = b"ACG"[: min(3, len(s))] + s[3:]
        elif u < p_n + p_str + p_short:  # shorter than trim + 10 (not trimmed), down to 1 nt
            s = s[: int(rng.integers(1, 22))]
        elif u < p_n + p_str + p_short + p_space:  # leading spaces / trailing junk after a space
            s = b"  " + s + (b" trailing" if rng.random() < 0.5 else b"")
        out.append(s)
    return out


INPUT_CASES = ["in_fasta_se", "in_fasta_pe", "in_fastq_pe", "in_plain", "in_rna", "in_u_dna", "in_truncated", "in_no_eol",
               "in_empty", "in_long"]


def input_case(name):
    """-> (text1, text2 | None, file_type, extra harness args): file contents for InputReader::readInput."""
    from oracle.oracle import INPUT_FASTA, INPUT_FASTQ, INPUT_PLAIN
    rng = np.random.default_rng(sum(name.encode()) * 7919)
    g = synth.make_genome(20_000, rng)
    if name == "in_fasta_se":
        m = synth.sample_single_end(g, 100, 12, rng, 0.0)
        return _fasta(_spice([_seq(r) for r in m], rng)), None, INPUT_FASTA, []
    if name in ("in_fasta_pe", "in_fastq_pe"):
        m1, m2 = synth.sample_paired_end(g, 150, 14, rng, 0.005)
        s1 = _spice([_seq(r[: int(rng.integers(60, 151))]) for r in m1], rng)
        s2 = _spice([_seq(r[: int(rng.integers(60, 151))]) for r in m2], rng)
        if name == "in_fasta_pe":
            return _fasta(s1), _fasta(s2, b"m"), INPUT_FASTA, []
        return _fastq(s1, rng), _fastq(s2, rng), INPUT_FASTQ, []
    if name == "in_plain":          # whitespace-separated sequences, mixed separators
        m = synth.sample_single_end(g, 80, 6, rng, 0.0)
        seqs = _spice([_seq(r) for r in m], rng, p_space=0.0)
        seps = [b"\n", b" ", b"\t", b"\n\n", b" \n"]
        return b"".join(s + seps[int(rng.integers(0, len(seps)))] for s in seqs), None, INPUT_PLAIN, []
    if name in ("in_rna", "in_u_dna"):  # U characters: T with --rna=1, otherwise kept (packs as 0, complement unchanged)
        m = synth.sample_single_end(g, 100, 4, rng, 0.0)
        seqs = [_seq(r).replace(b"T", b"U") if rng.random() < 0.5 else _seq(r) for r in m]
        return _fasta(_spice(seqs, rng)), None, INPUT_FASTA, (["--rna=1"] if name == "in_rna" else [])
    if name == "in_truncated":      # an empty sequence line in the middle: reading stops there (InputReader.cpp:284)
        m = synth.sample_single_end(g, 100, 2, rng, 0.0)
        seqs = [_seq(r) for r in m]
        seqs[len(seqs) // 2] = b""
        return _fasta(seqs), None, INPUT_FASTA, []
    if name == "in_no_eol":         # last line without a newline; header-only tail in the mate file
        m1, m2 = synth.sample_paired_end(g, 100, 2, rng, 0.0)
        return _fasta([_seq(r) for r in m1])[:-1], _fasta([_seq(r) for r in m2]) + b">dangling header", INPUT_FASTA, []
    if name == "in_empty":
        return b"", None, INPUT_FASTA, []
    if name == "in_long":           # reads of several hundred nucleotides to a few thousand (more than one 32-lane step per line)
        seqs = []
        for _ in range(300):
            ln = int(rng.integers(200, 3000))
            s0 = int(rng.integers(0, len(g) - ln))
            seqs.append(_seq(g[s0:s0 + ln]))
        return _fasta(_spice(seqs, rng)), None, INPUT_FASTA, []
    raise KeyError(name)


FRONT_CASES = ["front_se", "front_pe", "front_short"]


def front_case(name):
    """-> (text1, text2 | None, file_type): inputs for the whole path main.cpp:82-291 (reader, duplicate / prefix-read
    removal, renumbering, GraphCreatorPrefSuf), pinned by the graph the stock binary serialises."""
    from oracle.oracle import INPUT_FASTA
    rng = np.random.default_rng(sum(name.encode()) * 104729)
    if name == "front_se":          # BASELINE config 1 shape: 100 bp single-end, 30x, error-free
        g = synth.make_genome(30_000, rng)
        m = synth.sample_single_end(g, 100, 30, rng, 0.0)
        return _fasta(_spice([_seq(r) for r in m], rng, p_short=0.0, p_space=0.0)), None, INPUT_FASTA
    if name == "front_pe":          # 2 x 150 bp, 40x, ragged lengths (contained reads), a repeat, a few errors
        g = synth.make_genome(25_000, rng, repeats=2, repeat_len=(300, 800))
        m1, m2 = synth.sample_paired_end(g, 150, 40, rng, 0.002)
        cut = lambda r: _seq(r[: int(rng.integers(110, 151))] if rng.random() < 0.3 else r)
        s1 = _spice([cut(r) for r in m1], rng, p_short=0.0, p_space=0.0)
        s2 = _spice([cut(r) for r in m2], rng, p_short=0.0, p_space=0.0)
        return _fasta(s1), _fasta(s2, b"m"), INPUT_FASTA
    if name == "front_short":       # a few reads shorter than LI_KMER_INTERVALS + LI_KMER_LENGTH: removed at main.cpp:253-266
        g = synth.make_genome(12_000, rng)
        m = synth.sample_single_end(g, 100, 25, rng, 0.0)
        seqs = [_seq(r) for r in m]
        for k in range(0, len(seqs), 23):
            seqs[k] = seqs[k][: int(rng.integers(24, 44))]
        return _fasta(seqs), None, INPUT_FASTA
    raise KeyError(name)


# ---- first simplifier step (cutNonAndWeaklyMetricTriangles) on graphs the reference itself produced ------------------
# (fixture the input graph comes from, key of its edge array, Params::MAX_OFFSET_PARALLEL_PATHS)
TRIANGLE_CASES = {"tri_cfg1": ("cfg1_small", "edges", 175), "tri_cfg3": ("cfg3_small", "edges", 262),
                  "tri_sup_cfg3": ("sup_cfg3", "after", 262), "tri_varlen_dups": ("varlen_dups", "edges", 250),
                  "tri_periodic": ("periodic_dups", "edges", 250), "tri_periodic_tight": ("periodic_dups", "edges", 20),
                  "tri_contigs": ("rs_eq_lmin", "edges", 250), "tri_empty": ("empty", "edges", 250)}


def triangle_case(name, golden_dir):
    """-> (edges_in (E, 3) sorted by (src, dst), n_nodes, max_offset)"""
    import os
    src, key, mx = TRIANGLE_CASES[name]
    e = np.load(os.path.join(golden_dir, f"{src}.npz"))[key]
    if src.startswith("sup_"):
        n = supplement_case(src)[0].n
    else:
        n = build_case(src)[0].n
    return e, n, mx

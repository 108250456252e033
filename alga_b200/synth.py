"""Seeded synthetic read sets and the host-side preprocessing that precedes the hot path.

The generators follow SURVEY.md §8(d): iid uniform genome, uniform start and strand,
paired fragments N(500, 50), iid substitutions, optional copied repeat segments.

``strand_nodes`` / ``remove_prefix_reads`` restate what the reference does to a read
before the graph creator sees it:
  * trim 3 nt at both ends                       (InputReader.cpp:298-303, Params.cpp:729-730)
  * ids: record k of file 1 -> (4k, 4k+1) = (revcomp, forward), its mate -> (4k+2, 4k+3);
    single-end (2k, 2k+1)                         (InputReader.cpp:53-80)
  * drop a read whose successor in sorted order starts with it (duplicates keep the
    greatest id; a proper prefix also drops its reverse-complement twin)
                                                 (ReadPreprocess.cpp:13-77)
  * compact and renumber in order                (main.cpp:150-232)
  * derive min overlap / small-overlap threshold (main.cpp:93-110)
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .readset import ReadSet, from_code_list, from_code_matrix, pack_matrix, revcomp_codes

TRIM = 3  # Params.cpp:729-730
SCALE = 0.55  # Params.cpp:678


@dataclass
class PrefSufParams:
    min_overlap: int
    rs_min_overlap: int
    min_offset: int = 0  # Params.cpp:709
    max_len_cap: int = 500  # GraphCreatorPrefSuf.cpp:92


def derive_params(avg_trimmed_len: float) -> PrefSufParams:
    """main.cpp:93-110 -- LEN = int(avg)+6, L = int(LEN*SCALE), RSOEMO = int(LEN*(SCALE+1)/2) in float."""
    LEN = int(avg_trimmed_len) + 2 * TRIM
    L = int(np.float32(LEN) * np.float32(SCALE))
    RS = int(np.float32(LEN) * (np.float32(SCALE) + np.float32(1)) / np.float32(2))
    return PrefSufParams(min_overlap=L, rs_min_overlap=RS)


def make_genome(size: int, rng: np.random.Generator, repeats: int = 0, repeat_len=(1000, 10000)) -> np.ndarray:
    g = rng.integers(0, 4, size=size, dtype=np.uint8)
    for _ in range(repeats):
        ln = int(rng.integers(repeat_len[0], repeat_len[1] + 1))
        ln = min(ln, size // 4)
        src = int(rng.integers(0, size - ln))
        dst = int(rng.integers(0, size - ln))
        g[dst : dst + ln] = g[src : src + ln].copy()
    return g


def _substitute(codes: np.ndarray, p: float, rng: np.random.Generator) -> np.ndarray:
    if p <= 0:
        return codes
    hit = rng.random(codes.shape) < p
    delta = rng.integers(1, 4, size=codes.shape, dtype=np.uint8)
    return np.where(hit, (codes + delta) & 3, codes).astype(np.uint8)


def sample_single_end(genome: np.ndarray, read_len: int, coverage: float, rng, error: float = 0.0) -> np.ndarray:
    """(n, read_len) code matrix of single-end reads (random strand)."""
    n = int(genome.shape[0] * coverage / read_len)
    start = rng.integers(0, genome.shape[0] - read_len + 1, size=n)
    idx = start[:, None] + np.arange(read_len)[None, :]
    reads = genome[idx]
    flip = rng.random(n) < 0.5
    reads[flip] = revcomp_codes(reads[flip])
    return _substitute(reads, error, rng)


def sample_paired_end(genome: np.ndarray, read_len: int, coverage: float, rng, error: float = 0.0):
    """Two (n, read_len) matrices: mate 1 and mate 2 (reverse complement of the fragment's far end)."""
    n = int(genome.shape[0] * coverage / (2 * read_len))
    frag = np.maximum(read_len, np.rint(rng.normal(500, 50, size=n)).astype(np.int64))
    frag = np.minimum(frag, genome.shape[0])
    start = (rng.random(n) * (genome.shape[0] - frag + 1)).astype(np.int64)
    ar = np.arange(read_len)[None, :]
    m1 = genome[start[:, None] + ar]
    m2 = revcomp_codes(genome[(start + frag - read_len)[:, None] + ar])
    swap = rng.random(n) < 0.5
    m1s = np.where(swap[:, None], m2, m1)
    m2s = np.where(swap[:, None], m1, m2)
    return _substitute(m1s, error, rng), _substitute(m2s, error, rng)


def fasta_text(m: np.ndarray) -> bytes:
    """Two-line FASTA (the only FASTA layout the reference reads, InputReader.cpp:151-154) of an (n, len) code matrix."""
    n, ln = m.shape
    rec = np.empty((n, 1 + 8 + 1 + ln + 1), np.uint8)  # ">" + 8-digit id + "\n" + sequence + "\n"
    rec[:, 0] = ord(">")
    ids = np.arange(n, dtype=np.int64)
    for d in range(8):
        rec[:, 8 - d] = ord("0") + (ids // 10 ** d) % 10
    rec[:, 9] = ord("\n")
    rec[:, 10 : 10 + ln] = np.frombuffer(b"ACGT", np.uint8)[m]
    rec[:, 10 + ln] = ord("\n")
    return rec.tobytes()


def strand_nodes(m1: np.ndarray, m2: np.ndarray | None = None) -> np.ndarray:
    """Trim 3+3 and lay records out as graph nodes in the reference's id order."""
    def trim(m):
        return m[:, TRIM : m.shape[1] - TRIM]

    t1 = trim(m1)
    n, l = t1.shape
    if m2 is None:
        out = np.empty((2 * n, l), dtype=np.uint8)
        out[0::2] = revcomp_codes(t1)
        out[1::2] = t1
        return out
    t2 = trim(m2)
    out = np.empty((4 * n, l), dtype=np.uint8)
    out[0::4] = revcomp_codes(t1)
    out[1::4] = t1
    out[2::4] = revcomp_codes(t2)
    out[3::4] = t2
    return out


def remove_duplicate_nodes(nodes: np.ndarray) -> np.ndarray:
    """Equal-length case of ReadPreprocess: among identical strand-reads only the greatest id survives.

    Returns the surviving rows in their original order (the renumbering of main.cpp:150-232).
    """
    packed = np.ascontiguousarray(pack_matrix(nodes))
    key = packed.view(np.dtype((np.void, packed.shape[1] * 4))).reshape(-1)
    n = key.shape[0]
    _, first_rev = np.unique(key[::-1], return_index=True)
    keep = np.zeros(n, dtype=bool)
    keep[n - 1 - first_rev] = True
    # a strand-read and its twin are duplicates of mirrored partners, so pairs stay intact
    assert np.array_equal(keep[0::2], keep[1::2])
    return nodes[keep]


def remove_prefix_reads_general(reads: list) -> list:
    """Variable-length case of ReadPreprocess.cpp:13-77 (small inputs; plain Python).

    ``reads`` is a list of uint8 code arrays laid out as (revcomp, forward) twins.  A read is
    dropped when another read starts with it (equal reads: all but the greatest id); a proper
    prefix also drops its twin ``id ^ 1``.  Survivors keep their order; twins are dropped together.
    """
    n = len(reads)
    keys = [bytes(r.tolist()) for r in reads]
    order = sorted(range(n), key=lambda i: (keys[i], i))
    drop = np.zeros(n, dtype=bool)
    for a, b in zip(order[:-1], order[1:]):
        ka, kb = keys[a], keys[b]
        if kb.startswith(ka):
            drop[a] = True
            if len(ka) < len(kb):
                drop[a ^ 1] = True
    drop_pair = drop[0::2] | drop[1::2]
    drop = np.repeat(drop_pair, 2)
    return [r for i, r in enumerate(reads) if not drop[i]]


@dataclass
class Workload:
    name: str
    reads: ReadSet
    params: PrefSufParams
    records: int  # FASTA records that produced the nodes
    genome_size: int


def make_workload(name: str, genome_size: int, read_len: int, coverage: float, paired: bool, seed: int,
                  error: float = 0.0, repeats: int = 0) -> Workload:
    cache = os.environ.get("ALGA_SYNTH_CACHE")  # experiments on the GPU box: several processes, one generation
    if cache:
        path = os.path.join(cache, f"synth_{genome_size}_{read_len}_{coverage}_{int(paired)}_{seed}_{error}_{repeats}.npz")
        if os.path.exists(path):
            z = np.load(path)
            rs = ReadSet(z["words"], z["word_off"], z["len_nt"])
            return Workload(name, rs, PrefSufParams(int(z["p"][0]), int(z["p"][1])), int(z["p"][2]), genome_size)
        w = _make_workload(name, genome_size, read_len, coverage, paired, seed, error, repeats)
        os.makedirs(cache, exist_ok=True)
        np.savez(path, words=w.reads.words, word_off=w.reads.word_off, len_nt=w.reads.len_nt,
                 p=np.array([w.params.min_overlap, w.params.rs_min_overlap, w.records], np.int64))
        return w
    return _make_workload(name, genome_size, read_len, coverage, paired, seed, error, repeats)


def _make_workload(name: str, genome_size: int, read_len: int, coverage: float, paired: bool, seed: int,
                   error: float = 0.0, repeats: int = 0) -> Workload:
    rng = np.random.default_rng(seed)
    genome = make_genome(genome_size, rng, repeats=repeats)
    if paired:
        m1, m2 = sample_paired_end(genome, read_len, coverage, rng, error)
        nodes = strand_nodes(m1, m2)
        records = 2 * m1.shape[0]
    else:
        m1 = sample_single_end(genome, read_len, coverage, rng, error)
        nodes = strand_nodes(m1)
        records = m1.shape[0]
    nodes = remove_duplicate_nodes(nodes)
    params = derive_params(float(nodes.shape[1]))
    return Workload(name, from_code_matrix(nodes), params, records, genome_size)


# BASELINE.json configs (seeds per SURVEY.md §8(d))
CONFIGS = {
    "cfg1": dict(genome_size=1_000_000, read_len=100, coverage=30, paired=False, seed=1),
    "cfg2": dict(genome_size=4_600_000, read_len=150, coverage=50, paired=True, seed=2),
    "cfg3": dict(genome_size=4_600_000, read_len=150, coverage=50, paired=True, seed=3, error=0.01),
    "cfg4": dict(genome_size=100_000_000, read_len=150, coverage=50, paired=True, seed=4),
    "cfg5": dict(genome_size=250_000_000, read_len=100, coverage=40, paired=True, seed=5, repeats=200),
}


def make_config(name: str, scale: float = 1.0) -> Workload:
    """One of the BASELINE.json configs; ``scale`` shrinks the genome (same read length and coverage)."""
    kw = dict(CONFIGS[name])
    kw["genome_size"] = max(20_000, int(kw["genome_size"] * scale))
    return make_workload(name if scale == 1.0 else f"{name}@{scale:g}", **kw)


def make_variable_length(genome_size: int, n_reads: int, len_lo: int, len_hi: int, seed: int, error: float = 0.0,
                         repeats: int = 0, dedupe: bool = True) -> ReadSet:
    """Variable-length single-end reads (contig-like second call site)."""
    rng = np.random.default_rng(seed)
    genome = make_genome(genome_size, rng, repeats=repeats, repeat_len=(200, 2000))
    reads = []
    for _ in range(n_reads):
        ln = int(rng.integers(len_lo, len_hi + 1))
        s = int(rng.integers(0, genome_size - ln + 1))
        r = _substitute(genome[s : s + ln].copy(), error, rng)
        if rng.random() < 0.5:
            r = revcomp_codes(r)
        reads.append(revcomp_codes(r))
        reads.append(r)
    if dedupe:
        reads = remove_prefix_reads_general(reads)
    return from_code_list(reads)

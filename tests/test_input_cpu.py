"""CPU tests of the input stage's oracle (oracle/input_oracle.c): InputReader::readInput against the reads the
unmodified reference produced (tests/golden/in_*.npz, live against oracle/_ref where it exists), and the whole path
from the files to the graph -- reader, ReadPreprocess, the renumbering of main.cpp:150-232, GraphCreatorPrefSuf --
against the graph the STOCK binary serialises (tests/golden/front_*.npz)."""
import hashlib
import os

import numpy as np
import pytest

from alga_b200.input_reader import driver_params
from alga_b200.readset import ReadSet
from oracle import harness, oracle
from tests.cases import FRONT_CASES, INPUT_CASES, front_case, input_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def text_sha(t1, t2):
    h = hashlib.sha256(t1)
    if t2 is not None:
        h.update(b"|")
        h.update(t2)
    return h.hexdigest()


def same_reads(a: ReadSet, len_nt, word_off, words):
    return (a.n == len_nt.shape[0] and np.array_equal(a.len_nt, len_nt) and np.array_equal(a.word_off, word_off)
            and np.array_equal(a.words, words))


def gather(rs: ReadSet, ids) -> ReadSet:
    """Reads `ids` of rs, in that order."""
    cnt = (rs.word_off[1:] - rs.word_off[:-1])[ids].astype(np.int64)
    off = np.zeros(len(ids) + 1, np.uint64)
    np.cumsum(cnt, out=off[1:])
    src = np.repeat(rs.word_off[:-1][ids].astype(np.int64) - off[:-1].astype(np.int64), cnt) + np.arange(int(off[-1]))
    return ReadSet(rs.words[src] if len(src) else np.zeros(0, np.uint32), off, rs.len_nt[ids])


def oracle_front(t1, t2, ft):
    """main.cpp:82-291 composed from the oracle's pieces -> (n, edges)."""
    rs, _ = oracle.read_input(t1, t2, ft)
    prm = driver_params(rs)
    mask = oracle.prefix_reads(rs, 2)
    old, po = oracle.remap(rs.len_nt, mask)
    rs2 = gather(rs, old)
    ln = rs2.len_nt.copy()
    ln[ln < prm["li_kmer_intervals"] + prm["li_kmer_length"]] = 0  # main.cpp:253-266
    rs3 = ReadSet(rs2.words, rs2.word_off, ln)
    return rs3, po, harness.sort_edges(oracle.prefsuf(rs3, prm["min_overlap"], prm["rs_min_overlap"]))


@pytest.mark.parametrize("name", INPUT_CASES)
def test_read_input_oracle_matches_reference_fixture(name):
    t1, t2, ft, extra = input_case(name)
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    assert str(g["input_sha"]) == text_sha(t1, t2), "generator drifted: rerun tests/golden/make_golden.py"
    rs, info = oracle.read_input(t1, t2, ft, rna=int("--rna=1" in extra))
    assert same_reads(rs, g["len_nt"], g["word_off"], g["words"])
    assert rs.n % 2 == 0 and np.array_equal(rs.len_nt[0::2], rs.len_nt[1::2])
    assert 2 * (info["n_with_n"] + info["n_str"]) == int((rs.len_nt == 0).sum())


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["in_fasta_pe", "in_plain", "in_rna", "in_truncated"])
def test_read_input_oracle_matches_reference_live(name):
    t1, t2, ft, extra = input_case(name)
    ref = harness.run_read_input(t1, t2, ft, extra=extra)
    rs, _ = oracle.read_input(t1, t2, ft, rna=int("--rna=1" in extra))
    assert same_reads(rs, ref.len_nt, ref.word_off, ref.words)


def test_read_input_error_cases():
    ok = b">a\nACGTACGTACGTACGTACGTAGCATCGATCGACTAGCTAGCTACGACTAGC\n"
    with pytest.raises(ValueError):
        oracle.read_input(ok + b">b\nACGTacgtACGTACGTACGTACGTACGT\n", None, oracle.INPUT_FASTA)  # lower case: the reference exits
    # CR is part of the line: a long read loses it to the end trimming, a short (untrimmed) one keeps it -> bad character
    rs, _ = oracle.read_input(ok.replace(b"\n", b"\r\n"), None, oracle.INPUT_FASTA)
    assert rs.len_nt.tolist() == [46, 46]
    with pytest.raises(ValueError):
        oracle.read_input(b">s\r\nACGTTGCA\r\n", None, oracle.INPUT_FASTA)
    with pytest.raises(ValueError):
        oracle.read_input(ok + ok, ok, oracle.INPUT_FASTA)  # mate files of different length
    if harness.available():
        with pytest.raises(RuntimeError):
            harness.run_read_input(ok + b">b\nACGTacgtACGTACGTACGTACGTACGT\n", None, oracle.INPUT_FASTA)
    # a bad character behind the point where reading stops is never seen
    rs, _ = oracle.read_input(ok + b">e\n\n>b\nxxxx\n", None, oracle.INPUT_FASTA)
    assert rs.n == 2


def test_remap_rules():
    # units: (0,1) mates, (2,3) mates, ... ; unit u = reads 2u, 2u+1
    ln = np.full(16, 50, np.uint32)
    mask = np.zeros(16, np.uint8)
    mask[[2, 3]] = 1      # unit 1 (second mate of pair 0) removed -> unit 0 alone: offset 0
    mask[[8, 9]] = 1      # unit 4 (first mate of pair 2) removed -> unit 5 alone: offset 0
    ln[[12, 13]] = 0      # unit 6 nullptr from the reader, unit 7 alone
    old, po = oracle.remap(ln, mask)
    assert old.tolist() == [0, 1, 4, 5, 6, 7, 10, 11, 14, 15]
    assert po.tolist() == [0, 0, 1, 1, 2, 2, 0, 0, 0, 0]
    with pytest.raises(ValueError):
        m2 = np.zeros(16, np.uint8)
        m2[5] = 1         # a read without its reverse complement: main.cpp:173 asserts
        oracle.remap(ln, m2)
    old, po = oracle.remap(np.zeros(0, np.uint32), None)
    assert old.shape == (0,)
    # odd number of units: the last first-mate has no partner slot
    old, po = oracle.remap(np.full(6, 9, np.uint32), None)
    assert po.tolist() == [1, 1, 2, 2, 0, 0]


@pytest.mark.parametrize("name", FRONT_CASES)
def test_files_to_graph_matches_stock_binary(name):
    t1, t2, ft = front_case(name)
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    assert str(g["input_sha"]) == text_sha(t1, t2), "generator drifted: rerun tests/golden/make_golden.py"
    rs, po, edges = oracle_front(t1, t2, ft)
    assert rs.n == int(g["n"])
    assert np.array_equal(edges, g["edges"])
    assert np.array_equal(po, g["paired_offset"])  # Global::pairedReadOffset after the reference's own main()


def test_file_type_follows_the_extension():
    """Params.cpp:332-335."""
    from alga_b200.input_reader import FASTA, FASTQ, PLAIN, file_type_of
    assert file_type_of("/data/x_1.fasta") == FASTA and file_type_of("reads.pfasta") == FASTA
    assert file_type_of("a.b/reads.fastq") == FASTQ and file_type_of("reads.fq") == FASTQ
    assert file_type_of("reads.txt") == PLAIN and file_type_of("dir.v2/reads") == PLAIN and file_type_of("reads.fa") == PLAIN


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref not built")
def test_read_input_oracle_matches_reference_on_random_files():
    """Random small files with every oddity the reader has a rule for (spaces, short lines, N, U, repeats, empty lines, missing
    final newline, FASTQ quality noise): the oracle must produce exactly the reference's Global::READS -- or fail where it exits."""
    n_checked = n_failed = 0
    for trial, (t1, t2, ft, paired) in enumerate(random_input_files(2024, 40)):
        try:
            ref = harness.run_read_input(t1, t2, ft)
        except RuntimeError:
            ref = None
        try:
            got, _ = oracle.read_input(t1, t2, ft)
        except ValueError:
            got = None
        if ref is None:
            # the reference died: a bad character, or -- with unequal mate files -- its out-of-range interleave (undefined)
            n_failed += 1
            assert got is None or paired, f"trial {trial}: the reference exits, the oracle does not"
            continue
        if got is None:
            assert paired, f"trial {trial}: the oracle fails, the reference does not"  # unequal mate files only
            continue
        assert same_reads(got, ref.len_nt, ref.word_off, ref.words), f"trial {trial} (type {ft}, paired {paired})"
        n_checked += 1
    assert n_checked >= 20


def random_input_files(seed, n_trials):
    """Small random input files of all three types with the reader's special cases sprinkled in -> (text1, text2, type, paired)."""
    rng = np.random.default_rng(seed)
    nt = b"ACGT"

    def seq():
        u = rng.random()
        ln = int(rng.integers(1, 60)) if u < 0.25 else int(rng.integers(60, 200))
        s = bytes(nt[i] for i in rng.integers(0, 4, size=ln))
        if u > 0.9:
            per = int(rng.integers(1, 25))
            s = (s[:per] * (ln // per + 1))[:ln]
        if rng.random() < 0.1:
            k = int(rng.integers(0, len(s)))
            s = s[:k] + b"N" + s[k + 1:]
        if rng.random() < 0.1:
            s = s.replace(b"T", b"U")
        if rng.random() < 0.1:
            s = b" " * int(rng.integers(1, 4)) + s
        if rng.random() < 0.1:
            s = s + b" extra words"
        return s

    for trial in range(n_trials):
        ft = int(rng.integers(0, 3))
        paired = rng.random() < 0.5
        n = int(rng.integers(0, 12))
        texts = []
        for _ in range(2 if paired else 1):
            recs = []
            for i in range(n):
                s = seq()
                if rng.random() < 0.03:
                    s = b""                      # reading stops here (InputReader.cpp:284)
                if rng.random() < 0.02:
                    s = s[:1] + b"x" + s[2:]     # the reference exits (InputReader.cpp:324-327)
                if ft == oracle.INPUT_FASTA:
                    recs.append(b">r%d some text\n" % i + s + b"\n")
                elif ft == oracle.INPUT_FASTQ:
                    q = bytes(rng.integers(33, 74, size=len(s), dtype=np.uint8).tolist())
                    recs.append(b"@r%d\n" % i + s + b"\n+\n" + q + b"\n")
                else:
                    recs.append(s.split(b" ")[-1 if s.startswith(b" ") else 0] + [b"\n", b" ", b"\t", b"\n\n"][int(rng.integers(0, 4))])
            t = b"".join(recs)
            if t.endswith(b"\n") and rng.random() < 0.3:
                t = t[:-1]
            texts.append(t)
        yield texts[0], (texts[1] if paired else None), ft, paired


@pytest.mark.skipif(not os.path.isfile(harness.STOCK), reason="oracle/_ref/ALGA not built")
def test_minimum_overlap_override_follows_the_driver():
    """`-l 40` (Params.cpp:488-497, main.cpp:112-115): minimum overlap 40, RSOEMO = (40 + LEN) / 2, LI_KMER_LENGTH = 40 -- the
    rule alga_gpu_files_to_graph applies for alga_driver_params.min_overlap > 0, checked against the stock binary's graph."""
    t1, t2, ft = front_case("front_short")
    n_ref, e_ref = harness.run_stock_graph(t1, t2, ft, extra=("-l", "40"))
    rs, _ = oracle.read_input(t1, t2, ft)
    alive = rs.len_nt[rs.len_nt > 0]
    LEN = int(alive.astype(np.float64).sum() / alive.shape[0]) + 6
    lmin, rsmin, li = 40, (40 + LEN) // 2, 40
    old, _ = oracle.remap(rs.len_nt, oracle.prefix_reads(rs, 2))
    rs2 = gather(rs, old)
    ln = rs2.len_nt.copy()
    ln[ln < 3 + li] = 0
    edges = harness.sort_edges(oracle.prefsuf(ReadSet(rs2.words, rs2.word_off, ln), lmin, rsmin))
    assert rs2.n == n_ref
    assert np.array_equal(edges, e_ref)


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref not built")
def test_paired_read_offsets_match_the_reference_driver_on_random_sets():
    """Global::pairedReadOffset after the reference's own main() (harness mode `driver`) against the oracle's renumbering, on small
    random paired and single-end sets with N reads, repeats and duplicates (so that mates lose their partners in every way)."""
    from alga_b200 import synth
    from tests.cases import _fasta, _seq, _spice
    for seed in range(6):
        rng = np.random.default_rng(900 + seed)
        g = synth.make_genome(6_000, rng)
        if seed % 2 == 0:
            m1, m2 = synth.sample_paired_end(g, 120, 25, rng, 0.0)
            t1 = _fasta(_spice([_seq(r) for r in m1], rng, p_n=0.08, p_str=0.05, p_short=0.0, p_space=0.0))
            t2 = _fasta(_spice([_seq(r) for r in m2], rng, p_n=0.08, p_str=0.05, p_short=0.0, p_space=0.0), b"m")
        else:
            m = synth.sample_single_end(g, 100, 25, rng, 0.0)
            t1, t2 = _fasta(_spice([_seq(r) for r in m], rng, p_n=0.08, p_str=0.05, p_short=0.0, p_space=0.0)), None
        po_ref = harness.run_driver_paired_offsets(t1, t2, oracle.INPUT_FASTA)
        rs, _ = oracle.read_input(t1, t2, oracle.INPUT_FASTA)
        _, po = oracle.remap(rs.len_nt, oracle.prefix_reads(rs, 2))
        assert np.array_equal(po, po_ref), f"seed {seed}"
        assert set(np.unique(po).tolist()) <= {0, 1, 2} and (po == 0).any()

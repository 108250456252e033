"""GPU parity of the input stage through the C ABI: alga_gpu_read_input (InputReader::readInput), alga_gpu_remap_reads
(main.cpp:150-232) and the whole path from the files to the overlap graph (main.cpp:82-291) against the oracle, the
reads the unmodified reference produced (tests/golden/in_*.npz) and the graph the stock binary serialises
(tests/golden/front_*.npz)."""
import os

import numpy as np
import pytest

from alga_b200 import _lib, synth
from alga_b200.input_reader import (FASTA, InputReader, PinnedText, build_overlap_graph, build_overlap_graph_staged,
                                    remap_reads)
from alga_b200.readset import ReadSet
from oracle import oracle
from tests.cases import FRONT_CASES, INPUT_CASES, PREPROCESS_CASES, _fasta, _seq, front_case, input_case, preprocess_case
from tests.test_input_cpu import gather, oracle_front

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def dense(len_nt, word_off, words, width=None):
    """(n, width) matrix of the blocks of every read, zero behind its last block; removed reads (length 0) are all zero."""
    n = len_nt.shape[0]
    cnt = (len_nt.astype(np.int64) + 15) // 16
    width = int(cnt.max()) if width is None and n else (width or 1)
    m = np.zeros((n, max(width, 1)), np.uint32)
    col = np.arange(m.shape[1])[None, :]
    sel = col < cnt[:, None]
    src = (word_off[:-1].astype(np.int64)[:, None] + col)[sel]
    m[sel] = words[src]
    return m


def assert_same_reads(got: ReadSet, len_nt, word_off, words):
    """`got` has a fixed stride, the reference side ragged offsets: lengths equal, blocks equal, padding zero."""
    assert got.n == len_nt.shape[0]
    assert np.array_equal(got.len_nt, len_nt)
    if got.n == 0:
        return
    stride = int(got.word_off[1] - got.word_off[0])
    assert stride >= max(1, (int(len_nt.max()) + 15) // 16)
    assert np.array_equal(got.words.reshape(got.n, stride), dense(len_nt, word_off, words, stride))


@pytest.mark.parametrize("name", INPUT_CASES)
def test_read_input_matches_oracle_and_reference(gpu, name):
    t1, t2, ft, extra = input_case(name)
    rd = InputReader(ft, rna="--rna=1" in extra)
    got = rd.readInput(t1, t2)
    want, info = oracle.read_input(t1, t2, ft, rna=int("--rna=1" in extra))
    assert_same_reads(got, want.len_nt, want.word_off, want.words)
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    assert_same_reads(got, g["len_nt"], g["word_off"], g["words"])
    assert rd.info["n_with_n"] == info["n_with_n"] and rd.info["n_str"] == info["n_str"]
    assert rd.timing["kernel_launches"] > 0 or got.n == 0


def test_read_input_trim_and_threshold_options(gpu):
    t1, _, ft, _ = input_case("in_fasta_se")
    for tl, tr, thr in [(0, 0, 20), (5, 1, 20), (3, 3, 7), (0, 12, 40)]:
        got = InputReader(ft, trim_left=tl, trim_right=tr, str_threshold=thr).readInput(t1)
        want, _ = oracle.read_input(t1, None, ft, trim_left=tl, trim_right=tr, str_threshold=thr)
        assert_same_reads(got, want.len_nt, want.word_off, want.words)


def test_read_input_errors(gpu):
    ok = b">a\nACGTACGTACGTACGTACGTAGCATCGATCGACTAGCTAGCTACGACTAGC\n"
    with pytest.raises(_lib.AlgaGpuError) as e:
        InputReader(FASTA).readInput(ok + b">b\nACGTacgtACGTACGTACGTACGTACGT\n")
    assert e.value.code == -1 and "record 1" in str(e.value)
    with pytest.raises(_lib.AlgaGpuError):
        InputReader(FASTA).readInput(b">s\r\nACGTTGCA\r\n")
    with pytest.raises(_lib.AlgaGpuError):
        InputReader(FASTA).readInput(ok + ok, ok)
    with pytest.raises(_lib.AlgaGpuError):
        InputReader(7).readInput(ok)
    # a bad character behind the point where reading stops is never seen (InputReader.cpp:284)
    assert InputReader(FASTA).readInput(ok + b">e\n\n>b\nxxxx\n").n == 2
    # CR LF on a long read: the CR falls to the end trimming
    assert InputReader(FASTA).readInput(ok.replace(b"\n", b"\r\n")).len_nt.tolist() == [46, 46]


def test_read_input_block_boundaries(gpu):
    """Record starts and line ends at every position relative to the 16-byte / 4 KiB granules of the mark kernels."""
    rng = np.random.default_rng(5)
    for pad in range(0, 40, 3):
        seqs = [_seq(rng.integers(0, 4, size=int(rng.integers(30, 130)))) for _ in range(400)]
        text = b">" + b"x" * pad + b"\n" + seqs[0] + b"\n" + _fasta(seqs[1:])
        got = InputReader(FASTA).readInput(text)
        want, _ = oracle.read_input(text, None, oracle.INPUT_FASTA)
        assert_same_reads(got, want.len_nt, want.word_off, want.words)


def test_read_input_full_size_properties(gpu):
    """BASELINE config 2 sized files (2 x 766 666 records of 150 bp): strands are reverse complements of each other,
    mates interleave, lengths are 144, and the packed reads equal the host-side packing of the same records."""
    rng = np.random.default_rng(2)
    g = synth.make_genome(4_600_000, rng)
    m1, m2 = synth.sample_paired_end(g, 150, 50, rng, 0.0)
    rd = InputReader(FASTA)
    got = rd.readInput(synth.fasta_text(m1), synth.fasta_text(m2))
    n = m1.shape[0]
    assert got.n == 4 * n
    want = synth.strand_nodes(m1, m2)  # host-side trimming / strand layout of the same records
    from alga_b200.readset import pack_matrix
    ref_words = pack_matrix(want)
    alive = got.len_nt > 0
    assert np.array_equal(got.len_nt[alive], np.full(int(alive.sum()), 144, np.uint32))
    assert np.array_equal(alive[0::2], alive[1::2])
    stride = int(got.word_off[1] - got.word_off[0])
    assert stride == 9
    assert np.array_equal(got.words.reshape(-1, 9)[alive], ref_words[alive])
    assert int((~alive).sum()) == 2 * (rd.info["n_with_n"] + rd.info["n_str"])
    assert (~alive).sum() < 100  # random 144-mers are practically never short-period repeats


@pytest.mark.parametrize("name", PREPROCESS_CASES + ["in_fasta_pe", "in_fasta_se"])
def test_remap_matches_oracle(gpu, name):
    if name.startswith("in_"):
        t1, t2, ft, _ = input_case(name)
        rs, _ = oracle.read_input(t1, t2, ft)
    else:
        rs = preprocess_case(name)
        if rs.n % 2:
            pytest.skip("odd read count")
    mask = oracle.prefix_reads(rs, 2)
    try:
        old, po = oracle.remap(rs.len_nt, mask)
    except ValueError:  # a palindromic duplicate leaves one strand without its twin: drop pairs together
        mask = np.repeat(mask[0::2] | mask[1::2], 2).astype(np.uint8)
        old, po = oracle.remap(rs.len_nt, mask)
    rm = remap_reads(rs, mask)
    assert np.array_equal(rm.old_id, old)
    assert np.array_equal(rm.paired_offset, po)
    want = gather(rs, old)
    assert_same_reads(rm.reads, want.len_nt, want.word_off, want.words)
    # nothing to remove, no mask: identity
    if rm.reads.n:
        again = remap_reads(rm.reads)
        assert np.array_equal(again.old_id, np.arange(rm.reads.n, dtype=np.uint32))


def test_remap_edge_cases(gpu):
    ln = np.full(16, 50, np.uint32)
    rs = ReadSet(np.arange(16 * 4, dtype=np.uint32), np.arange(17, dtype=np.uint64) * np.uint64(4), ln.copy())
    mask = np.zeros(16, np.uint8)
    mask[[2, 3, 8, 9]] = 1
    rs.len_nt[[12, 13]] = 0
    rm = remap_reads(rs, mask)
    assert rm.old_id.tolist() == [0, 1, 4, 5, 6, 7, 10, 11, 14, 15]
    assert rm.paired_offset.tolist() == [0, 0, 1, 1, 2, 2, 0, 0, 0, 0]
    bad = np.zeros(16, np.uint8)
    bad[5] = 1
    with pytest.raises(_lib.AlgaGpuError):
        remap_reads(rs, bad)
    empty = ReadSet(np.zeros(0, np.uint32), np.zeros(1, np.uint64), np.zeros(0, np.uint32))
    assert remap_reads(empty).reads.n == 0
    allgone = remap_reads(rs, np.ones(16, np.uint8))
    assert allgone.reads.n == 0


@pytest.mark.parametrize("name", FRONT_CASES)
@pytest.mark.parametrize("how", ["fused", "fused_pinned", "staged"])
def test_files_to_graph_matches_stock_binary(gpu, name, how):
    """main.cpp:82-291 on the GPU end to end: same node count and edge set as the graph the stock binary serialised."""
    t1, t2, ft = front_case(name)
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    if how == "staged":
        og = build_overlap_graph_staged(t1, t2, ft)
    elif how == "fused_pinned":
        og = build_overlap_graph(PinnedText(t1), PinnedText(t2) if t2 is not None else None, ft)
    else:
        og = build_overlap_graph(t1, t2, ft)
    assert og.reads.n == int(g["n"])
    assert np.array_equal(og.graph.edges(), g["edges"])


@pytest.mark.parametrize("name", ["in_fasta_se", "in_fasta_pe", "in_fastq_pe", "in_plain", "in_long", "in_no_eol", "in_truncated",
                                  "in_empty"])
def test_files_to_graph_matches_oracle_pipeline(gpu, name):
    """The fused call against the same path composed from the oracle's pieces: renumbered reads, pairedReadOffset, the
    removal of short reads, parameters and the edge set (inputs with N reads, repeats, short reads, ragged lengths)."""
    t1, t2, ft, _ = input_case(name)
    og = build_overlap_graph(t1, t2, ft)
    raw, _ = oracle.read_input(t1, t2, ft)
    if raw.n == 0 or not (raw.len_nt > 0).any():
        assert og.reads.n == 0 and og.graph.n_edges == 0
        return
    rs, po, edges = oracle_front(t1, t2, ft)
    assert_same_reads(og.reads, rs.len_nt, rs.word_off, rs.words)
    assert np.array_equal(og.paired_offset, po)
    assert np.array_equal(og.graph.edges(), edges)
    old, _ = oracle.remap(raw.len_nt, oracle.prefix_reads(raw, 2))
    assert np.array_equal(og.old_id, old)


def test_files_to_graph_errors_leave_the_library_usable(gpu):
    """A failing call (bad character, unequal mate files) must not poison the cached workspace of the next one."""
    t1, t2, ft = front_case("front_pe")
    g = np.load(os.path.join(GOLD, "front_pe.npz"))
    cut = t1.index(b"\n>", 2000) + 1  # a record boundary
    bad = t1[:cut] + b">x\nACGTACGTacgtACGTACGTACGTACGT\n" + t1[cut:]
    for _ in range(2):
        with pytest.raises(_lib.AlgaGpuError) as e:
            build_overlap_graph(bad, t2, ft)
        assert "character" in str(e.value)
        with pytest.raises(_lib.AlgaGpuError):
            build_overlap_graph(t1, t2[: len(t2) // 2], ft)
        og = build_overlap_graph(t1, t2, ft)
        assert og.reads.n == int(g["n"]) and np.array_equal(og.graph.edges(), g["edges"])


def test_files_to_graph_overrides_and_remove_types(gpu):
    """-l style overrides of the derived parameters and the other REMOVE_PREF_READS_TYPE settings, against the oracle's pieces."""
    from alga_b200.input_reader import driver_params
    from tests.test_input_cpu import gather
    t1, t2, ft = front_case("front_se")
    raw, _ = oracle.read_input(t1, t2, ft)
    for remove_type in (1, 0):
        og = build_overlap_graph(t1, t2, ft, remove_type=remove_type)
        prm = driver_params(raw)
        mask = oracle.prefix_reads(raw, 1) if remove_type else None
        old, po = oracle.remap(raw.len_nt, mask)
        rs2 = gather(raw, old)
        ln = rs2.len_nt.copy()
        ln[ln < 3 + prm["li_kmer_length"]] = 0
        want = oracle.prefsuf(ReadSet(rs2.words, rs2.word_off, ln), prm["min_overlap"], prm["rs_min_overlap"])
        assert np.array_equal(og.old_id, old) and np.array_equal(og.paired_offset, po)
        assert np.array_equal(og.graph.edges(), want)
    og = build_overlap_graph(t1, t2, ft, min_overlap=40, rs_min_overlap=70)
    assert og.params["min_overlap"] == 40 and og.params["rs_min_overlap"] == 70 and og.params["li_kmer_length"] == 40


def test_read_input_offsets_beyond_4_gib(gpu):
    """Records that start behind byte 2^32 of the file (a 4 GiB header line in front of them): positions are 64-bit from the
    mark kernels to the packing.  Pageable text, so the chunked upload is exercised with hundreds of chunks as well."""
    rng = np.random.default_rng(9)
    seqs = [_seq(rng.integers(0, 4, size=int(rng.integers(60, 160)))) for _ in range(300)]
    tail = _fasta(seqs[1:])
    text = b">" + b"x" * (2**32 + 12345) + b"\n" + seqs[0] + b"\n" + tail
    got = InputReader(FASTA).readInput(text)
    del text
    want, _ = oracle.read_input(b">h\n" + seqs[0] + b"\n" + tail, None, oracle.INPUT_FASTA)
    assert_same_reads(got, want.len_nt, want.word_off, want.words)


def test_build_graph_cli_writes_the_file_alga_loads(gpu, tmp_path):
    """python -m alga_b200 build-graph: the graph file of the zero-code-change boundary (main.cpp:242), from the input files."""
    import json
    import subprocess
    import sys

    from alga_b200.graph_file import graph_file_name, read_graph

    t1, t2, ft = front_case("front_pe")
    (tmp_path / "x_1.fasta").write_bytes(t1)
    (tmp_path / "x_2.fasta").write_bytes(t2)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "alga_b200", "build-graph", "--file1", str(tmp_path / "x_1.fasta"), "--file2",
                        str(tmp_path / "x_2.fasta"), "--out-dir", str(tmp_path)], cwd=root, stdout=subprocess.PIPE,
                       stderr=subprocess.PIPE, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    g = np.load(os.path.join(GOLD, "front_pe.npz"))
    back = read_graph(str(tmp_path / graph_file_name("x_1.fasta")))
    assert info["nodes"] == back.n == int(g["n"])
    assert np.array_equal(back.edges(), g["edges"])


def test_read_input_random_files(gpu):
    """The random small files of the CPU suite (all three types, spaces, short lines, N, U, repeats, empty lines, bad characters,
    unequal mate files): the CUDA reader and the oracle agree on the reads, and fail on the same inputs."""
    from tests.test_input_cpu import random_input_files
    n_ok = n_err = 0
    for trial, (t1, t2, ft, paired) in enumerate(random_input_files(77, 120)):
        try:
            want, _ = oracle.read_input(t1, t2, ft)
        except ValueError:
            want = None
        try:
            got = InputReader(ft).readInput(t1, t2)
        except _lib.AlgaGpuError:
            got = None
        assert (got is None) == (want is None), f"trial {trial} (type {ft}, paired {paired})"
        if want is None:
            n_err += 1
            continue
        assert_same_reads(got, want.len_nt, want.word_off, want.words)
        n_ok += 1
    assert n_ok >= 60 and n_err >= 1

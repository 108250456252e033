"""CPU: the counter-based generator (alga_b200/synth_dev.py), the edge-set digest and the full-size reference goldens."""
import glob
import hashlib
import json
import os

import numpy as np
import torch

from alga_b200 import edge_hash, readset, synth, synth_dev
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))


def test_generator_is_deterministic_and_twinned():
    a = synth_dev.make_config("cfg2", 0.01)
    b = synth_dev.make_config("cfg2", 0.01)
    assert torch.equal(a.words, b.words) and a.params == b.params and a.records == b.records
    rs = a.to_readset()
    assert rs.n % 2 == 0 and a.len_nt == 144 and (a.params.min_overlap, a.params.rs_min_overlap) == (82, 116)
    for i in (0, 2, rs.n - 2):  # even id = reverse complement of the odd one (InputReader.cpp:78-85)
        assert np.array_equal(readset.revcomp_codes(rs.codes(i + 1)), rs.codes(i))
    other = synth_dev.make_config("cfg2", 0.01, seed_offset=1)
    assert not torch.equal(a.words[:1000], other.words[:1000])


def test_duplicate_removal_matches_the_numpy_restatement():
    """remove_duplicate_rows == synth.remove_duplicate_nodes (greatest id among identical strand-reads survives)."""
    g = synth_dev.make_genome(3000, 7, "cpu")
    words, _ = synth_dev._strand_words(g, 100, 60, True, 7, 0.0)  # 60x over 3 kbp: plenty of duplicates
    keep = synth_dev.remove_duplicate_rows(words).numpy()
    w = words.numpy().view(np.uint32)
    key = np.ascontiguousarray(w).view(np.dtype((np.void, w.shape[1] * 4))).reshape(-1)
    n = key.shape[0]
    _, first_rev = np.unique(key[::-1], return_index=True)
    want = np.zeros(n, bool)
    want[n - 1 - first_rev] = True
    assert 0 < want.sum() < n and np.array_equal(keep, want)


def test_single_end_and_errors():
    w = synth_dev.make_config("cfg1", 0.02)
    assert w.len_nt == 94 and w.records * 2 >= w.n
    e = synth_dev.make_config("cfg3", 0.005)
    c = synth_dev.make_config("cfg2", 0.005)
    assert e.n >= c.n  # substitutions break duplicates


def test_packing_matches_readset():
    codes = torch.from_numpy(np.random.default_rng(3).integers(0, 4, size=(50, 94), dtype=np.uint8))
    got = synth_dev._pack(codes).numpy().view(np.uint32)
    assert np.array_equal(got, readset.pack_matrix(codes.numpy()))


def test_digest_is_order_independent_and_additive():
    rng = np.random.default_rng(5)
    n = 1000
    deg = rng.integers(0, 4, size=n)
    row_off = np.concatenate([[0], np.cumsum(deg)]).astype(np.int64)
    E = int(row_off[-1])
    nbr = rng.integers(0, n, size=E).astype(np.int32)
    off = rng.integers(0, 63, size=E).astype(np.int32)
    src = np.repeat(np.arange(n, dtype=np.int32), deg)
    edges = np.stack([src, nbr, off], axis=1)
    d0 = edge_hash.digest_edges(edges)
    assert d0 == edge_hash.digest_edges(edges[rng.permutation(E)])
    assert d0 == edge_hash.digest_csr(torch.from_numpy(row_off), torch.from_numpy(nbr), torch.from_numpy(off))
    k = 400  # two disjoint row ranges add up (how the ranks of a sharded build combine theirs)
    a = edge_hash.digest_csr(torch.from_numpy(row_off[: k + 1]), torch.from_numpy(nbr[: row_off[k]]), torch.from_numpy(off[: row_off[k]]))
    b = edge_hash.digest_csr(torch.from_numpy(row_off[k:] - row_off[k]), torch.from_numpy(nbr[row_off[k]:]),
                             torch.from_numpy(off[row_off[k]:]), first_row=k)
    assert edge_hash.add(a, b) == d0
    changed = edges.copy()
    changed[7, 2] += 1
    assert edge_hash.digest_edges(changed) != d0
    assert edge_hash.digest_edges(np.zeros((0, 3), np.int32)) == (0, 0)


def test_full_goldens_are_well_formed():
    files = sorted(glob.glob(os.path.join(HERE, "golden", "full_*.json")))
    assert len(files) >= 5
    for f in files:
        g = json.load(open(f))
        assert g["nodes"] > 0 and g["edges"] > 0 and len(g["digest"]) == 2 and len(g["sha256"]) == 64
        if g["workload"].startswith("cfg3"):
            assert g["threads"] == 1  # data with errors: only the --threads=1 order is canonical


def test_oracle_reproduces_the_full_cfg1_golden():
    """The plain-C restatement against the UNMODIFIED reference on the whole of BASELINE config 1 (518 k nodes)."""
    for gen in ("dev", "np"):
        g = json.load(open(os.path.join(HERE, "golden", f"full_cfg1_{gen}.json")))
        if gen == "dev":
            w = synth_dev.make_config("cfg1")
            rs, p = w.to_readset(), w.params
        else:
            w = synth.make_config("cfg1")
            rs, p = w.reads, w.params
        assert hashlib.sha256(np.ascontiguousarray(rs.words).tobytes()).hexdigest() == g["input_sha"], "generator drift"
        e = oracle.prefsuf(rs, p.min_overlap, p.rs_min_overlap)
        assert (rs.n, e.shape[0]) == (g["nodes"], g["edges"])
        assert [f"{x:016x}" for x in edge_hash.digest_edges(e)] == g["digest"]
        assert hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest() == g["sha256"]

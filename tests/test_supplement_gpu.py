"""GPU parity of the error-rate supplement (main.cpp:300-355): LI k-mers and the whole supplement through the C ABI
against the oracle and the golden graphs produced by the unmodified reference."""
import os

import numpy as np
import pytest

from alga_b200.graph_creator import Graph, GraphCreatorLI, GraphCreatorPrefSuf, li_kmers
from oracle import oracle
from tests.cases import SUPPLEMENT_CASES, supplement_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _graph_from_edges(n, e):
    deg = np.bincount(e[:, 0], minlength=n).astype(np.uint64)
    row_off = np.zeros(n + 1, np.uint64)
    np.cumsum(deg, out=row_off[1:])
    return Graph(n, row_off, np.ascontiguousarray(e[:, 1]), np.ascontiguousarray(e[:, 2]))


@pytest.mark.parametrize("prio", [(0, 1, 2, 3), (1, 2, 3, 0), (2, 3, 0, 1), (3, 0, 1, 2)])
@pytest.mark.parametrize("K,IV", [(35, 6), (20, 3), (63, 2)])
def test_li_kmers_match_oracle(gpu, prio, K, IV):
    """Read::getLIKmers (Read.cpp:145-226): hash mod 10^18+3 and position of the leftmost minimal K-mer per interval."""
    rs, *_ = supplement_case("sup_varlen")
    ids = np.flatnonzero(rs.len_nt >= max(K, 60)).astype(np.uint32)
    want_h, want_i = oracle.li_kmers(rs, ids, prio, K, IV)
    got_h, got_i = li_kmers(rs, ids, prio, K, IV)
    assert np.array_equal(got_i, want_i)
    assert np.array_equal(got_h, want_h)


@pytest.mark.parametrize("name", SUPPLEMENT_CASES)
def test_supplement_matches_oracle_and_reference(gpu, name):
    rs, lmin, rsmin, sp = supplement_case(name)
    g0 = GraphCreatorPrefSuf(rs, lmin, rsmin).startAlignmentGraphCreation()  # main.cpp:282-291 on the GPU
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    assert np.array_equal(g0.edges(), z["before"])
    li = GraphCreatorLI(rs, g0, **sp)
    g1 = li.startAlignmentGraphCreation()
    got = g1.edges()
    assert np.array_equal(got, oracle.supplement(rs, z["before"], **sp))
    assert np.array_equal(got, z["after"])  # the unmodified reference's graph after main.cpp:346
    assert li.timing["n_pairs_verified"] > 0 and li.timing["kernel_launches"] >= 8


def test_supplement_without_dead_ends_is_identity(gpu):
    """A graph whose every node has in- and out-edges (or neither) takes no part: only the final row dedupe acts."""
    rs, lmin, rsmin, sp = supplement_case("sup_cfg3")
    n = rs.n
    ring = np.stack([np.arange(n), (np.arange(n) + 1) % n, np.full(n, 5)], axis=1).astype(np.int32)
    g1 = GraphCreatorLI(rs, _graph_from_edges(n, ring), **sp).startAlignmentGraphCreation()
    assert np.array_equal(g1.edges(), ring)

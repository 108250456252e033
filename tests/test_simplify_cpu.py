"""CPU test of the oracle's first simplifier step (oracle/simplify_oracle.c) against the edges that survive the unmodified
reference's Graph::sortEdgesByIncreasingOffset + GraphSimplifier::cutNonAndWeaklyMetricTriangles (tests/golden/tri_*.npz)."""
import os

import numpy as np
import pytest

from oracle import harness, oracle
from tests.cases import TRIANGLE_CASES, triangle_case

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", list(TRIANGLE_CASES))
def test_cut_triangles_oracle_matches_reference_fixture(name):
    e, n, mx = triangle_case(name, GOLD)
    g = np.load(os.path.join(GOLD, f"{name}.npz"))
    assert int(g["n_in"]) == e.shape[0], "input fixture changed: rerun tests/golden/make_golden.py"
    got = oracle.cut_triangles(e, n, mx)
    assert np.array_equal(harness.sort_edges(got), g["edges"])
    # rows come back ordered by (offset, neighbour), as Graph::sortEdgesByIncreasingOffset leaves them
    if got.shape[0] > 1:
        same = got[1:, 0] == got[:-1, 0]
        assert np.all((got[1:, 2] > got[:-1, 2]) | ((got[1:, 2] == got[:-1, 2]) & (got[1:, 1] >= got[:-1, 1])) | ~same)


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref not built")
def test_cut_triangles_oracle_matches_reference_live():
    e, n, _ = triangle_case("tri_periodic", GOLD)
    for mx in (0, 33, 1000):
        assert np.array_equal(harness.sort_edges(oracle.cut_triangles(e, n, mx)), harness.run_cut_triangles(e, n, mx))


def test_cut_triangles_small_cases():
    # i -> a -> b of length 5 + 7 equals the direct edge i -> b of offset 12: removed; 13 is kept; beyond max_offset kept
    e = np.array([[0, 1, 5], [0, 2, 12], [1, 2, 7], [3, 4, 5], [3, 5, 13], [4, 5, 7]], np.int32)
    out = oracle.cut_triangles(e, 6, 250)
    assert out.tolist() == [[0, 1, 5], [1, 2, 7], [3, 4, 5], [3, 5, 13], [4, 5, 7]]
    assert oracle.cut_triangles(e, 6, 11).shape[0] == 6
    # a shorter two-hop path does not remove the edge (the reference tests equality, GraphSimplifier.cpp:309)
    e2 = np.array([[0, 1, 2], [0, 2, 12], [1, 2, 7]], np.int32)
    assert oracle.cut_triangles(e2, 3, 250).shape[0] == 3
    assert oracle.cut_triangles(np.zeros((0, 3), np.int32), 0, 250).shape[0] == 0


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref not built")
def test_cut_triangles_oracle_matches_reference_on_random_graphs():
    """Dense random graphs (many triangles, equal and unequal two-hop lengths, self loops, parallel entries with different offsets)."""
    rng = np.random.default_rng(31)
    for trial in range(12):
        n = int(rng.integers(5, 80))
        e = rng.integers(0, n, size=(int(rng.integers(1, 6 * n)), 2))
        w = rng.integers(1, 25, size=(e.shape[0], 1))
        edges = np.concatenate([e, w], axis=1).astype(np.int32)
        if trial % 3:
            edges = np.unique(edges, axis=0)  # otherwise: exact duplicates stay in
        mx = int(rng.integers(5, 60))
        want = harness.run_cut_triangles(edges, n, mx)
        got = harness.sort_edges(oracle.cut_triangles(edges, n, mx))
        assert np.array_equal(got, want), f"trial {trial}"

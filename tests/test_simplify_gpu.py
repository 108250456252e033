"""GPU parity of alga_gpu_cut_triangles (Graph::sortEdgesByIncreasingOffset + GraphSimplifier::cutNonAndWeaklyMetricTriangles)
through the C ABI against the oracle and the unmodified reference's results (tests/golden/tri_*.npz)."""
import os

import numpy as np
import pytest

from alga_b200.graph_creator import Graph, GraphSimplifier
from oracle import harness, oracle
from tests.cases import TRIANGLE_CASES, triangle_case

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def graph_of(edges, n) -> Graph:
    e = harness.sort_edges(edges)
    row_off = np.zeros(n + 1, np.uint64)
    if e.shape[0]:
        np.cumsum(np.bincount(e[:, 0], minlength=n), out=row_off[1:])
    return Graph(n, row_off, np.ascontiguousarray(e[:, 1]), np.ascontiguousarray(e[:, 2]))


@pytest.mark.parametrize("name", list(TRIANGLE_CASES))
def test_cut_triangles_matches_oracle_and_reference(gpu, name):
    e, n, mx = triangle_case(name, GOLD)
    gs = GraphSimplifier(graph_of(e, n), mx)
    out = gs.cutNonAndWeaklyMetricTriangles()
    got = out.edges()
    want = oracle.cut_triangles(e, n, mx)
    assert np.array_equal(got, want)  # same edges AND the same row order: (offset, neighbour)
    assert np.array_equal(harness.sort_edges(got), np.load(os.path.join(GOLD, f"{name}.npz"))["edges"])


def test_cut_triangles_thresholds_and_long_rows(gpu):
    e, n, _ = triangle_case("tri_periodic", GOLD)
    for mx in (0, 33, 1000):
        got = GraphSimplifier(graph_of(e, n), mx).cutNonAndWeaklyMetricTriangles().edges()
        assert np.array_equal(got, oracle.cut_triangles(e, n, mx))
    # a hub with a row of 200 entries (more than the 32 the in-register row sort takes) and duplicated targets
    rng = np.random.default_rng(3)
    rows = [(0, int(t), int(rng.integers(1, 60))) for t in rng.integers(1, 120, size=200)]
    rows += [(int(a), int(b), int(rng.integers(1, 60))) for a, b in rng.integers(1, 120, size=(600, 2))]
    e2 = harness.sort_edges(np.array(rows, np.int32))
    got = GraphSimplifier(graph_of(e2, 120), 250).cutNonAndWeaklyMetricTriangles().edges()
    assert np.array_equal(got, oracle.cut_triangles(e2, 120, 250))


def test_cut_triangles_takes_rows_in_any_order(gpu):
    """The output of a call is sorted by (offset, neighbour); fed back in -- or a reference graph after
    sortEdgesByIncreasingOffset -- it must give what the oracle gives for the same edge set (the rows are re-sorted on the device)."""
    e, n, mx = triangle_case("tri_periodic", GOLD)
    first = GraphSimplifier(graph_of(e, n), mx).cutNonAndWeaklyMetricTriangles()
    e1 = first.edges()
    again = GraphSimplifier(first, mx).cutNonAndWeaklyMetricTriangles().edges()  # rows by (offset, neighbour) going in
    assert np.array_equal(again, oracle.cut_triangles(harness.sort_edges(e1), n, mx))
    # rows shuffled arbitrarily
    rng = np.random.default_rng(5)
    g = graph_of(e, n)
    nbr, off = g.nbr.copy(), g.off.copy()
    for i in range(n):
        s, t = int(g.row_off[i]), int(g.row_off[i + 1])
        if t - s > 1:
            p = rng.permutation(t - s)
            nbr[s:t], off[s:t] = nbr[s:t][p], off[s:t][p]
    got = GraphSimplifier(Graph(n, g.row_off, nbr, off), mx).cutNonAndWeaklyMetricTriangles().edges()
    assert np.array_equal(got, oracle.cut_triangles(e, n, mx))

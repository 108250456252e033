"""The zero-code-change boundary: a graph written by alga_b200.graph_file is what the UNMODIFIED binary loads with
--serialize=1 instead of building its own -- same bytes as the file the stock binary writes, and the same contigs afterwards."""
import os
import subprocess

import numpy as np
import pytest

from alga_b200.graph_creator import Graph
from alga_b200.graph_file import graph_file_name, read_graph, test_name, write_graph
from oracle import harness
from tests.cases import front_case
from tests.test_input_cpu import oracle_front

test_name.__test__ = False  # a helper of the package, not a test


def graph_of(edges, n) -> Graph:
    e = harness.sort_edges(edges)
    row_off = np.zeros(n + 1, np.uint64)
    if e.shape[0]:
        np.cumsum(np.bincount(e[:, 0], minlength=n), out=row_off[1:])
    return Graph(n, row_off, np.ascontiguousarray(e[:, 1]), np.ascontiguousarray(e[:, 2]))


def test_names_follow_the_driver():
    assert test_name("/data/run7/x_1.fasta") == "ALGA_x_1_scale55_noN"
    assert graph_file_name("reads.fq") == "ALGA_reads_scale55_noN_beforeSimplifier.graph"


def test_round_trip(tmp_path):
    rng = np.random.default_rng(4)
    e = np.unique(rng.integers(0, 50, size=(300, 3)).astype(np.int32), axis=0)
    g = graph_of(e, 50)
    write_graph(str(tmp_path / "g.graph"), g)
    back = read_graph(str(tmp_path / "g.graph"))
    assert np.array_equal(back.edges(), g.edges())
    n, e2 = harness.read_graph_file(str(tmp_path / "g.graph"))
    assert n == 50 and np.array_equal(e2, g.edges())
    write_graph(str(tmp_path / "empty.graph"), graph_of(np.zeros((0, 3), np.int32), 0))
    assert read_graph(str(tmp_path / "empty.graph")).n == 0


@pytest.mark.skipif(not os.path.isfile(harness.STOCK), reason="oracle/_ref/ALGA not built")
def test_stock_binary_loads_our_file(tmp_path):
    t1, t2, ft = front_case("front_pe")
    rs, _, edges = oracle_front(t1, t2, ft)  # CPU stand-in for the GPU build: the same graph (tests/test_input_gpu.py)
    runs = {}
    for name in ("stock", "injected"):
        d = tmp_path / name
        d.mkdir()
        (d / "x_1.fasta").write_bytes(t1)
        (d / "x_2.fasta").write_bytes(t2)
        if name == "injected":
            write_graph(str(d / graph_file_name("x_1.fasta")), graph_of(edges, rs.n))
        r = subprocess.run([harness.STOCK, "--file1=x_1.fasta", "--file2=x_2.fasta", "--threads=1", "--output=contigs.fasta",
                            "--serialize=1"], cwd=d, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        runs[name] = (r.stdout, (d / "contigs.fasta").read_bytes())
    # the stock run wrote its own graph file: ours must be the same bytes
    own = (tmp_path / "stock" / graph_file_name("x_1.fasta")).read_bytes()
    assert own == (tmp_path / "injected" / graph_file_name("x_1.fasta")).read_bytes()
    assert "Creating GraphCreator" in runs["stock"][0] and "Creating GraphCreator" not in runs["injected"][0]
    assert runs["stock"][1] == runs["injected"][1] and len(runs["stock"][1]) > 0

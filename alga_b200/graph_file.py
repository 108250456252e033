"""The reference's serialised graph (``Graph::serializeGraph`` / ``deserializeGraph``, Graph.cpp:220-297) -- the zero-code-change
boundary of SURVEY.md 8-b2: the unmodified binary run with ``--serialize=1`` loads ``<TEST_NAME>_beforeSimplifier.graph``
instead of building the graph (main.cpp:242) and goes on with its own supplement, simplifier and contig stages.

Format: little-endian ``u32 n``, then for every node ``i32 id, i32 degree, degree x (i32 neighbour, i32 offset)``.
"""
from __future__ import annotations

import os

import numpy as np

from .graph_creator import Graph


def test_name(file1: str, scale: float = 0.55, remove_reads_with_n: bool = True) -> str:
    """Params::TEST_NAME as the driver builds it (Params.cpp:343, 554-557): ``ALGA_<file1 without directory and extension>
    _scale<100 * SCALE>_noN``."""
    base = os.path.basename(file1)
    stem = base[: base.rfind(".")] if "." in base else base
    return f"ALGA_{stem}_scale{int(np.float32(100) * np.float32(scale))}" + ("_noN" if remove_reads_with_n else "_randN")


def graph_file_name(file1: str, **kw) -> str:
    return test_name(file1, **kw) + "_beforeSimplifier.graph"


def write_graph(path: str, graph: Graph) -> None:
    n = int(graph.n)
    deg = np.diff(graph.row_off.astype(np.int64)).astype(np.int64)
    total = 1 + 2 * n + 2 * int(graph.n_edges)
    out = np.empty(total, dtype="<i4")
    out[0] = n  # u32 in the file; node counts are < 2^31
    head = 1 + 2 * np.arange(n, dtype=np.int64) + 2 * graph.row_off[:-1].astype(np.int64)  # position of node i's record
    out[head] = np.arange(n, dtype=np.int32)
    out[head + 1] = deg.astype(np.int32)
    if graph.n_edges:
        row = np.repeat(np.arange(n, dtype=np.int64), deg)
        k = np.arange(int(graph.n_edges), dtype=np.int64)
        pos = head[row] + 2 + 2 * (k - graph.row_off[:-1].astype(np.int64)[row])
        out[pos] = graph.nbr
        out[pos + 1] = graph.off
    out.tofile(path)


def read_graph(path: str) -> Graph:
    raw = np.fromfile(path, dtype="<i4")
    n = int(raw[0])
    row_off = np.zeros(n + 1, np.uint64)
    nbr_parts, off_parts = [None] * n, [None] * n
    p = 1
    for _ in range(n):
        v, d = int(raw[p]), int(raw[p + 1])
        e = raw[p + 2 : p + 2 + 2 * d].reshape(-1, 2)
        nbr_parts[v], off_parts[v] = e[:, 0], e[:, 1]
        row_off[v + 1] = d
        p += 2 + 2 * d
    np.cumsum(row_off[1:], out=row_off[1:])
    nbr = np.concatenate(nbr_parts).astype(np.int32) if n else np.zeros(0, np.int32)
    off = np.concatenate(off_parts).astype(np.int32) if n else np.zeros(0, np.int32)
    return Graph(n, row_off, nbr, off)

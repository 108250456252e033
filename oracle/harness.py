"""TEST INFRASTRUCTURE ONLY -- runs the compiled, unmodified reference (``oracle/_ref``).

``oracle/_ref/alga_ref_harness`` is ``oracle/ref_harness.cpp`` linked against the reference's own
sources (built by ``oracle/Makefile`` in the dev container; the binary travels to the GPU box).
"""
from __future__ import annotations

import json
import os
import struct
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
HARNESS = os.path.join(HERE, "_ref", "alga_ref_harness")


def available() -> bool:
    return os.path.isfile(HARNESS) and os.access(HARNESS, os.X_OK)


def write_reads(path, reads, min_overlap, rs_min_overlap, min_offset=0):
    """ALGR: magic, u32 n, i32 x4 (min_overlap, rs, min_offset, 0), len, from, to, word_off, words."""
    with open(path, "wb") as f:
        f.write(b"ALGR")
        f.write(struct.pack("<I4i", reads.n, min_overlap, rs_min_overlap, min_offset, 0))
        f.write(reads.len_nt.astype("<u4").tobytes())
        f.write(reads.align_from.astype("u1").tobytes())
        f.write(reads.align_to.astype("u1").tobytes())
        f.write(reads.word_off.astype("<u8").tobytes())
        f.write(reads.words.astype("<u4").tobytes())


def read_edges(path) -> np.ndarray:
    """ALGE: magic, u32 n, u64 E, E x (i32 src, i32 dst, i32 offset) -> (E, 3) int32 sorted."""
    with open(path, "rb") as f:
        assert f.read(4) == b"ALGE"
        _n, e = struct.unpack("<IQ", f.read(12))
        arr = np.frombuffer(f.read(12 * e), dtype="<i4").reshape(-1, 3).copy()
    return sort_edges(arr)


def sort_edges(e: np.ndarray) -> np.ndarray:
    e = np.asarray(e, dtype=np.int32).reshape(-1, 3)
    if e.shape[0] == 0:
        return e
    order = np.lexsort((e[:, 2], e[:, 1], e[:, 0]))
    return np.ascontiguousarray(e[order])


def run_prefsuf(reads, min_overlap, rs_min_overlap, min_offset=0, threads=1, want_edges=True):
    """Run the reference's GraphCreatorPrefSuf + retainOnlySmallestOffset; returns (edges, info)."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        rp = os.path.join(d, "in.algr")
        ep = os.path.join(d, "out.alge")
        write_reads(rp, reads, min_overlap, rs_min_overlap, min_offset)
        out = subprocess.run([HARNESS, "prefsuf", rp, ep if want_edges else "-", str(threads)], check=True,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, cwd=d)
        info = json.loads(out.stdout.decode().strip().splitlines()[-1])
        edges = read_edges(ep) if want_edges else None
    return edges, info


def run_verify(reads, pairs, thr, max_offset_pct, min_overlap_area, min_offset=0, lcs_rate_pct=0, lcs_band=2) -> np.ndarray:
    """Run the reference's AlignmentControllerHybrid::canAlign on (a, b, offset) triples."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    pairs = np.ascontiguousarray(pairs, dtype="<i4").reshape(-1, 3)
    with tempfile.TemporaryDirectory() as d:
        rp = os.path.join(d, "in.algr")
        pp = os.path.join(d, "pairs.algp")
        vp = os.path.join(d, "verdict.bin")
        write_reads(rp, reads, 0, 0, 0)
        with open(pp, "wb") as f:
            f.write(b"ALGP")
            f.write(struct.pack("<Q4i", pairs.shape[0], thr, max_offset_pct, min_overlap_area, min_offset))
            f.write(pairs.tobytes())
        extra = [str(lcs_rate_pct), str(lcs_band)] if lcs_rate_pct > 0 else []  # USE_ACLER_INSTEAD_OF_ACLCS = 0
        subprocess.run([HARNESS, "verify", rp, pp, vp] + extra, check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, cwd=d)
        return np.fromfile(vp, dtype=np.uint8)


def write_edges(path, n, edges):
    e = np.ascontiguousarray(edges, dtype="<i4").reshape(-1, 3)
    with open(path, "wb") as f:
        f.write(b"ALGE")
        f.write(struct.pack("<IQ", n, e.shape[0]))
        f.write(e.tobytes())


def run_supplement(reads, edges_in, threshold_pct, max_offset_pct, min_overlap_area, kmer_length_bucket, min_offset=0,
                   threads=1):
    """Run the reference's error-rate supplement (main.cpp:300-355) on a given graph; returns (edges, info)."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        rp, ip, op = (os.path.join(d, x) for x in ("in.algr", "in.alge", "out.alge"))
        write_reads(rp, reads, 0, 0, min_offset)
        write_edges(ip, reads.n, edges_in)
        out = subprocess.run([HARNESS, "supplement", rp, ip, op, str(min_overlap_area), str(max_offset_pct),
                              str(threshold_pct), str(kmer_length_bucket), str(threads)], check=True,
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, cwd=d)
        info = json.loads(out.stdout.decode().strip().splitlines()[-1])
        return read_edges(op), info


def run_prefix_reads(reads, remove_type=2, threads=1) -> np.ndarray:
    """Run the reference's ReadPreprocess::getPrefixReads; returns the uint8 removal mask."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        rp, mp = os.path.join(d, "in.algr"), os.path.join(d, "mask.bin")
        write_reads(rp, reads, 0, 0, 0)
        subprocess.run([HARNESS, "prefixreads", rp, mp, str(remove_type), str(threads)], check=True,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d)
        return np.fromfile(mp, dtype=np.uint8)


def read_reads(path):
    """ALGR file -> ReadSet."""
    from alga_b200.readset import ReadSet
    with open(path, "rb") as f:
        assert f.read(4) == b"ALGR"
        n, *_ = struct.unpack("<I4i", f.read(20))
        ln = np.frombuffer(f.read(4 * n), "<u4").copy()
        af = np.frombuffer(f.read(n), "u1").copy()
        at = np.frombuffer(f.read(n), "u1").copy()
        off = np.frombuffer(f.read(8 * (n + 1)), "<u8").copy()
        w = np.frombuffer(f.read(4 * int(off[n])), "<u4").copy()
    return ReadSet(w, off, ln, align_from=af, align_to=at)


_EXT = {0: "txt", 1: "fasta", 2: "fastq"}


def run_read_input(text1: bytes, text2: bytes | None = None, file_type=1, threads=1, extra=()):
    """Run the reference's InputReader::readInput on the given file contents; returns the ReadSet (nullptr = length 0)."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        f1 = os.path.join(d, "x_1." + _EXT[file_type])
        f2 = os.path.join(d, "x_2." + _EXT[file_type])
        op = os.path.join(d, "out.algr")
        open(f1, "wb").write(text1)
        if text2 is not None:
            open(f2, "wb").write(text2)
        r = subprocess.run([HARNESS, "readinput", f1, f2 if text2 is not None else "-", op, str(threads), *extra],
                           stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d)
        if r.returncode != 0:  # the reference exits on a bad character (InputReader.cpp:324-327) / asserts
            raise RuntimeError(f"the reference's reader exited with status {r.returncode}")
        return read_reads(op)


STOCK = os.path.join(HERE, "_ref", "ALGA")


def read_graph_file(path) -> np.ndarray:
    """Graph::serializeGraph format (Graph.cpp:269-297) -> (n, (E, 3) int32 sorted edges)."""
    raw = np.fromfile(path, dtype="<i4")
    n = int(raw[0])
    edges = []
    p = 1
    for _ in range(n):
        v, deg = int(raw[p]), int(raw[p + 1])
        p += 2
        if deg:
            e = raw[p : p + 2 * deg].reshape(-1, 2)
            edges.append(np.column_stack([np.full(deg, v, np.int32), e]))
            p += 2 * deg
    e = np.concatenate(edges) if edges else np.zeros((0, 3), np.int32)
    return n, sort_edges(e)


def run_stock_graph(text1: bytes, text2: bytes | None = None, file_type=1, threads=1, extra=()):
    """The stock ALGA binary (main.cpp:57-293) with --serialize=1 on the given files: (n_nodes, edges) of the graph it
    writes after GraphCreatorPrefSuf + retainOnlySmallestOffset, i.e. on the ids left by the reference's own reader,
    duplicate removal and renumbering."""
    if not os.path.isfile(STOCK):
        raise RuntimeError("oracle/_ref/ALGA is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        f1 = "x_1." + _EXT[file_type]
        args = [STOCK, "--file1=" + f1]
        open(os.path.join(d, f1), "wb").write(text1)
        if text2 is not None:
            f2 = "x_2." + _EXT[file_type]
            open(os.path.join(d, f2), "wb").write(text2)
            args.append("--file2=" + f2)
        args += [f"--threads={threads}", "--output=contigs.fasta", "--serialize=1", *extra]
        subprocess.run(args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d)
        g = [x for x in os.listdir(d) if x.endswith("_beforeSimplifier.graph")]
        assert len(g) == 1, g
        return read_graph_file(os.path.join(d, g[0]))


def run_cut_triangles(edges_in, n_nodes: int, max_offset: int, threads=1) -> np.ndarray:
    """The reference's Graph::sortEdgesByIncreasingOffset + GraphSimplifier::cutNonAndWeaklyMetricTriangles on a graph;
    returns the surviving edges sorted by (src, dst, offset)."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        ip, op = os.path.join(d, "in.alge"), os.path.join(d, "out.alge")
        write_edges(ip, n_nodes, edges_in)
        subprocess.run([HARNESS, "triangles", ip, op, str(max_offset), str(threads)], check=True, stdout=subprocess.DEVNULL,
                       stderr=subprocess.DEVNULL, cwd=d)
        return read_edges(op)


def run_driver_paired_offsets(text1: bytes, text2: bytes | None = None, file_type=1, threads=1) -> np.ndarray:
    """Run the reference's whole driver (its main(), compiled as a function) on the given files and return
    Global::pairedReadOffset as the renumbering of main.cpp:150-232 left it."""
    if not available():
        raise RuntimeError("oracle/_ref/alga_ref_harness is not built (make -C oracle ref)")
    with tempfile.TemporaryDirectory() as d:
        f1 = "x_1." + _EXT[file_type]
        open(os.path.join(d, f1), "wb").write(text1)
        args = ["--file1=" + f1]
        if text2 is not None:
            f2 = "x_2." + _EXT[file_type]
            open(os.path.join(d, f2), "wb").write(text2)
            args.append("--file2=" + f2)
        args += [f"--threads={threads}", "--output=contigs.fasta"]
        subprocess.run([HARNESS, "driver", "po.bin"] + args, check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, cwd=d)
        raw = np.fromfile(os.path.join(d, "po.bin"), dtype=np.uint8)
        n = int(raw[:4].view("<u4")[0])
        return raw[4 : 4 + n].copy()

"""Host-side mirror of the reference's input stage: ``InputReader`` (include/IO/InputReader.h, src/IO/InputReader.cpp),
the renumbering of ``main.cpp:150-232`` and, on top of them, the part of the driver that leads from the input files to
the overlap graph (``main.cpp:82-291``).  All compute goes through the C ABI of ``libalga_gpu.so``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .graph_creator import Graph, GraphCreatorPrefSuf, ReadPreprocess, _csr_to_graph, _PinnedArray
from .readset import ReadSet

PLAIN, FASTA, FASTQ = _lib.INPUT_PLAIN, _lib.INPUT_FASTA, _lib.INPUT_FASTQ


def file_type_of(path: str) -> int:
    """Params.cpp:332-335: the extension picks the parser (pfasta with paired reads is read like fasta)."""
    ext = path.rsplit("/", 1)[-1].rsplit(".", 1)[-1] if "." in path.rsplit("/", 1)[-1] else ""
    return FASTA if ext in ("fasta", "pfasta") else (FASTQ if ext in ("fastq", "fq") else PLAIN)


def _timing(tm: _lib.Timing) -> dict:
    return {"h2d_ms": tm.h2d_ms, "device_ms": tm.device_ms, "d2h_ms": tm.d2h_ms, "total_ms": tm.total_ms,
            "kernel_launches": tm.kernel_launches}


def _take(out: _lib.ReadSetOut):
    """Copy a library-allocated read set into numpy arrays and release it."""
    n, stride = out.n_reads, out.stride_words
    words = np.ctypeslib.as_array(out.words, shape=(max(n * stride, 1),))[: n * stride].copy()
    ln = np.ctypeslib.as_array(out.len_nt, shape=(max(n, 1),))[:n].copy()
    old = np.ctypeslib.as_array(out.old_id, shape=(max(n, 1),))[:n].copy() if out.old_id else None
    po = np.ctypeslib.as_array(out.paired_offset, shape=(max(n, 1),))[:n].copy() if out.paired_offset else None
    info = {"n_records": (out.n_records[0], out.n_records[1]), "n_with_n": out.n_with_n, "n_str": out.n_str,
            "max_len_nt": out.max_len_nt, "stride_words": stride}
    _lib.load().alga_gpu_free_read_set(C.byref(out))
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(stride)
    return ReadSet(words, off, ln), old, po, info


class InputReader:
    """``InputReader::readInput`` (InputReader.cpp:44-139): file contents in, ``Global::READS`` out -- both strands of
    every record, reverse complement at the even id, mates interleaved, removed reads (N, short-period repeats) as
    length 0."""

    def __init__(self, file_type: int = FASTA, trim_left: int = 3, trim_right: int = 3, rna: bool = False,
                 str_threshold: int = 20, device: int = 0):
        self.params = _lib.InputParams(file_type, trim_left, trim_right, int(rna), str_threshold, device)
        self.timing: dict | None = None
        self.info: dict | None = None

    def readInput(self, text1: bytes, text2: bytes | None = None) -> ReadSet:
        lib = _lib.load()
        b1 = np.frombuffer(text1, np.uint8)
        b2 = np.frombuffer(text2, np.uint8) if text2 is not None else None
        keep = C.create_string_buffer(1)
        p1 = b1.ctypes.data if b1.size else C.addressof(keep)
        p2 = (b2.ctypes.data if b2.size else C.addressof(keep)) if b2 is not None else None
        out, tm = _lib.ReadSetOut(), _lib.Timing()
        _lib.check(lib.alga_gpu_read_input(p1, b1.size, p2, b2.size if b2 is not None else 0, C.byref(self.params),
                                           C.byref(out), C.byref(tm)))
        self.timing = _timing(tm)
        rs, _, _, self.info = _take(out)
        return rs


@dataclass
class Remapped:
    reads: ReadSet
    old_id: np.ndarray  # previous id of every read
    paired_offset: np.ndarray  # Global::pairedReadOffset
    timing: dict


def remap_reads(reads: ReadSet, remove_mask: np.ndarray | None = None, device: int = 0) -> Remapped:
    """main.cpp:133-140 + 150-232: drop the marked reads and the nullptr reads, renumber the rest in order."""
    from .graph_creator import _reads_struct
    lib = _lib.load()
    st = _reads_struct(reads)
    m = np.ascontiguousarray(remove_mask, dtype=np.uint8) if remove_mask is not None else None
    out, tm = _lib.ReadSetOut(), _lib.Timing()
    _lib.check(lib.alga_gpu_remap_reads(C.byref(st), m.ctypes.data if m is not None else None, device, C.byref(out),
                                        C.byref(tm)))
    t = _timing(tm)
    rs, old, po, _ = _take(out)
    return Remapped(rs, old, po, t)


def driver_params(reads: ReadSet, trim_left: int = 3, trim_right: int = 3, scale: float = 0.55) -> dict:
    """main.cpp:93-110 with the float arithmetic of the reference: LEN = int(avg length) + trims, L = int(LEN * SCALE),
    RSOEMO = int(LEN * (SCALE + 1) / 2); LI_KMER_LENGTH = min(2L/3, 60)."""
    alive = reads.len_nt[reads.len_nt > 0]
    avg = float(alive.astype(np.float64).sum() / max(alive.shape[0], 1))
    LEN = int(avg) + trim_left + trim_right
    L = int(np.float32(LEN) * np.float32(scale))
    RS = int(np.float32(LEN) * (np.float32(scale) + np.float32(1)) / np.float32(2))
    return {"min_overlap": L, "rs_min_overlap": RS, "li_kmer_length": min(2 * L // 3, 60), "li_kmer_intervals": 3}


@dataclass
class OverlapGraph:
    reads: ReadSet  # the renumbered read set the graph is built on (Global::READS at main.cpp:237)
    graph: Graph  # Global::GRAPH after main.cpp:291
    paired_offset: np.ndarray
    params: dict
    timing: dict
    old_id: np.ndarray | None = None  # id of every read before the renumbering (i.e. in the reader's output)


class PinnedText:
    """The bytes of an input file in page-locked memory (``alga_gpu_host_alloc``) -- where a caller that cares about the
    upload rate reads its file to."""

    def __init__(self, data: bytes):
        src = np.frombuffer(data, np.uint8)
        self.size = int(src.size)
        self._pin = _PinnedArray(src if src.size else np.zeros(1, np.uint8))
        self.array = self._pin.array


def _text_ptr(t, keep):
    if t is None:
        return None, 0
    if isinstance(t, PinnedText):
        return t.array.ctypes.data, t.size
    a = np.frombuffer(t, np.uint8)
    return (a.ctypes.data if a.size else C.addressof(keep)), int(a.size)


def build_overlap_graph(text1, text2=None, file_type: int = FASTA, device: int = 0, remove_type: int = 2,
                        min_overlap: int = 0, rs_min_overlap: int = 0) -> OverlapGraph:
    """The reference driver from the input files to the overlap graph (main.cpp:82-291) in ONE call of the C ABI
    (``alga_gpu_files_to_graph``): reader, parameter derivation, ReadPreprocess::getPrefixReads + removal, renumbering,
    the short-read rule of main.cpp:253-266, GraphCreatorPrefSuf + retainOnlySmallestOffset -- all on the GPU, the read
    set device-resident between the stages.  ``text1`` / ``text2``: ``bytes`` or ``PinnedText``."""
    lib = _lib.load()
    keep = C.create_string_buffer(1)
    p1, n1 = _text_ptr(text1, keep)
    p2, n2 = _text_ptr(text2, keep)
    dp = _lib.DriverParams(_lib.InputParams(file_type, 3, 3, 0, 20, device), remove_type, 0.55, min_overlap, rs_min_overlap)
    out, tm = _lib.OverlapGraphOut(), _lib.Timing()
    _lib.check(lib.alga_gpu_files_to_graph(p1, n1, p2, n2, C.byref(dp), C.byref(out), C.byref(tm)))
    try:
        graph = _csr_to_graph(out.graph)
    finally:
        lib.alga_gpu_free_csr(C.byref(out.graph))
    params = {"min_overlap": out.min_overlap, "rs_min_overlap": out.rs_min_overlap, "li_kmer_length": out.li_kmer_length,
              "li_kmer_intervals": 3, "avg_len": out.avg_len, "n_reads_in": out.n_reads_in,
              "n_records": (out.n_records[0], out.n_records[1]), "n_with_n": out.n_with_n, "n_str": out.n_str}
    timing = _timing(tm)
    # read_input: wall time from the call to "packed reads resident" (uploads included, they overlap the scan of file 1)
    timing["stage_ms"] = dict(zip(("read_input", "prefix_reads", "remap", "prefsuf_device", "prefsuf_host"), list(tm.stage_ms)[:5]))
    rs, old, po, _ = _take(out.reads)
    return OverlapGraph(rs, graph, po, params, timing, old)


def build_overlap_graph_staged(text1: bytes, text2: bytes | None = None, file_type: int = FASTA, device: int = 0,
                               remove_type: int = 2) -> OverlapGraph:
    """The same path as ``build_overlap_graph`` composed from the separate entry points (host buffers between the
    stages) -- the form a driver that keeps its own ``vector<Read*>`` would use; also the cross-check of the fused call."""
    reader = InputReader(file_type, device=device)
    rs = reader.readInput(text1, text2)
    timing = {"read_input": reader.timing}
    params = driver_params(rs)
    pre = ReadPreprocess(rs, device)
    mask = pre.getPrefixReads(remove_type)
    timing["prefix_reads"] = pre.timing
    rm = remap_reads(rs, mask, device)
    timing["remap"] = rm.timing
    # main.cpp:253-266: reads shorter than LI_KMER_INTERVALS + LI_KMER_LENGTH take no part and are removed (nullptr)
    ln = rm.reads.len_nt.copy()
    ln[ln < params["li_kmer_intervals"] + params["li_kmer_length"]] = 0
    final = ReadSet(rm.reads.words, rm.reads.word_off, ln)
    gc = GraphCreatorPrefSuf(final, params["min_overlap"], params["rs_min_overlap"], device=device)
    graph = gc.startAlignmentGraphCreation()
    timing["prefsuf"] = gc.timing
    return OverlapGraph(final, graph, rm.paired_offset, params, timing, rm.old_id)

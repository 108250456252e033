"""Host-side mirror of the reference's graph-creator interface over the C ABI.

``GraphCreatorPrefSuf`` follows the reference class of the same name
(``include/GraphCreators/GraphCreator.h:12-62``, ``include/GraphCreators/GraphCreatorPrefSuf.h:22-28``):
constructed from the read set, ``setAlignFrom`` / ``setAlignTo`` / ``getAlignFrom`` / ``getAlignTo``,
``startAlignmentGraphCreation()``, ``clear()``.  The result is what ``Graph::V`` holds after
``main.cpp:282-291`` (graph creation + ``retainOnlySmallestOffset``), as a CSR.

All compute happens in ``libalga_gpu.so`` (hand-written CUDA, sm_100a).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from .readset import ReadSet


@dataclass
class Graph:
    """Forward adjacency: row b holds (nbr, off) = "read nbr starts at position off of read b"."""

    n: int
    row_off: np.ndarray  # uint64, n+1
    nbr: np.ndarray  # int32, E
    off: np.ndarray  # int32, E

    @property
    def n_edges(self) -> int:
        return int(self.nbr.shape[0])

    def edges(self) -> np.ndarray:
        """(E, 3) int32 array of (src, dst, offset) sorted by (src, dst)."""
        deg = np.diff(self.row_off.astype(np.int64))
        src = np.repeat(np.arange(self.n, dtype=np.int32), deg)
        return np.stack([src, self.nbr, self.off], axis=1).astype(np.int32)

    def neighbors(self, b: int):
        s, e = int(self.row_off[b]), int(self.row_off[b + 1])
        return list(zip(self.nbr[s:e].tolist(), self.off[s:e].tolist()))


def fixed_stride(reads: ReadSet) -> int:
    """Words per read if every read starts at a multiple of the same word count (else 0)."""
    n = reads.n
    w = int(reads.word_off[1] - reads.word_off[0]) if n else 0
    if w > 0 and np.array_equal(reads.word_off, np.arange(n + 1, dtype=np.uint64) * np.uint64(w)):
        return w
    return 0


def _reads_struct(reads: ReadSet, stride: int | None = None) -> _lib.Reads:
    """``stride``: result of ``fixed_stride`` if the caller already knows it (it costs a pass over word_off)."""
    if stride is None:
        stride = fixed_stride(reads)
    # fixed-stride layout lets the library skip the offset array (alga_gpu.h: word_off == NULL)
    word_off = None if stride else reads.word_off.ctypes.data
    return _lib.Reads(reads.n, reads.words.ctypes.data, word_off, stride, reads.len_nt.ctypes.data,
                      reads.align_from.ctypes.data, reads.align_to.ctypes.data)


def _csr_to_graph(csr: _lib.Csr, borrow: bool = False) -> Graph:
    """Results are COPIED out of the library's buffers by default.  ``borrow=True`` wraps a borrowed (page-locked,
    library-owned) result without a copy: such a Graph aliases process-wide staging memory and is valid only until the next
    build / supplement / triangle cut in this process (exactly like the ``alga_csr`` it came from, include/alga_gpu.h)."""
    n, e = csr.n_reads, csr.n_edges
    keep = (lambda a: a) if (csr.borrowed and borrow) else (lambda a: a.copy())
    row_off = keep(np.ctypeslib.as_array(csr.row_off, shape=(n + 1,)))
    if e:
        nbr = keep(np.ctypeslib.as_array(csr.nbr, shape=(e,)))
        off = keep(np.ctypeslib.as_array(csr.off, shape=(e,)))
    else:
        nbr = np.zeros(0, np.int32)
        off = np.zeros(0, np.int32)
    return Graph(n, row_off, nbr, off)


class _PinnedArray:
    """numpy array over page-locked host memory from ``alga_gpu_host_alloc`` (freed with the object)."""

    def __init__(self, src: np.ndarray):
        self._lib = _lib.load()
        nbytes = max(int(src.nbytes), 1)
        self._ptr = self._lib.alga_gpu_host_alloc(nbytes)
        if not self._ptr:
            raise MemoryError(self._lib.alga_gpu_last_error().decode(errors="replace"))
        buf = (C.c_uint8 * nbytes).from_address(self._ptr)
        self.array = np.frombuffer(buf, dtype=src.dtype, count=src.size).reshape(src.shape)
        self.array[...] = src

    def __del__(self):
        try:
            if self._ptr:
                self._lib.alga_gpu_host_free(self._ptr)
                self._ptr = None
        except Exception:
            pass


class GraphCreatorPrefSuf:
    """Drop-in for the reference's ``GraphCreatorPrefSuf`` (GraphCreatorPrefSuf.cpp:15-126).

    ``pinned=True`` stages the packed reads in page-locked host memory (``alga_gpu_host_alloc``), which is what
    the C++ shim does when it gathers ``vector<Read*>``; the upload then runs at full host->device rate."""

    def __init__(self, reads: ReadSet, min_overlap: int, rs_min_overlap: int, min_offset: int = 0,
                 max_len_cap: int = 500, device: int = 0, list_cap: int = 0, pinned: bool = False,
                 force_generic: bool = False, borrow: bool = False, n_gpus: int = 1):
        self.n_gpus = n_gpus  # > 1: alga_gpu_prefsuf_build_multi (one process, up to n_gpus GPUs of the box)
        self.borrow = borrow  # True: results alias the library's page-locked staging until the next build (_csr_to_graph)
        self.params = _lib.PsParams(min_overlap, rs_min_overlap, min_offset, max_len_cap, device, list_cap,
                                    _lib.PS_FORCE_GENERIC if force_generic else 0)
        self._pins = []
        if pinned:
            def pin(a):
                self._pins.append(_PinnedArray(a))
                return self._pins[-1].array
            reads = ReadSet(pin(reads.words), pin(reads.word_off), pin(reads.len_nt), reads.align_from, reads.align_to)
            # GraphCreator::GraphCreator (GraphCreator.cpp:9-17): flags start as true for every read
            self.alignFrom = pin(reads.align_from)
            self.alignTo = pin(reads.align_to)
        else:
            self.alignFrom = reads.align_from.copy()
            self.alignTo = reads.align_to.copy()
        self.reads = reads
        self._stride = fixed_stride(reads)
        self.graph: Graph | None = None
        self.timing: dict | None = None

    # GraphCreator.h:22-43
    def setAlignTo(self, i: int, val: bool):
        self.alignTo[i] = 1 if val else 0

    def setAlignFrom(self, i: int, val: bool):
        self.alignFrom[i] = 1 if val else 0

    def getAlignFrom(self, i: int) -> bool:
        return bool(self.alignFrom[i])

    def getAlignTo(self, i: int) -> bool:
        return bool(self.alignTo[i])

    def startAlignmentGraphCreation(self) -> Graph:
        """GraphCreatorPrefSuf.cpp:73-126 + Graph::retainOnlySmallestOffset (main.cpp:291)."""
        lib = _lib.load()
        r = self.reads
        st = _lib.Reads(r.n, r.words.ctypes.data, None if self._stride else r.word_off.ctypes.data, self._stride,
                        r.len_nt.ctypes.data, self.alignFrom.ctypes.data, self.alignTo.ctypes.data)
        csr = _lib.Csr()
        tm = _lib.Timing()
        if self.n_gpus > 1:
            _lib.check(lib.alga_gpu_prefsuf_build_multi(C.byref(st), C.byref(self.params), self.n_gpus, C.byref(csr), C.byref(tm)))
        else:
            _lib.check(lib.alga_gpu_prefsuf_build(C.byref(st), C.byref(self.params), C.byref(csr), C.byref(tm)))
        try:
            self.graph = _csr_to_graph(csr, self.borrow)
        finally:
            lib.alga_gpu_free_csr(C.byref(csr))
        self.timing = {k: getattr(tm, k) for k, _ in _lib.Timing._fields_ if k != "stage_ms"}
        self.timing["stage_ms"] = dict(zip(("index", "phase1", "transpose", "phase2", "csr"), list(tm.stage_ms)[:5]))
        self.timing["n_row_overflow"] = int(tm.stage_ms[5])
        self.timing["n_hard_sources"] = int(tm.stage_ms[6])
        self.timing["n_gpus_used"] = int(tm.stage_ms[7])  # alga_gpu_prefsuf_build_multi only (else 0)
        return self.graph

    def clear(self):
        """GraphCreatorPrefSuf::clear (GraphCreatorPrefSuf.cpp:62-71): drop the working state."""
        self.graph = None


def fingerprints(reads: ReadSet, L: int, device: int = 0):
    """(pre64, pre32, suf64, suf32) of every read with len >= L (GraphCreatorPrefSuf.cpp:213-236)."""
    lib = _lib.load()
    n = reads.n
    p64 = np.zeros(n, np.uint64); p32 = np.zeros(n, np.uint32)
    s64 = np.zeros(n, np.uint64); s32 = np.zeros(n, np.uint32)
    st = _reads_struct(reads)
    _lib.check(lib.alga_gpu_fingerprints(C.byref(st), L, device, p64.ctypes.data, p32.ctypes.data, s64.ctypes.data,
                                         s32.ctypes.data))
    return p64, p32, s64, s32


def pack_reads(ascii_reads: np.ndarray, device: int = 0) -> np.ndarray:
    """2-bit pack an (n, len) uint8 matrix of ASCII nucleotides (Read::createSequence, Read.cpp:40-68)."""
    lib = _lib.load()
    a = np.ascontiguousarray(ascii_reads, dtype=np.uint8)
    n, ln = a.shape
    words = np.zeros((n, (ln + 15) // 16), np.uint32)
    _lib.check(lib.alga_gpu_pack_reads(a.ctypes.data, n, ln, device, words.ctypes.data))
    return words


def verify_pairs(reads: ReadSet, pairs: np.ndarray, threshold_pct: int, max_offset_pct: int, min_overlap_area: int,
                 min_offset: int = 0, same_ends: int = 3, device: int = 0, lcs_rate_pct: int = 0, lcs_band: int = 2) -> np.ndarray:
    """Batch ``AlignmentControllerHybrid::canAlign`` (AlignmentControllerHybrid.cpp:46-83).  ``lcs_rate_pct`` > 0 is
    ``Params::USE_ACLER_INSTEAD_OF_ACLCS = 0`` with that ``MINIMAL_OVERLAP_RATE_FOR_LCS``: pairs the low-error test rejects go on
    to the banded LCS of ``AlignmentControllerLCS`` (band ``lcs_band`` = ``MAX_ERROR_RATE_FOR_LCS``)."""
    lib = _lib.load()
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 3)
    out = np.zeros(pairs.shape[0], np.uint8)
    st = _reads_struct(reads)
    vp = _lib.VerifyParams(max_offset_pct, min_offset, min_overlap_area, threshold_pct, same_ends, device, lcs_rate_pct, lcs_band)
    _lib.check(lib.alga_gpu_verify_pairs(C.byref(st), pairs.ctypes.data, pairs.shape[0], C.byref(vp), out.ctypes.data))
    return out


def li_kmers(reads: ReadSet, ids: np.ndarray, priorities=(0, 1, 2, 3), kmer_length: int = 35, intervals: int = 6,
             device: int = 0):
    """``Read::getLIKmers`` (Read.cpp:145-226) of the reads ``ids``: (hash [n, intervals] uint64, ind [n, intervals] int32,
    ind = -1 for intervals beyond the last window)."""
    lib = _lib.load()
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    h = np.zeros((ids.shape[0], intervals), np.uint64)
    ind = np.zeros((ids.shape[0], intervals), np.int32)
    pr = np.ascontiguousarray(priorities, dtype=np.int32)
    st = _reads_struct(reads)
    _lib.check(lib.alga_gpu_li_kmers(C.byref(st), ids.ctypes.data, ids.shape[0], pr.ctypes.data, kmer_length, intervals,
                                     device, h.ctypes.data, ind.ctypes.data))
    return h, ind


def supplement_params(avg_len: float, error_rate_pct: int = 2, scale: float = 0.55):
    """Params of the supplement as the reference driver derives them (main.cpp:93-115, 332-340), float arithmetic as there."""
    LEN = int(avg_len) + 6
    L = int(np.float32(LEN) * np.float32(scale))
    return dict(threshold_pct=99 - error_rate_pct,
                max_offset_pct=int((np.float32(1.0) - np.float32(scale)) * np.float32(avg_len) / 2),
                min_overlap_area=int((np.float32(1.0) + np.float32(scale)) * np.float32(avg_len) / 2),
                kmer_length_bucket=min(2 * L // 3, 60))


class GraphCreatorLI:
    """Drop-in for the error-rate supplement of the reference driver (main.cpp:300-355): ``new GraphCreatorLI(READS, G)``,
    the dead-end flags of main.cpp:308-323, the Params of main.cpp:332-340 and ``startAlignmentGraphCreation()`` followed
    by ``G->retainOnlySmallestOffset()`` (main.cpp:343-346).

    ``threshold_pct`` = MINIMAL_OVERLAP_FOR_LCS_LOW_ERROR (99 - ERROR_RATE), ``max_offset_pct`` =
    MAX_OFFSET_CONSIDERED_FOR_ALIGNMENT ((1 - SCALE) * avg_len / 2), ``min_overlap_area`` = MIN_OVERLAP_AREA
    ((1 + SCALE) * avg_len / 2), ``kmer_length_bucket`` = KMER_LENGTH_BUCKET (main.cpp:104)."""

    def __init__(self, reads: ReadSet, graph: Graph, threshold_pct: int, max_offset_pct: int, min_overlap_area: int,
                 kmer_length_bucket: int, min_offset: int = 0, same_ends: int = 3, kmer_length: int = 35,
                 intervals: int = 6, device: int = 0):
        self.reads, self.graph_in = reads, graph
        self.params = _lib.SupParams(max_offset_pct, min_offset, min_overlap_area, threshold_pct, same_ends, kmer_length,
                                     intervals, kmer_length_bucket, device)
        self.graph: Graph | None = None
        self.timing: dict | None = None

    def startAlignmentGraphCreation(self) -> Graph:
        lib = _lib.load()
        g = self.graph_in
        row_off = np.ascontiguousarray(g.row_off, dtype=np.uint64)
        nbr = np.ascontiguousarray(g.nbr, dtype=np.int32)
        off = np.ascontiguousarray(g.off, dtype=np.int32)
        cin = _lib.Csr(g.n, g.n_edges, row_off.ctypes.data_as(C.POINTER(C.c_uint64)), nbr.ctypes.data_as(C.POINTER(C.c_int32)),
                       off.ctypes.data_as(C.POINTER(C.c_int32)), 1)
        cout = _lib.Csr()
        tm = _lib.Timing()
        st = _reads_struct(self.reads)
        _lib.check(lib.alga_gpu_supplement(C.byref(st), C.byref(cin), C.byref(self.params), C.byref(cout), C.byref(tm)))
        try:
            self.graph = _csr_to_graph(cout)
        finally:
            lib.alga_gpu_free_csr(C.byref(cout))
        self.timing = {"h2d_ms": tm.h2d_ms, "device_ms": tm.device_ms, "total_ms": tm.total_ms,
                       "kernel_launches": tm.kernel_launches,
                       "stage_ms": dict(zip(("li_kmers", "bucket_sort", "enumerate", "can_align", "replay"), list(tm.stage_ms)[:5])),
                       "n_dead_end_reads": int(tm.stage_ms[5]),
                       "n_pairs_verified": int(tm.stage_ms[6]), "replay_levels": int(tm.stage_ms[7])}
        return self.graph


class ReadPreprocess:
    """Mirror of the reference's ``ReadPreprocess`` (include/IO/ReadPreprocess.h:11-17) for the one method the driver
    calls before the graph build (main.cpp:132-134)."""

    def __init__(self, reads: ReadSet, device: int = 0):
        self.reads, self.device = reads, device
        self.timing: dict | None = None

    def getPrefixReads(self, remove_type: int = 2) -> np.ndarray:
        """uint8 mask: 1 = removed (duplicates except the greatest id; with ``remove_type`` 2 also proper prefixes of
        other reads and their reverse complements ``id ^ 1``).  ReadPreprocess.cpp:13-77."""
        lib = _lib.load()
        mask = np.zeros(self.reads.n, np.uint8)
        st = _reads_struct(self.reads)
        tm = _lib.Timing()
        _lib.check(lib.alga_gpu_prefix_reads(C.byref(st), remove_type, self.device, mask.ctypes.data, C.byref(tm)))
        self.timing = {"h2d_ms": tm.h2d_ms, "device_ms": tm.device_ms, "total_ms": tm.total_ms,
                       "kernel_launches": tm.kernel_launches}
        return mask


class GraphSimplifier:
    """Mirror of the reference's ``GraphSimplifier`` (include/GraphSimplifiers/GraphSimplifier.h:24-37) for the first step of
    ``simplifyGraphOld`` (GraphSimplifier.cpp:110-130), the one that works on the graph exactly as the graph creators leave
    it: ``Graph::sortEdgesByIncreasingOffset`` + ``cutNonAndWeaklyMetricTriangles``."""

    def __init__(self, graph: Graph, max_offset_parallel_paths: int, device: int = 0):
        self.graph, self.max_offset, self.device = graph, max_offset_parallel_paths, device
        self.timing: dict | None = None

    def cutNonAndWeaklyMetricTriangles(self) -> Graph:
        """GraphSimplifier.cpp:228-349; rows of the result are sorted by (offset, neighbour)."""
        lib = _lib.load()
        g = self.graph
        row_off = np.ascontiguousarray(g.row_off, dtype=np.uint64)
        nbr = np.ascontiguousarray(g.nbr, dtype=np.int32)
        off = np.ascontiguousarray(g.off, dtype=np.int32)
        cin = _lib.Csr(g.n, g.n_edges, row_off.ctypes.data_as(C.POINTER(C.c_uint64)), nbr.ctypes.data_as(C.POINTER(C.c_int32)),
                       off.ctypes.data_as(C.POINTER(C.c_int32)), 1)
        cout, tm = _lib.Csr(), _lib.Timing()
        _lib.check(lib.alga_gpu_cut_triangles(C.byref(cin), self.max_offset, self.device, C.byref(cout), C.byref(tm)))
        try:
            self.graph = _csr_to_graph(cout)
        finally:
            lib.alga_gpu_free_csr(C.byref(cout))
        self.timing = {"h2d_ms": tm.h2d_ms, "device_ms": tm.device_ms, "d2h_ms": tm.d2h_ms, "total_ms": tm.total_ms,
                       "kernel_launches": tm.kernel_launches}
        return self.graph

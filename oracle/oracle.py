"""TEST INFRASTRUCTURE ONLY -- ctypes binding of oracle/liboracle.so (the plain-C restatement)."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")


class _Reads(C.Structure):
    _fields_ = [("n", C.c_uint32), ("words", C.c_void_p), ("word_off", C.c_void_p), ("len_nt", C.c_void_p),
                ("align_from", C.c_void_p), ("align_to", C.c_void_p)]


class _VerifyParams(C.Structure):
    _fields_ = [("max_offset_pct", C.c_int32), ("min_offset", C.c_int32), ("min_overlap_area", C.c_int32),
                ("threshold_pct", C.c_int32), ("same_ends", C.c_int32), ("lcs_rate_pct", C.c_int32), ("lcs_band", C.c_int32)]


class _SupParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("max_offset_pct", "min_offset", "min_overlap_area", "threshold_pct", "same_ends",
                                          "kmer_length", "intervals", "kmer_length_bucket")]


class _InputParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("file_type", "trim_left", "trim_right", "rna", "str_threshold")]


INPUT_PLAIN, INPUT_FASTA, INPUT_FASTQ = 0, 1, 2


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in ("prefsuf_oracle.c", "verify_oracle.c", "supplement_oracle.cpp", "preprocess_oracle.c", "input_oracle.c", "simplify_oracle.c", "oracle.h",
                                            "Makefile")]
    if force or not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs):
        subprocess.run(["make", "-C", HERE, "liboracle.so"], check=True, stdout=subprocess.DEVNULL)
    return LIB


_lib = None


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.oracle_prefsuf.restype = C.POINTER(C.c_int32)
        _lib.oracle_prefsuf.argtypes = [C.POINTER(_Reads), C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                        C.POINTER(C.c_uint64)]
        _lib.oracle_fingerprints.restype = None
        _lib.oracle_fingerprints.argtypes = [C.POINTER(_Reads), C.c_int32] + [C.c_void_p] * 4
        _lib.oracle_verify_pairs.restype = None
        _lib.oracle_verify_pairs.argtypes = [C.POINTER(_Reads), C.c_void_p, C.c_uint64, C.POINTER(_VerifyParams),
                                             C.c_void_p]
        _lib.oracle_supplement.restype = C.POINTER(C.c_int32)
        _lib.oracle_supplement.argtypes = [C.POINTER(_Reads), C.c_void_p, C.c_uint64, C.POINTER(_SupParams),
                                           C.POINTER(C.c_uint64)]
        _lib.oracle_li_kmers.restype = None
        _lib.oracle_li_kmers.argtypes = [C.POINTER(_Reads), C.c_void_p, C.c_uint32, C.c_void_p, C.c_int32, C.c_int32,
                                         C.c_void_p, C.c_void_p]
        _lib.oracle_prefix_reads.restype = None
        _lib.oracle_prefix_reads.argtypes = [C.POINTER(_Reads), C.c_int32, C.c_void_p]
        _lib.oracle_read_input.restype = C.c_int
        _lib.oracle_read_input.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(_InputParams),
                                           C.POINTER(C.c_uint32), C.POINTER(C.c_void_p), C.POINTER(C.c_void_p),
                                           C.POINTER(C.c_void_p), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        _lib.oracle_remap.restype = C.c_int
        _lib.oracle_remap.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_uint32)]
        _lib.oracle_cut_triangles.restype = C.POINTER(C.c_int32)
        _lib.oracle_cut_triangles.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.POINTER(C.c_uint64)]
        _lib.oracle_free.argtypes = [C.c_void_p]
    return _lib


def _reads_struct(reads):
    return _Reads(reads.n, reads.words.ctypes.data, reads.word_off.ctypes.data, reads.len_nt.ctypes.data,
                  reads.align_from.ctypes.data, reads.align_to.ctypes.data)


def prefsuf(reads, min_overlap, rs_min_overlap, min_offset=0, max_len_cap=500) -> np.ndarray:
    """(E, 3) int32 edges (src, dst, offset) sorted by (src, dst)."""
    lib = _load()
    rs = _reads_struct(reads)
    ne = C.c_uint64(0)
    p = lib.oracle_prefsuf(C.byref(rs), min_overlap, rs_min_overlap, min_offset, max_len_cap, C.byref(ne))
    if not p:
        raise ValueError("oracle_prefsuf rejected the parameters")
    out = np.ctypeslib.as_array(p, shape=(max(ne.value, 1) * 3,))[: ne.value * 3].reshape(-1, 3).copy()
    lib.oracle_free(p)
    return out


def fingerprints(reads, L):
    lib = _load()
    rs = _reads_struct(reads)
    n = reads.n
    p64 = np.zeros(n, np.uint64); p32 = np.zeros(n, np.uint32)
    s64 = np.zeros(n, np.uint64); s32 = np.zeros(n, np.uint32)
    lib.oracle_fingerprints(C.byref(rs), L, p64.ctypes.data, p32.ctypes.data, s64.ctypes.data, s32.ctypes.data)
    return p64, p32, s64, s32


def verify_pairs(reads, pairs, threshold_pct, max_offset_pct, min_overlap_area, min_offset=0, same_ends=3, lcs_rate_pct=0, lcs_band=2):
    lib = _load()
    rs = _reads_struct(reads)
    pairs = np.ascontiguousarray(pairs, dtype=np.int32).reshape(-1, 3)
    vp = _VerifyParams(max_offset_pct, min_offset, min_overlap_area, threshold_pct, same_ends, lcs_rate_pct, lcs_band)
    out = np.zeros(pairs.shape[0], np.uint8)
    lib.oracle_verify_pairs(C.byref(rs), pairs.ctypes.data, pairs.shape[0], C.byref(vp), out.ctypes.data)
    return out


def supplement(reads, edges_in, threshold_pct, max_offset_pct, min_overlap_area, kmer_length_bucket, min_offset=0,
               same_ends=3, kmer_length=35, intervals=6) -> np.ndarray:
    """Graph after the error-rate supplement (main.cpp:300-355): (E, 3) int32 edges sorted by (src, dst)."""
    lib = _load()
    rs = _reads_struct(reads)
    e = np.ascontiguousarray(edges_in, dtype=np.int32).reshape(-1, 3)
    sp = _SupParams(max_offset_pct, min_offset, min_overlap_area, threshold_pct, same_ends, kmer_length, intervals,
                    kmer_length_bucket)
    ne = C.c_uint64(0)
    p = lib.oracle_supplement(C.byref(rs), e.ctypes.data, e.shape[0], C.byref(sp), C.byref(ne))
    out = np.ctypeslib.as_array(p, shape=(max(ne.value, 1) * 3,))[: ne.value * 3].reshape(-1, 3).copy()
    lib.oracle_free(p)
    return out


def li_kmers(reads, ids, priorities=(0, 1, 2, 3), kmer_length=35, intervals=6):
    """Read::getLIKmers of the reads ``ids``: (hash [n, intervals] uint64, ind [n, intervals] int32, -1 = absent)."""
    lib = _load()
    rs = _reads_struct(reads)
    ids = np.ascontiguousarray(ids, dtype=np.uint32)
    pr = np.ascontiguousarray(priorities, dtype=np.int32)
    h = np.zeros((ids.shape[0], intervals), np.uint64)
    ind = np.zeros((ids.shape[0], intervals), np.int32)
    lib.oracle_li_kmers(C.byref(rs), ids.ctypes.data, ids.shape[0], pr.ctypes.data, kmer_length, intervals, h.ctypes.data,
                        ind.ctypes.data)
    return h, ind


def prefix_reads(reads, remove_type=2) -> np.ndarray:
    """ReadPreprocess::getPrefixReads: uint8 mask of the reads that are removed (duplicates, prefix reads + revcomps)."""
    lib = _load()
    rs = _reads_struct(reads)
    mask = np.zeros(reads.n, np.uint8)
    lib.oracle_prefix_reads(C.byref(rs), remove_type, mask.ctypes.data)
    return mask


def read_input(text1: bytes, text2: bytes | None = None, file_type=INPUT_FASTA, trim_left=3, trim_right=3, rna=0,
               str_threshold=20):
    """InputReader::readInput on in-memory files -> (ReadSet, info dict); raises ValueError where the reference exits."""
    from alga_b200.readset import ReadSet
    lib = _load()
    ip = _InputParams(file_type, trim_left, trim_right, rna, str_threshold)
    b1 = np.frombuffer(text1, np.uint8)
    b2 = np.frombuffer(text2, np.uint8) if text2 is not None else None
    n = C.c_uint32(0)
    pl, po, pw = C.c_void_p(), C.c_void_p(), C.c_void_p()
    cn, cs = C.c_uint64(0), C.c_uint64(0)
    rc = lib.oracle_read_input(b1.ctypes.data if b1.size else None, b1.size,
                               (b2.ctypes.data if b2.size else C.cast(C.create_string_buffer(1), C.c_void_p)) if b2 is not None else None,
                               b2.size if b2 is not None else 0, C.byref(ip), C.byref(n), C.byref(pl), C.byref(po),
                               C.byref(pw), C.byref(cn), C.byref(cs))
    if rc != 0:
        raise ValueError({-1: "invalid character in a read", -2: "mate files differ in record count"}.get(rc, str(rc)))
    nn = n.value
    ln = np.ctypeslib.as_array(C.cast(pl, C.POINTER(C.c_uint32)), shape=(max(nn, 1),))[:nn].copy()
    off = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(nn + 1,)).copy()
    nw = int(off[nn])
    w = np.ctypeslib.as_array(C.cast(pw, C.POINTER(C.c_uint32)), shape=(max(nw, 1),))[:nw].copy()
    for q in (pl, po, pw):
        lib.oracle_free(q)
    return ReadSet(w, off, ln), dict(n_with_n=cn.value, n_str=cs.value)


def remap(len_nt, mask=None):
    """main.cpp:133-232 -> (old_id uint32 [n_out], paired_offset uint8 [n_out])."""
    lib = _load()
    ln = np.ascontiguousarray(len_nt, dtype=np.uint32)
    n = ln.shape[0]
    m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
    old = np.zeros(max(n, 1), np.uint32)
    po = np.zeros(max(n, 1), np.uint8)
    no = C.c_uint32(0)
    rc = lib.oracle_remap(ln.ctypes.data, m.ctypes.data if m is not None else None, n, old.ctypes.data, po.ctypes.data,
                          C.byref(no))
    if rc != 0:
        raise ValueError("a read is present without its reverse complement (the reference asserts, main.cpp:173)")
    return old[: no.value].copy(), po[: no.value].copy()


def cut_triangles(edges, n_nodes: int, max_offset: int) -> np.ndarray:
    """sortEdgesByIncreasingOffset + cutNonAndWeaklyMetricTriangles: (E', 3) int32 edges sorted by (src, offset, dst)."""
    lib = _load()
    e = np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 3)
    e = e[np.argsort(e[:, 0], kind="stable")]
    ne = C.c_uint64(0)
    p = lib.oracle_cut_triangles(e.ctypes.data, e.shape[0], n_nodes, max_offset, C.byref(ne))
    out = np.ctypeslib.as_array(p, shape=(max(ne.value, 1) * 3,))[: ne.value * 3].reshape(-1, 3).copy()
    lib.oracle_free(p)
    return out

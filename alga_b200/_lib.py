"""ctypes binding of libalga_gpu.so -- the C ABI declared in include/alga_gpu.h.

The library is built in-tree (``alga_b200/libalga_gpu.so``) by ``__graft_entry__.build()`` /
``make -C alga_b200/csrc``.  Loading fails loudly when it is missing: there is no other code path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ALGA_GPU_LIB") or os.path.join(HERE, "libalga_gpu.so")  # override: A/B builds only

ALGA_OK = 0
ERRORS = {-1: "ALGA_E_INVALID", -2: "ALGA_E_CUDA", -3: "ALGA_E_NOMEM", -4: "ALGA_E_CAPACITY"}


class AlgaGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{ERRORS.get(code, code)}: {msg}")
        self.code = code


class Reads(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("words", C.c_void_p), ("word_off", C.c_void_p),
                ("stride_words", C.c_uint32), ("len_nt", C.c_void_p), ("align_from", C.c_void_p),
                ("align_to", C.c_void_p)]


class PsParams(C.Structure):
    _fields_ = [("min_overlap", C.c_int32), ("rs_min_overlap", C.c_int32), ("min_offset", C.c_int32),
                ("max_len_cap", C.c_int32), ("device", C.c_int32), ("list_cap", C.c_int32), ("flags", C.c_int32)]


PS_FORCE_GENERIC = 1


class Csr(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("n_edges", C.c_uint64), ("row_off", C.POINTER(C.c_uint64)),
                ("nbr", C.POINTER(C.c_int32)), ("off", C.POINTER(C.c_int32)), ("borrowed", C.c_int32)]


class Timing(C.Structure):
    _fields_ = [("h2d_ms", C.c_double), ("device_ms", C.c_double), ("d2h_ms", C.c_double), ("total_ms", C.c_double),
                ("kernel_launches", C.c_uint64), ("n_spilled_targets", C.c_uint64), ("stage_ms", C.c_double * 8)]


class Shard(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("n_shard", C.c_uint32), ("n_total", C.c_uint32),
                ("peer_ws", C.c_void_p * 8), ("table_prefix", C.c_void_p), ("table_suffix", C.c_void_p)]


class VerifyParams(C.Structure):
    _fields_ = [("max_offset_pct", C.c_int32), ("min_offset", C.c_int32), ("min_overlap_area", C.c_int32),
                ("threshold_pct", C.c_int32), ("same_ends", C.c_int32), ("device", C.c_int32), ("lcs_rate_pct", C.c_int32),
                ("lcs_band", C.c_int32)]


class SupParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("max_offset_pct", "min_offset", "min_overlap_area", "threshold_pct", "same_ends",
                                          "kmer_length", "intervals", "kmer_length_bucket", "device")]


class InputParams(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("file_type", "trim_left", "trim_right", "rna", "str_threshold", "device")]


INPUT_PLAIN, INPUT_FASTA, INPUT_FASTQ = 0, 1, 2


class ReadSetOut(C.Structure):
    _fields_ = [("n_reads", C.c_uint32), ("stride_words", C.c_uint32), ("max_len_nt", C.c_uint32),
                ("words", C.POINTER(C.c_uint32)), ("len_nt", C.POINTER(C.c_uint32)), ("old_id", C.POINTER(C.c_uint32)),
                ("paired_offset", C.POINTER(C.c_uint8)), ("n_records", C.c_uint64 * 2), ("n_with_n", C.c_uint64),
                ("n_str", C.c_uint64), ("borrowed", C.c_int32)]


class DriverParams(C.Structure):
    _fields_ = [("input", InputParams), ("remove_type", C.c_int32), ("scale", C.c_float), ("min_overlap", C.c_int32),
                ("rs_min_overlap", C.c_int32)]


class OverlapGraphOut(C.Structure):
    _fields_ = [("reads", ReadSetOut), ("graph", Csr), ("avg_len", C.c_double), ("min_overlap", C.c_int32),
                ("rs_min_overlap", C.c_int32), ("li_kmer_length", C.c_int32), ("n_reads_in", C.c_uint32),
                ("n_records", C.c_uint64 * 2), ("n_with_n", C.c_uint64), ("n_str", C.c_uint64)]


# every symbol include/alga_gpu.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "alga_gpu_prefsuf_build": (C.c_int, [C.POINTER(Reads), C.POINTER(PsParams), C.POINTER(Csr), C.POINTER(Timing)]),
    "alga_gpu_prefsuf_build_multi": (C.c_int, [C.POINTER(Reads), C.POINTER(PsParams), C.c_int32, C.POINTER(Csr), C.POINTER(Timing)]),
    "alga_gpu_free_csr": (None, [C.POINTER(Csr)]),
    "alga_ps_plan_create": (C.c_int, [C.POINTER(_P), C.POINTER(PsParams)]),
    "alga_ps_plan_destroy": (None, [_P]),
    "alga_ps_plan_bind_reads_device": (C.c_int, [_P, C.POINTER(Reads), C.c_uint32]),
    "alga_ps_plan_upload_reads": (C.c_int, [_P, C.POINTER(Reads)]),
    "alga_ps_plan_run": (C.c_int, [_P, _P]),
    "alga_ps_stage_index": (C.c_int, [_P, _P]),
    "alga_ps_stage_phase1": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "alga_ps_stage_phase2": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64, _P, C.POINTER(_P),
                                       C.POINTER(C.c_uint64)]),
    "alga_ps_stage_csr": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.c_uint64, C.c_int, _P]),
    "alga_ps_plan_bind_reads_uniform": (C.c_int, [_P, C.POINTER(Reads), C.c_uint32]),
    "alga_ps_stage_index_range": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_int, _P]),
    "alga_ps_shard_ws_bytes": (C.c_uint64, [C.c_uint32, C.c_int32]),
    "alga_ps_shard_table_bytes": (C.c_uint64, [C.c_uint32, C.c_int32]),
    "alga_ps_shard_index_range": (C.c_int, [_P, C.POINTER(Shard), C.c_uint32, C.c_uint32, C.c_int, _P]),
    "alga_ps_shard_seed_keys": (C.c_int, [_P, C.POINTER(Shard), _P, C.c_uint32, C.c_uint32, _P, _P]),
    "alga_ps_shard_index_keys": (C.c_int, [_P, C.POINTER(Shard), _P, C.c_uint32, C.c_uint32, C.c_int, _P]),
    "alga_ps_set_bucket_load": (None, [C.c_int32]),
    "alga_ps_shard_phase1": (C.c_int, [_P, C.POINTER(Shard), _P]),
    "alga_ps_shard_phase2": (C.c_int, [_P, C.POINTER(Shard), _P]),
    "alga_ps_shard_csr": (C.c_int, [_P, C.POINTER(Shard), _P]),
    "alga_ps_plan_result_device": (C.c_int, [_P, C.POINTER(_P), C.POINTER(_P), C.POINTER(_P), C.POINTER(C.c_uint64)]),
    "alga_ps_plan_result_rows": (C.c_uint32, [_P]),
    "alga_ps_plan_result_host": (C.c_int, [_P, C.POINTER(Csr)]),
    "alga_ps_plan_result_host_pinned": (C.c_int, [_P, C.POINTER(Csr)]),
    "alga_gpu_host_alloc": (C.c_void_p, [C.c_size_t]),
    "alga_gpu_host_free": (None, [_P]),
    "alga_ps_plan_stats": (C.c_int, [_P, C.POINTER(Timing)]),
    "alga_gpu_fingerprints": (C.c_int, [C.POINTER(Reads), C.c_int32, C.c_int32, _P, _P, _P, _P]),
    "alga_gpu_pack_reads": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_int32, _P]),
    "alga_gpu_verify_pairs": (C.c_int, [C.POINTER(Reads), _P, C.c_uint64, C.POINTER(VerifyParams), _P]),
    "alga_gpu_prefix_reads": (C.c_int, [C.POINTER(Reads), C.c_int32, C.c_int32, _P, C.POINTER(Timing)]),
    "alga_gpu_read_input": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64, C.POINTER(InputParams), C.POINTER(ReadSetOut),
                                      C.POINTER(Timing)]),
    "alga_gpu_remap_reads": (C.c_int, [C.POINTER(Reads), _P, C.c_int32, C.POINTER(ReadSetOut), C.POINTER(Timing)]),
    "alga_gpu_free_read_set": (None, [C.POINTER(ReadSetOut)]),
    "alga_gpu_files_to_graph": (C.c_int, [_P, C.c_uint64, _P, C.c_uint64, C.POINTER(DriverParams), C.POINTER(OverlapGraphOut),
                                          C.POINTER(Timing)]),
    "alga_gpu_cut_triangles": (C.c_int, [C.POINTER(Csr), C.c_int32, C.c_int32, C.POINTER(Csr), C.POINTER(Timing)]),
    "alga_gpu_supplement": (C.c_int, [C.POINTER(Reads), C.POINTER(Csr), C.POINTER(SupParams), C.POINTER(Csr),
                                      C.POINTER(Timing)]),
    "alga_gpu_li_kmers": (C.c_int, [C.POINTER(Reads), _P, C.c_uint32, _P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "alga_gpu_device_count": (C.c_int, []),
    "alga_gpu_last_error": (C.c_char_p, []),
    "alga_gpu_version": (C.c_char_p, []),
}

_lib = None


def load() -> C.CDLL:
    """Load libalga_gpu.so; raises if the CUDA library has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "or `make -C alga_b200/csrc` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(code: int):
    if code != ALGA_OK:
        raise AlgaGpuError(code, load().alga_gpu_last_error().decode(errors="replace"))

"""Device-resident interface over the staged C ABI (``alga_ps_plan_*`` of include/alga_gpu.h).

PyTorch is used only as plumbing: it owns the device buffers the read set lives in, supplies the CUDA
stream and (in ``alga_b200.distributed``) the NCCL process group.  Every kernel that runs is one of
libalga_gpu.so's hand-written sm_100a kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from .readset import ReadSet

READ_PAD_BYTES = 256  # alga_gpu.h: `words` must be followed by this many readable bytes


class _DevArray:
    """Zero-copy view of plan-owned device memory for torch (``__cuda_array_interface__``)."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _as_tensor(ptr: int, shape, typestr: str, dtype, device) -> torch.Tensor:
    n = int(np.prod(shape))
    if n == 0 or not ptr:
        return torch.empty(tuple(shape), dtype=dtype, device=device)
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)


class DeviceReads:
    """A packed read set resident in HBM (torch tensors own the memory)."""

    def __init__(self, reads: ReadSet, device: torch.device, pinned: bool = False):
        self.n = reads.n
        self.device = device
        n_words = int(reads.words.shape[0])
        self.words = torch.zeros(n_words + READ_PAD_BYTES // 4, dtype=torch.int32, device=device)
        self.words[:n_words].copy_(torch.from_numpy(reads.words.view(np.int32)))
        self.len_nt = torch.from_numpy(reads.len_nt.view(np.int32)).to(device)
        self.align_from = torch.from_numpy(reads.align_from).to(device)
        self.align_to = torch.from_numpy(reads.align_to).to(device)
        self.stride = 0
        self.word_off = None
        w = int(reads.word_off[1] - reads.word_off[0]) if reads.n else 0
        if reads.n and w > 0 and np.array_equal(reads.word_off, np.arange(reads.n + 1, dtype=np.uint64) * np.uint64(w)):
            self.stride = w
        else:
            self.word_off = torch.from_numpy(reads.word_off.view(np.int64)).to(device)
        self.max_len = int(reads.len_nt.max()) if reads.n else 0

    @classmethod
    def from_tensors(cls, words: torch.Tensor, len_nt: torch.Tensor, stride: int, n: int, max_len: int,
                     word_off: torch.Tensor | None = None, align_from: torch.Tensor | None = None,
                     align_to: torch.Tensor | None = None) -> "DeviceReads":
        """Wrap device tensors that already hold a packed read set (``words`` padded by READ_PAD_BYTES)."""
        self = cls.__new__(cls)
        self.n, self.device = n, words.device
        self.words, self.len_nt, self.stride, self.word_off = words, len_nt, stride, word_off
        self.align_from, self.align_to = align_from, align_to
        self.max_len = max_len
        return self

    def struct(self) -> _lib.Reads:
        return _lib.Reads(self.n, self.words.data_ptr(), self.word_off.data_ptr() if self.word_off is not None else None,
                          self.stride, self.len_nt.data_ptr(),
                          self.align_from.data_ptr() if self.align_from is not None else None,
                          self.align_to.data_ptr() if self.align_to is not None else None)


class PrefSufPlan:
    """One GPU's workspace for GraphCreatorPrefSuf (GraphCreatorPrefSuf.cpp:73-126 + main.cpp:291)."""

    def __init__(self, min_overlap: int, rs_min_overlap: int, min_offset: int = 0, max_len_cap: int = 500,
                 device: int | torch.device = 0, list_cap: int = 0, force_generic: bool = False):
        self.lib = _lib.load()
        self.device = torch.device("cuda", device) if isinstance(device, int) else device
        self.params = _lib.PsParams(min_overlap, rs_min_overlap, min_offset, max_len_cap, self.device.index or 0, list_cap,
                                    _lib.PS_FORCE_GENERIC if force_generic else 0)
        self._h = C.c_void_p()
        _lib.check(self.lib.alga_ps_plan_create(C.byref(self._h), C.byref(self.params)))
        self.reads: DeviceReads | None = None

    def close(self):
        if self._h:
            self.lib.alga_ps_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def bind(self, reads: DeviceReads):
        self.reads = reads
        st = reads.struct()
        torch.cuda.current_stream(self.device).synchronize()
        _lib.check(self.lib.alga_ps_plan_bind_reads_device(self._h, C.byref(st), reads.max_len))

    def bind_uniform(self, reads: DeviceReads, len_nt: int):
        """Equal-length reads at a fixed stride, no flags: bound without a pass over them (they may still be arriving)."""
        self.reads = reads
        st = reads.struct()
        _lib.check(self.lib.alga_ps_plan_bind_reads_uniform(self._h, C.byref(st), len_nt))

    # ---- whole pipeline on this GPU -------------------------------------------------------------
    def run(self):
        _lib.check(self.lib.alga_ps_plan_run(self._h, self._stream()))

    def stats(self) -> dict:
        tm = _lib.Timing()
        _lib.check(self.lib.alga_ps_plan_stats(self._h, C.byref(tm)))
        d = {k: getattr(tm, k) for k, _ in _lib.Timing._fields_ if k != "stage_ms"}
        d["stage_ms"] = dict(zip(("index", "phase1", "transpose", "phase2", "csr"), list(tm.stage_ms)[:5]))
        d["n_row_overflow"] = int(tm.stage_ms[5])
        d["n_hard_sources"] = int(tm.stage_ms[6])
        return d

    # ---- stages (sharded runs) ------------------------------------------------------------------
    def stage_index(self):
        _lib.check(self.lib.alga_ps_stage_index(self._h, self._stream()))

    def stage_phase1(self, lo: int, hi: int) -> torch.Tensor:
        """(n, 3) int32 device tensor of (b, c, offset) for source reads b in [lo, hi) -- a view of plan memory."""
        p, n = C.c_void_p(), C.c_uint64()
        _lib.check(self.lib.alga_ps_stage_phase1(self._h, lo, hi, self._stream(), C.byref(p), C.byref(n)))
        return _as_tensor(p.value, (n.value, 3), "<i4", torch.int32, self.device)

    def stage_phase2(self, lo: int, hi: int, triples: torch.Tensor) -> torch.Tensor:
        """Transitive reduction for target reads c in [lo, hi); ``triples`` = phase-1 edges with c in range."""
        triples = triples.contiguous()
        p, n = C.c_void_p(), C.c_uint64()
        _lib.check(self.lib.alga_ps_stage_phase2(self._h, lo, hi, triples.data_ptr() if triples.numel() else None,
                                                 triples.shape[0], self._stream(), C.byref(p), C.byref(n)))
        return _as_tensor(p.value, (n.value, 3), "<i4", torch.int32, self.device)

    def stage_csr(self, lo: int, hi: int, triples: torch.Tensor, swap_direction: bool = False):
        triples = triples.contiguous()
        _lib.check(self.lib.alga_ps_stage_csr(self._h, lo, hi, triples.data_ptr() if triples.numel() else None,
                                              triples.shape[0], 1 if swap_direction else 0, self._stream()))

    # ---- sharded build, exchange inside the kernels over peer memory (alga_gpu.h: alga_ps_shard_*) ---
    def stage_index_range(self, lo: int, hi: int, first: bool):
        _lib.check(self.lib.alga_ps_stage_index_range(self._h, lo, hi, 1 if first else 0, self._stream()))

    def shard_ws_bytes(self, n_shard: int, world: int) -> int:
        return int(self.lib.alga_ps_shard_ws_bytes(n_shard, world))

    @staticmethod
    def shard_struct(rank: int, world: int, n_shard: int, n_total: int, peer_ws_ptrs, table_prefix: int = 0,
                     table_suffix: int = 0) -> _lib.Shard:
        sh = _lib.Shard(rank, world, n_shard, n_total)
        for p, ptr in enumerate(peer_ws_ptrs):
            sh.peer_ws[p] = int(ptr)
        sh.table_prefix = table_prefix or None
        sh.table_suffix = table_suffix or None
        return sh

    def shard_table_bytes(self, n_total: int, world: int) -> int:
        return int(self.lib.alga_ps_shard_table_bytes(n_total, world))

    def shard_index_range(self, sh: _lib.Shard, lo: int, hi: int, first: bool, which: int = 0):
        """which: 0 = both tables, 1 = the prefix table only, 2 = the suffix table only (alga_gpu.h)."""
        _lib.check(self.lib.alga_ps_shard_index_range(self._h, C.byref(sh), lo, hi, (1 if first else 0) | (which << 1), self._stream()))

    def shard_seed_keys(self, sh: _lib.Shard, shard_words: torch.Tensor, stride: int, n_reads: int, keys: torch.Tensor):
        """12-byte seed records of this rank's own reads (``shard_words``: their packed words at ``stride`` words per read)."""
        _lib.check(self.lib.alga_ps_shard_seed_keys(self._h, C.byref(sh), shard_words.data_ptr(), stride, n_reads, keys.data_ptr(),
                                                    self._stream()))

    def shard_index_keys(self, sh: _lib.Shard, keys: torch.Tensor, lo: int, hi: int, first: bool):
        """Insert the seeds of the reads [lo, hi) that fall into this rank's slice, out of their seed records ``keys``."""
        _lib.check(self.lib.alga_ps_shard_index_keys(self._h, C.byref(sh), keys.data_ptr() if hi > lo else None, lo, hi,
                                                     1 if first else 0, self._stream()))

    def shard_phase1(self, sh: _lib.Shard):
        _lib.check(self.lib.alga_ps_shard_phase1(self._h, C.byref(sh), self._stream()))

    def shard_phase2(self, sh: _lib.Shard):
        _lib.check(self.lib.alga_ps_shard_phase2(self._h, C.byref(sh), self._stream()))

    def shard_csr(self, sh: _lib.Shard):
        _lib.check(self.lib.alga_ps_shard_csr(self._h, C.byref(sh), self._stream()))

    # ---- results --------------------------------------------------------------------------------
    def result_device(self):
        """(row_off int64 [rows+1], nbr int32 [E], off int32 [E]) -- views of plan memory."""
        ro, nb, of, ne = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_uint64()
        _lib.check(self.lib.alga_ps_plan_result_device(self._h, C.byref(ro), C.byref(nb), C.byref(of), C.byref(ne)))
        rows = self.result_rows()
        return (_as_tensor(ro.value, (rows + 1,), "<i8", torch.int64, self.device),
                _as_tensor(nb.value, (ne.value,), "<i4", torch.int32, self.device),
                _as_tensor(of.value, (ne.value,), "<i4", torch.int32, self.device))

    def result_rows(self) -> int:
        return int(self.lib.alga_ps_plan_result_rows(self._h))

    def n_edges(self) -> int:
        ne = C.c_uint64()
        _lib.check(self.lib.alga_ps_plan_result_device(self._h, None, None, None, C.byref(ne)))
        return int(ne.value)

    def result_host(self):
        from .graph_creator import _csr_to_graph

        csr = _lib.Csr()
        _lib.check(self.lib.alga_ps_plan_result_host(self._h, C.byref(csr)))
        try:
            return _csr_to_graph(csr)
        finally:
            self.lib.alga_gpu_free_csr(C.byref(csr))

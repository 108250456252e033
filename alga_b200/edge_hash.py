"""Order-independent 128-bit digest of an edge set {(source, target, offset)}.

The parity gate of every measured number: ``bench.py`` digests the graph the last timed step left in HBM (per rank;
the digests of disjoint row ranges add up) and compares it with the digest of the reference's own graph for the same
read set (``tests/golden/full_*.json``, produced by ``tests/golden/make_full_golden.py`` through ``oracle/_ref``).
Integer-only torch code: the same function digests a CSR on the GPU and an (E, 3) array on the host.
"""
from __future__ import annotations

import numpy as np
import torch

from .synth_dev import _c, _lsr, splitmix64

_M64 = (1 << 64) - 1


def _sum64(h: torch.Tensor) -> int:
    lo = int((h & 0xFFFFFFFF).sum().item())
    hi = int(_lsr(h, 32).sum().item())
    return (lo + (hi << 32)) & _M64


def digest_triples(src: torch.Tensor, dst: torch.Tensor, off: torch.Tensor) -> tuple:
    """-> (a, b): two 64-bit sums over per-edge hashes (mod 2^64); digests of disjoint edge sets add."""
    if src.numel() == 0:
        return (0, 0)
    s, d, o = src.to(torch.int64), dst.to(torch.int64), off.to(torch.int64)
    h = splitmix64(s * _c(0x9E3779B97F4A7C15) + d)
    h = splitmix64(h ^ (o * _c(0xC2B2AE3D27D4EB4F)))
    return (_sum64(h), _sum64(splitmix64(h)))


def digest_csr(row_off: torch.Tensor, nbr: torch.Tensor, off: torch.Tensor, first_row: int = 0) -> tuple:
    """Digest of CSR rows [first_row, first_row + n): row_off (n+1) int64, nbr / off (E) int32, on any device."""
    n = row_off.numel() - 1
    deg = (row_off[1:] - row_off[:-1]).to(torch.int64)
    src = torch.repeat_interleave(torch.arange(first_row, first_row + n, dtype=torch.int64, device=row_off.device), deg,
                                  output_size=int(nbr.numel()))
    return digest_triples(src, nbr, off)


def digest_edges(edges: np.ndarray) -> tuple:
    """Digest of an (E, 3) int32 host array of (source, target, offset)."""
    e = torch.from_numpy(np.ascontiguousarray(edges, dtype=np.int32).reshape(-1, 3))
    return digest_triples(e[:, 0], e[:, 1], e[:, 2])


def add(a: tuple, b: tuple) -> tuple:
    return ((a[0] + b[0]) & _M64, (a[1] + b[1]) & _M64)

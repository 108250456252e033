"""Counter-based synthetic read sets, generated where they are needed (CPU or GPU) with identical results.

Same distributions as ``synth`` (iid uniform genome, uniform fragment start and strand, fragment
length ~ N(500, 50), optional copied repeat segments, iid substitutions), but every random quantity is a pure function
``splitmix64(stream, counter)`` and all arithmetic is 64-bit integer, so a torch CPU tensor and a torch CUDA tensor give
the same bytes: BASELINE configs 4 and 5 (57 M / 200 M strand-reads) are generated on the GPU in seconds instead of minutes
on the host, every rank of a multi-GPU run can generate the same read set by itself, and a golden result computed once for
a (config, seed) stays valid.  The normal is an Irwin-Hall sum of eight 16-bit uniforms (integer only).

Preprocessing that precedes the hot path is restated as in ``synth``: trim 3+3 (InputReader.cpp:298-303), id layout
(4k, 4k+1, 4k+2, 4k+3) = (revcomp mate 1, mate 1, revcomp mate 2, mate 2) (InputReader.cpp:53-80), duplicates keep the
greatest id (ReadPreprocess.cpp:13-77; equal-length reads have no proper prefixes), survivors renumbered in order
(main.cpp:150-232).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from .synth import CONFIGS, TRIM, PrefSufParams, derive_params

_M64 = (1 << 64) - 1


def _c(x: int) -> int:
    """64-bit constant as a signed Python int (torch int64 arithmetic wraps)."""
    x &= _M64
    return x - (1 << 64) if x >> 63 else x


def _lsr(x: torch.Tensor, k: int) -> torch.Tensor:
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix64(x: torch.Tensor) -> torch.Tensor:
    z = x + _c(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _c(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _c(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _stream(seed: int, tag: int, idx: torch.Tensor) -> torch.Tensor:
    """Random 64-bit words: one per counter value, independent per (seed, tag)."""
    base = _c((seed * 0xD6E8FEB86659FD93 + tag * 0xA0761D6478BD642F) & _M64)
    return splitmix64(splitmix64(idx + base))


def _scalar(seed: int, tag: int, i: int) -> int:
    return int(_lsr(_stream(seed, tag, torch.tensor([i], dtype=torch.int64)), 1)[0])


def make_genome(size: int, seed: int, device, repeats: int = 0, repeat_len=(1000, 10000)) -> torch.Tensor:
    g = torch.empty(size, dtype=torch.uint8, device=device)
    step = 1 << 26
    for lo in range(0, size, step):
        idx = torch.arange(lo, min(size, lo + step), dtype=torch.int64, device=device)
        g[lo:lo + idx.numel()] = _lsr(_stream(seed, 1, idx), 62).to(torch.uint8)
    for j in range(repeats):
        ln = repeat_len[0] + _scalar(seed, 2, 3 * j) % (repeat_len[1] - repeat_len[0] + 1)
        ln = min(ln, size // 4)
        src = _scalar(seed, 2, 3 * j + 1) % (size - ln)
        dst = _scalar(seed, 2, 3 * j + 2) % (size - ln)
        g[dst:dst + ln] = g[src:src + ln].clone()
    return g


def _pack(codes: torch.Tensor) -> torch.Tensor:
    """(n, len) uint8 codes -> (n, ceil(len/16)) int32 words, reference layout (readset.pack_matrix)."""
    n, ln = codes.shape
    W = (ln + 15) // 16
    if W * 16 != ln:
        codes = torch.nn.functional.pad(codes, (0, W * 16 - ln))
    sh = (2 * torch.arange(16, dtype=torch.int64, device=codes.device))[None, None, :]
    w = (codes.view(n, W, 16).to(torch.int64) << sh).sum(dim=2)
    return _to_i32(w)


def _to_i32(w: torch.Tensor) -> torch.Tensor:
    """low 32 bits of non-negative int64 values as int32 bit patterns"""
    return torch.where(w >= (1 << 31), w - (1 << 32), w).to(torch.int32)


def _substitute(codes: torch.Tensor, p: float, seed: int, tag: int, first: int) -> torch.Tensor:
    if p <= 0:
        return codes
    n, ln = codes.shape
    idx = first * ln + torch.arange(n * ln, dtype=torch.int64, device=codes.device).view(n, ln)
    h = _stream(seed, tag, idx)
    hit = _lsr(h, 40) < int(p * (1 << 24))
    delta = (1 + _lsr(h, 8) % 3).to(torch.uint8)
    return torch.where(hit, (codes + delta) & 3, codes)


def _strand_words(genome: torch.Tensor, read_len: int, coverage: float, paired: bool, seed: int, error: float,
                  chunk: int = 1 << 20) -> tuple:
    """Packed strand-reads (before duplicate removal) in the reference's id order -> ((N, W) int32, records)."""
    dev = genome.device
    G = genome.shape[0]
    n = int(G * coverage / ((2 if paired else 1) * read_len))
    tl = read_len - 2 * TRIM
    W = (tl + 15) // 16
    per = 4 if paired else 2
    out = torch.empty((per * n, W), dtype=torch.int32, device=dev)
    ar = torch.arange(read_len, dtype=torch.int64, device=dev)[None, :]
    for lo in range(0, n, chunk):
        k = torch.arange(lo, min(n, lo + chunk), dtype=torch.int64, device=dev)
        m = k.numel()
        flip = (_lsr(_stream(seed, 5, k), 63) == 1)[:, None]
        if paired:
            h = _stream(seed, 3, k)
            h2 = splitmix64(h)
            S = sum(_lsr(h, 16 * j) & 0xFFFF for j in range(4)) + sum(_lsr(h2, 16 * j) & 0xFFFF for j in range(4))
            frag = 500 + torch.div((S - 262140) * 50, 53510, rounding_mode="floor")
            frag = torch.clamp(frag, min=read_len, max=G)
            start = _lsr(_stream(seed, 4, k), 1) % (G - frag + 1)
            a = genome[start[:, None] + ar]
            b = 3 - genome[(start + frag - read_len)[:, None] + ar].flip(1)
            m1 = torch.where(flip, b, a)
            m2 = torch.where(flip, a, b)
            m1 = _substitute(m1, error, seed, 6, lo)
            m2 = _substitute(m2, error, seed, 7, lo)
            t1, t2 = m1[:, TRIM:read_len - TRIM], m2[:, TRIM:read_len - TRIM]
            blk = out[per * lo:per * (lo + m)].view(m, 4, W)
            blk[:, 0] = _pack(3 - t1.flip(1))
            blk[:, 1] = _pack(t1)
            blk[:, 2] = _pack(3 - t2.flip(1))
            blk[:, 3] = _pack(t2)
        else:
            start = _lsr(_stream(seed, 4, k), 1) % (G - read_len + 1)
            a = genome[start[:, None] + ar]
            a = torch.where(flip, 3 - a.flip(1), a)
            a = _substitute(a, error, seed, 6, lo)
            t1 = a[:, TRIM:read_len - TRIM]
            blk = out[per * lo:per * (lo + m)].view(m, 2, W)
            blk[:, 0] = _pack(3 - t1.flip(1))
            blk[:, 1] = _pack(t1)
    return out, n * (2 if paired else 1)


def _row_keys(words: torch.Tensor) -> tuple:
    """Two independent 64-bit hashes per packed read."""
    n, W = words.shape
    h1 = torch.zeros(n, dtype=torch.int64, device=words.device)
    h2 = torch.full((n,), 0x1234567, dtype=torch.int64, device=words.device)
    for j in range(W):
        w = words[:, j].to(torch.int64) & 0xFFFFFFFF
        h1 = splitmix64(h1 ^ w)
        h2 = splitmix64(h2 + w * _c(0x9E3779B97F4A7C15) + j)
    return h1, h2


def remove_duplicate_rows(words: torch.Tensor) -> torch.Tensor:
    """Keep-mask of ReadPreprocess for equal-length reads: among identical strand-reads the greatest id survives."""
    n = words.shape[0]
    h1, h2 = _row_keys(words)
    o = torch.sort(h2, stable=True).indices
    o = o[torch.sort(h1[o], stable=True).indices]  # by (h1, h2), ties in id order
    del h1, h2
    same_as_next = torch.zeros(n, dtype=torch.bool, device=words.device)
    step = 1 << 24
    for lo in range(0, n - 1, step):
        hi = min(n - 1, lo + step)
        same_as_next[lo:hi] = (words[o[lo:hi]] == words[o[lo + 1:hi + 1]]).all(dim=1)
    keep = torch.ones(n, dtype=torch.bool, device=words.device)
    keep[o[same_as_next]] = False  # an identical read with a greater id follows
    return keep


@dataclass
class DevWorkload:
    name: str
    words: torch.Tensor  # (N, W) int32, packed strand-reads after duplicate removal, on `device`
    len_nt: int
    params: PrefSufParams
    records: int
    genome_size: int
    seed: int

    @property
    def n(self) -> int:
        return int(self.words.shape[0])

    def to_readset(self):
        """Host ReadSet (small workloads: tests, CPU legs)."""
        from .readset import ReadSet

        w = self.words.cpu().numpy().view(np.uint32)
        n, W = w.shape
        return ReadSet(w.reshape(-1), np.arange(n + 1, dtype=np.uint64) * W, np.full(n, self.len_nt, np.uint32))


def make_workload(name: str, genome_size: int, read_len: int, coverage: float, paired: bool, seed: int,
                  error: float = 0.0, repeats: int = 0, device="cpu") -> DevWorkload:
    device = torch.device(device)
    genome = make_genome(genome_size, seed, device, repeats=repeats)
    words, records = _strand_words(genome, read_len, coverage, paired, seed, error)
    del genome
    keep = remove_duplicate_rows(words)
    assert bool((keep[0::2] == keep[1::2]).all())  # a strand-read and its twin are duplicates of mirrored partners
    words = words[keep].contiguous()
    tl = read_len - 2 * TRIM
    return DevWorkload(name, words, tl, derive_params(float(tl)), records, genome_size, seed)


def make_config(name: str, scale: float = 1.0, device="cpu", seed_offset: int = 0) -> DevWorkload:
    kw = dict(CONFIGS[name])
    kw["genome_size"] = max(20_000, int(kw["genome_size"] * scale))
    kw["seed"] += seed_offset
    return make_workload(name if scale == 1.0 else f"{name}@{scale:g}", device=device, **kw)

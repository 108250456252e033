"""Packed read set in the reference's own 2-bit layout.

Layout follows ``Read::createSequence`` / ``Bitset`` of the reference
(``src/DataStructures/Read.cpp:40-68``, ``include/DataStructures/Bitset.h:38-45``):
A=0, C=1, G=2, T=3 (anything else -> 0); nucleotide ``j`` of a read occupies bits
``2*(j%16)`` (low) and ``2*(j%16)+1`` (high) of 32-bit block ``j//16``; unused tail
bits are zero.  A read of length 0 stands for a removed (``nullptr``) read.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

_NT = np.frombuffer(b"ACGT", dtype=np.uint8)
_CODE = np.zeros(256, dtype=np.uint8)
_CODE[ord("C")] = 1
_CODE[ord("G")] = 2
_CODE[ord("T")] = 3
_CODE[ord("c")] = 1
_CODE[ord("g")] = 2
_CODE[ord("t")] = 3


@dataclass
class ReadSet:
    """Concatenated packed reads: ``words[word_off[i]:word_off[i+1]]`` holds read ``i``."""

    words: np.ndarray  # uint32
    word_off: np.ndarray  # uint64, n+1
    len_nt: np.ndarray  # uint32, n ; 0 => null read
    align_from: np.ndarray = field(default=None)  # uint8, n
    align_to: np.ndarray = field(default=None)  # uint8, n

    def __post_init__(self):
        self.words = np.ascontiguousarray(self.words, dtype=np.uint32)
        self.word_off = np.ascontiguousarray(self.word_off, dtype=np.uint64)
        self.len_nt = np.ascontiguousarray(self.len_nt, dtype=np.uint32)
        n = self.len_nt.shape[0]
        if self.align_from is None:
            self.align_from = (self.len_nt > 0).astype(np.uint8)
        if self.align_to is None:
            self.align_to = (self.len_nt > 0).astype(np.uint8)
        self.align_from = np.ascontiguousarray(self.align_from, dtype=np.uint8)
        self.align_to = np.ascontiguousarray(self.align_to, dtype=np.uint8)
        if self.word_off.shape[0] != n + 1:
            raise ValueError("word_off must have n+1 entries")
        if self.align_from.shape[0] != n or self.align_to.shape[0] != n:
            raise ValueError("flag arrays must have n entries")

    @property
    def n(self) -> int:
        return int(self.len_nt.shape[0])

    def codes(self, i: int) -> np.ndarray:
        """2-bit codes of read ``i`` as a uint8 array (one entry per nucleotide)."""
        ln = int(self.len_nt[i])
        w = self.words[int(self.word_off[i]) : int(self.word_off[i + 1])]
        j = np.arange(ln)
        return ((w[j >> 4] >> ((j & 15) * 2).astype(np.uint32)) & 3).astype(np.uint8)

    def sequence(self, i: int) -> str:
        return _NT[self.codes(i)].tobytes().decode()


def codes_from_ascii(seq: bytes | str) -> np.ndarray:
    if isinstance(seq, str):
        seq = seq.encode()
    return _CODE[np.frombuffer(seq, dtype=np.uint8)]


def pack_matrix(codes: np.ndarray) -> np.ndarray:
    """Pack an (n, l) uint8 code matrix into (n, ceil(l/16)) uint32 words."""
    n, l = codes.shape
    w = (l + 15) // 16
    pad = np.zeros((n, w * 16), dtype=np.uint32)
    pad[:, :l] = codes
    sh = (np.arange(16, dtype=np.uint32) * 2)[None, None, :]
    return (pad.reshape(n, w, 16) << sh).sum(axis=2, dtype=np.uint64).astype(np.uint32)


def from_code_matrix(codes: np.ndarray, **flags) -> ReadSet:
    """Equal-length reads given as an (n, l) matrix of 2-bit codes."""
    n, l = codes.shape
    w = (l + 15) // 16
    words = pack_matrix(codes).reshape(-1)
    off = np.arange(n + 1, dtype=np.uint64) * np.uint64(w)
    return ReadSet(words, off, np.full(n, l, dtype=np.uint32), **flags)


def from_code_list(reads: list, **flags) -> ReadSet:
    """Variable-length reads; ``None`` or an empty array stands for a null read."""
    n = len(reads)
    lens = np.array([0 if r is None else len(r) for r in reads], dtype=np.uint32)
    wcnt = (lens.astype(np.uint64) + np.uint64(15)) // np.uint64(16)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(wcnt, out=off[1:])
    words = np.zeros(int(off[-1]), dtype=np.uint32)
    for i, r in enumerate(reads):
        if lens[i] == 0:
            continue
        words[int(off[i]) : int(off[i + 1])] = pack_matrix(np.asarray(r, dtype=np.uint8)[None, :])[0]
    return ReadSet(words, off, lens, **flags)


def revcomp_codes(codes: np.ndarray) -> np.ndarray:
    """Reverse complement along the last axis (complement of code c is 3-c)."""
    return (3 - codes[..., ::-1]).astype(np.uint8)

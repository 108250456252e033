"""Read-id-range sharding of GraphCreatorPrefSuf over the GPUs of one box (one process per GPU).

The reference has no distributed path; the unit that shards naturally is the read:
phase 1 of GraphCreatorPrefSuf.cpp:397-402 depends only on the source read b, phase 2 (:403-483) only on the
target read c, and the rows of the final adjacency are disjoint by source read.  Rank r owns the reads
[r * n_shard, (r + 1) * n_shard).  Per build:

    1. all-gather of the 2-bit packed reads (NCCL over NVLink; 4 W bytes per node)      -> every GPU holds all reads
    2. seed index over all reads (replicated; libalga_gpu kernel)
    3. phase 1 for own source reads                                                      -> (b, c, o) edges
    4. all-to-all of those edges to the owner of c
    5. phase 2 (transitive reduction) for own target reads                               -> surviving (a, c, o)
    6. all-to-all of the survivors to the owner of a
    7. CSR assembly of own rows

PyTorch supplies the device buffers and the process group only; all graph work is libalga_gpu.so.
``route_triples`` is device-agnostic so that the exchange logic is tested with gloo on CPU tensors.
"""
from __future__ import annotations

import time

import numpy as np
import torch
import torch.distributed as dist

from .plan import READ_PAD_BYTES, DeviceReads, PrefSufPlan


def owner_of(ids: torch.Tensor, n_shard: int, world: int) -> torch.Tensor:
    return torch.clamp(torch.div(ids, n_shard, rounding_mode="floor"), max=world - 1)


def route_triples(triples: torch.Tensor, col: int, n_shard: int, world: int, group=None) -> torch.Tensor:
    """Send every (b, c, o) row to the rank owning ``row[col]``; returns the rows this rank received."""
    triples = triples.reshape(-1, 3)
    dest = owner_of(triples[:, col].to(torch.int64), n_shard, world)
    order = torch.argsort(dest, stable=True)
    send = triples[order].contiguous()
    counts = torch.bincount(dest, minlength=world).to(torch.int64)
    recv_counts = torch.empty_like(counts)
    dist.all_to_all_single(recv_counts, counts, group=group)
    in_splits = counts.tolist()
    out_splits = recv_counts.tolist()
    out = torch.empty((int(sum(out_splits)), 3), dtype=triples.dtype, device=triples.device)
    dist.all_to_all_single(out, send, output_split_sizes=out_splits, input_split_sizes=in_splits, group=group)
    return out


def interleave_shards(reads, rank: int, world: int, device) -> tuple[torch.Tensor, int]:
    """Weak-scaling workload: every rank generated one chromosome's (equal-length) read set; redistribute so that
    global twin-pair g = t * world + r (t-th twin pair of chromosome r) and rank q owns a contiguous id range.
    Returns (this rank's packed words [n_shard, W] int32 on ``device``, total number of nodes)."""
    n = reads.n
    W = int(reads.word_off[1] - reads.word_off[0])
    words = torch.from_numpy(reads.words.view(np.int32).reshape(n, W)).to(device)
    cnt = torch.tensor([n // 2], dtype=torch.int64, device=device)
    dist.all_reduce(cnt, op=dist.ReduceOp.MIN)
    twins = (int(cnt.item()) // world) * world  # twin pairs kept per chromosome
    per = twins // world
    send = words[: 2 * twins].reshape(world, per, 2 * W).contiguous()  # slice q goes to rank q
    recv = torch.empty_like(send)  # recv[r] = twins [rank*per, (rank+1)*per) of chromosome r
    dist.all_to_all_single(recv, send)
    shard = recv.permute(1, 0, 2).reshape(per * world * 2, W).contiguous()  # local twin (t, r) -> row order t*world + r
    return shard, 2 * twins * world


class ShardedPrefSuf:
    """One rank of the sharded overlap-graph build for equal-length reads."""

    def __init__(self, min_overlap, rs_min_overlap, min_offset, max_len_cap, device, rank, world, len_nt, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.device = device
        self.len_nt = len_nt
        self.plan = PrefSufPlan(min_overlap, rs_min_overlap, min_offset, max_len_cap, device=device)
        self._full = None
        self._len = None
        self._launch_mark = 0
        self.stage_ms = {}
        self._ev = None

    def _buffers(self, n_shard: int, W: int):
        n_total = n_shard * self.world
        if self._full is None or self._full.numel() != n_total * W + READ_PAD_BYTES // 4:
            self._full = torch.zeros(n_total * W + READ_PAD_BYTES // 4, dtype=torch.int32, device=self.device)
            self._len = torch.full((n_total,), self.len_nt, dtype=torch.int32, device=self.device)
        return n_total

    def run(self, shard_words: torch.Tensor):
        n_shard, W = shard_words.shape
        n_total = self._buffers(n_shard, W)
        lo, hi = self.rank * n_shard, (self.rank + 1) * n_shard
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(8)]
        before = self.plan.stats()["kernel_launches"]
        marks[0].record()
        dist.all_gather_into_tensor(self._full[: n_total * W], shard_words.reshape(-1), group=self.group)
        marks[1].record()
        dr = DeviceReads.from_tensors(self._full, self._len, stride=W, n=n_total, max_len=self.len_nt)
        self.plan.bind(dr)
        self.plan.stage_index()
        marks[2].record()
        t1 = self.plan.stage_phase1(lo, hi)
        marks[3].record()
        t1r = route_triples(t1, 1, n_shard, self.world, self.group)
        marks[4].record()
        t2 = self.plan.stage_phase2(lo, hi, t1r)
        marks[5].record()
        t2r = route_triples(t2, 0, n_shard, self.world, self.group)
        marks[6].record()
        self.plan.stage_csr(lo, hi, t2r)
        marks[7].record()
        self._ev = marks
        self._launches = self.plan.stats()["kernel_launches"] - before

    def stats(self) -> dict:
        names = ("allgather", "index", "phase1", "route1", "phase2", "route2", "csr")
        self._ev[-1].synchronize()
        ms = {k: self._ev[i].elapsed_time(self._ev[i + 1]) for i, k in enumerate(names)}
        return {"kernel_launches": self._launches, "stage_ms": ms}

    def total_edges(self) -> int:
        t = torch.tensor([self.plan.n_edges()], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def result_device(self):
        return self.plan.result_device()

    def e2e(self, shard_words: torch.Tensor, steps: int, n_nodes_total: int) -> dict:
        """Same build with HOST buffers: pinned shard in, this rank's CSR rows out (pinned), wall clock, max over ranks."""
        host_in = shard_words.cpu().pin_memory()
        dev_in = torch.empty_like(shard_words)
        self.run(shard_words)
        ro, nb, of = self.result_device()
        cap = int(nb.numel() * 1.1) + 1024
        h_ro = torch.empty(ro.numel(), dtype=ro.dtype).pin_memory()
        h_nb = torch.empty(cap, dtype=nb.dtype).pin_memory()
        h_of = torch.empty(cap, dtype=of.dtype).pin_memory()
        d2h = 0

        def one():
            nonlocal d2h
            dev_in.copy_(host_in, non_blocking=True)
            self.run(dev_in)
            ro, nb, of = self.result_device()
            e = nb.numel()
            h_ro.copy_(ro, non_blocking=True)
            h_nb[:e].copy_(nb, non_blocking=True)
            h_of[:e].copy_(of, non_blocking=True)
            torch.cuda.synchronize(self.device)
            d2h = ro.numel() * 8 + 8 * e

        one()
        dist.barrier(group=self.group)
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dist.barrier(group=self.group)
        ms = torch.tensor([1e3 * (time.perf_counter() - t0) / steps], dtype=torch.float64, device=self.device)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX, group=self.group)
        b = torch.tensor([host_in.numel() * 4, d2h], dtype=torch.int64, device=self.device)
        dist.all_reduce(b, group=self.group)
        return {"value": n_nodes_total / (float(ms.item()) / 1e3), "unit": "reads/s", "h2d_bytes_per_step": int(b[0].item()),
                "d2h_bytes_per_step": int(b[1].item()), "ms_per_step": float(ms.item()), "steps": steps,
                "call": "alga_b200.distributed.ShardedPrefSuf.run with pinned host shard in / pinned host CSR rows out"}

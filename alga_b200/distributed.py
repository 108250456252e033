"""Read-id-range sharding of GraphCreatorPrefSuf over the GPUs of one box (one process per GPU).

The reference has no distributed path; the unit that shards naturally is the read:
phase 1 of GraphCreatorPrefSuf.cpp:397-402 depends only on the source read b, phase 2 (:403-483) only on the
target read c, and the rows of the final adjacency are disjoint by source read.  Rank r owns the reads
[r * n_shard, (r + 1) * n_shard).  ``ShardedPrefSuf`` is the product path: the exchanges happen inside the kernels
over peer memory.  ``route_triples`` is the same routing rule written with torch collectives; it documents the
semantics and is what the CPU (gloo) tests exercise, together with the triple-based stage calls of the C ABI.

PyTorch supplies device buffers, streams, symmetric memory and the process group only; all graph work is
libalga_gpu.so.
"""
from __future__ import annotations

import os

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
    """One rank of the sharded overlap-graph build for equal-length reads (one process per GPU, one box).

    Exchange design (alga_gpu.h, ``alga_ps_shard_*``): nothing is routed on the host and no NCCL collective carries
    graph data.  Each rank owns two peer-mapped buffers (``torch.distributed._symmetric_memory``): its shard of the
    packed reads and an exchange workspace.  Per build

        1. the other ranks' read shards are pulled over NVLink on a copy stream, one peer at a time, while the main
           stream inserts the shards that already arrived into the (replicated) seed index;
        2. phase 1 appends every edge to the segment of the rank owning its target, in the rank's own workspace;
        3. phase 2 reads the segments addressed to it straight out of all workspaces (NVLink loads in the kernel) and
           appends the surviving edges to the segment of the rank owning their source;
        4. the CSR stage reads those segments the same way.

    Three stream-ordered barriers on the symmetric-memory signal pads separate the stages.
    """

    def __init__(self, min_overlap, rs_min_overlap, min_offset, max_len_cap, device, rank, world, len_nt, n_shard, words_per_read,
                 group=None, n_total=None, bucket_load=None, seed_keys=None):
        import torch.distributed._symmetric_memory as symm

        self.rank, self.world = rank, world
        self.group = group if group is not None else dist.group.WORLD
        self.device = device
        self.len_nt, self.n_shard, self.W = len_nt, n_shard, words_per_read
        # rank r owns the reads [r * n_shard, min((r + 1) * n_shard, n_total)): the last ranks may own fewer (or none)
        self.n_total = n_shard * world if n_total is None else int(n_total)
        assert (world - 1) * n_shard < self.n_total <= world * n_shard or self.n_total <= n_shard, (n_shard, world, n_total)
        self._err = None
        self.plan = PrefSufPlan(min_overlap, rs_min_overlap, min_offset, max_len_cap, device=device)
        # the table slices travel over NVLink: from 4 ranks on denser tables (6 instead of 3 entries per bucket on average,
        # half the bytes) pay for the longer bucket chains
        if bucket_load is None:
            bucket_load = 6 if world >= 4 else 0
        self.plan.lib.alga_ps_set_bucket_load(int(bucket_load))
        # peer-mapped buffers
        self.shard_sym = symm.empty(n_shard * words_per_read, dtype=torch.int32, device=device)
        self.ws_sym = symm.empty(self.plan.shard_ws_bytes(n_shard, world), dtype=torch.uint8, device=device)
        self.ws_sym.zero_()
        self._h_shard = symm.rendezvous(self.shard_sym, self.group)
        self._h_ws = symm.rendezvous(self.ws_sym, self.group)
        self._peer_shards = [self._h_shard.get_buffer(p, (n_shard * words_per_read,), torch.int32) for p in range(world)]
        self._compact = torch.empty(n_shard * world * words_per_read, dtype=torch.int32, device=device)  # landing zone of the pulls
        # seed tables: every rank fills its slice of the bucket space, the slices are then copied from each other
        tb = self.plan.shard_table_bytes(self.n_total, world)
        self._slice_bytes = tb // world
        self.table_bytes = tb  # one seed table (bench.py: NVLink accounting)
        self.tp_sym = symm.empty(tb, dtype=torch.uint8, device=device)
        self.ts_sym = symm.empty(tb, dtype=torch.uint8, device=device)
        self._h_tp = symm.rendezvous(self.tp_sym, self.group)
        self._h_ts = symm.rendezvous(self.ts_sym, self.group)
        self._peer_tp = [self._h_tp.get_buffer(p, (tb,), torch.uint8) for p in range(world)]
        self._peer_ts = [self._h_ts.get_buffer(p, (tb,), torch.uint8) for p in range(world)]
        self._shard = PrefSufPlan.shard_struct(rank, world, n_shard, self.n_total, list(self._h_ws.buffer_ptrs),
                                               self.tp_sym.data_ptr(), self.ts_sym.data_ptr())
        # replicated read set + its binding (no pass over the reads: they arrive during the build)
        # (the shards travel in the caller's compact layout; the replicated copy has one sector-aligned slot per read, which is
        # what the fast kernels want: common.cuh aligned_stride_words)
        self.S = (words_per_read + 7) & ~7
        self._full = torch.zeros(n_shard * world * self.S + READ_PAD_BYTES // 4, dtype=torch.int32, device=device)
        self._slots = self._full[: n_shard * world * self.S].view(n_shard * world, self.S)
        self._len = torch.full((self.n_total,), len_nt, dtype=torch.int32, device=device)
        self._reads = DeviceReads.from_tensors(self._full, self._len, stride=self.S, n=self.n_total, max_len=len_nt)
        self.plan.bind_uniform(self._reads, len_nt)
        # seed records: every rank computes bucket and tag of the two seeds of ITS reads once and the peers pull the 12-byte records
        # with the shard, instead of every rank deriving the minimizers of all reads for its slice (world times the work).
        # Measured (config 4, r2): 54.6 against 54.9 ms on 2 GPUs, 19.0 against 18.0 ms on 8 -- the stage is bound by the pulls, and the
        # records add a third to them -- so it is OFF unless asked for (ALGA_SHARD_SEED_KEYS=1).
        if seed_keys is None:
            seed_keys = os.environ.get("ALGA_SHARD_SEED_KEYS") == "1"
        self.seed_keys = bool(seed_keys)
        # ALGA_SHARD_PREFIX_FIRST=1: the prefix slices are finished and exchanged first, the suffix slices filled meanwhile in a pass of
        # their own (phase 1 needs the prefix table only).  Measured: 54.65 against 54.68 ms on 2 GPUs, 18.4 against 17.9 on 8 -- the
        # second pass over the reads costs what the earlier exchange saves -- so it is off.
        self.prefix_first = os.environ.get("ALGA_SHARD_PREFIX_FIRST", "0") == "1"
        if self.seed_keys:
            self.keys_sym = symm.empty(n_shard * 3, dtype=torch.int32, device=device)
            self._h_keys = symm.rendezvous(self.keys_sym, self.group)
            self._peer_keys = [self._h_keys.get_buffer(p, (n_shard * 3,), torch.int32) for p in range(world)]
            self._keys = torch.empty(world * n_shard * 3, dtype=torch.int32, device=device)  # landing zone of the pulls
        # pulls over NVLink are DMA copies on a copy stream, one at a time.  Cutting every pull into chunks on several streams
        # (several copy engines) was measured and is SLOWER when all GPUs pull at once: 8 GPUs, config 4: 18.0 ms with one stream,
        # 19.5 with two, 22.4 with four (2 GPUs: 54.5 / 55.8 with four).  ALGA_SHARD_COPY_STREAMS=<n> for experiments.
        self._copy_streams = [torch.cuda.Stream(device=device) for _ in range(max(1, int(os.environ.get("ALGA_SHARD_COPY_STREAMS", "1"))))]
        self._ev = None
        self._launches = 0
        torch.cuda.synchronize(device)
        self._h_ws.barrier()

    def load_shard(self, shard_words: torch.Tensor):
        """Put this rank's packed reads ([n_shard, W] int32, device or pinned host) into its peer-visible buffer."""
        flat = shard_words.reshape(-1)
        self.shard_sym[: flat.numel()].copy_(flat, non_blocking=True)

    def run(self):
        """One build over the shards currently in the ranks' peer-visible buffers."""
        main = torch.cuda.current_stream(self.device)
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        n, W = self.n_shard, self.W
        marks[0].record()
        if self.seed_keys:
            self._h_shard.barrier()  # nobody still reads the previous build's records
            lo_own, hi_own = min(self.rank * n, self.n_total), min((self.rank + 1) * n, self.n_total)
            self._stage(lambda: self.plan.shard_seed_keys(self._shard, self.shard_sym, W, hi_own - lo_own, self.keys_sym))
        self._h_shard.barrier()  # every rank's shard (and its seed records) in place, and nobody still reads the previous one
        for st in self._copy_streams:
            st.wait_stream(main)
        pulls = []
        for k in range(self.world):  # all pulls are queued at once, peer by peer; the main stream follows them
            p = (self.rank + k) % self.world
            evs = []
            if self.seed_keys:
                evs += self._pull(self._keys[p * n * 3:(p + 1) * n * 3], self._peer_keys[p])
            evs += self._pull(self._compact[p * n * W:(p + 1) * n * W], self._peer_shards[p])  # DMA over NVLink, contiguous on both sides
            pulls.append(evs)
        for k in range(self.world):
            p = (self.rank + k) % self.world
            stage = self._compact[p * n * W:(p + 1) * n * W]
            for ev in pulls[k]:
                main.wait_event(ev)
            self._slots[p * n:(p + 1) * n, :W].copy_(stage.view(n, W))  # local: into the sector-aligned read slots
            # seeds of the arrived shard that fall into this rank's slice of the bucket space
            lo_p, hi_p = min(p * n, self.n_total), min((p + 1) * n, self.n_total)
            if self.seed_keys:
                self._stage(lambda: self.plan.shard_index_keys(self._shard, self._keys[p * n * 3:], lo_p, hi_p, first=(k == 0)))
            else:  # prefix_first: the prefix slice only -- the suffix slice is filled while the prefix slices travel
                self._stage(lambda: self.plan.shard_index_range(self._shard, lo_p, hi_p, first=(k == 0),
                                                                 which=1 if self.prefix_first else 0))
        sb = self._slice_bytes
        ev_tp, ev_ts = [], []

        def exchange(tabs, mine, evs):
            ev_done = torch.cuda.Event()
            ev_done.record(main)
            for st in self._copy_streams:
                st.wait_event(ev_done)
            for k in range(1, self.world):
                p = (self.rank + k) % self.world
                evs += self._pull(mine[p * sb:(p + 1) * sb], tabs[p][p * sb:(p + 1) * sb])

        self._h_ws.barrier()  # every rank's (prefix) slice is complete
        if self.prefix_first and not self.seed_keys:
            exchange(self._peer_tp, self.tp_sym, ev_tp)
            self._stage(lambda: self.plan.shard_index_range(self._shard, 0, self.n_total, first=True, which=2))
            self._h_ws.barrier()  # every rank's suffix slice is complete
            marks[1].record()
            exchange(self._peer_ts, self.ts_sym, ev_ts)
        else:
            marks[1].record()
            exchange(self._peer_tp, self.tp_sym, ev_tp)
            exchange(self._peer_ts, self.ts_sym, ev_ts)
        for ev in ev_tp:
            main.wait_event(ev)  # the suffix table keeps arriving while phase 1 runs
        self._stage(lambda: self.plan.shard_phase1(self._shard))
        marks[2].record()
        self._h_ws.barrier()
        for ev in ev_ts:
            main.wait_event(ev)
        self._stage(lambda: self.plan.shard_phase2(self._shard))
        marks[3].record()
        self._h_ws.barrier()
        marks[4].record()
        self._stage(lambda: self.plan.shard_csr(self._shard))
        marks[5].record()
        self._ev = marks
        self._launches = self.plan.stats()["kernel_launches"]
        self._raise_together()

    def _pull(self, dst: torch.Tensor, src: torch.Tensor) -> list:
        """dst <- src (1-D, the same length, src in a peer's memory), in chunks spread over the copy streams; -> their events."""
        total = dst.numel()
        c = len(self._copy_streams)
        step = ((total + c - 1) // c + 1023) & ~1023
        evs = []
        for i, st in enumerate(self._copy_streams):
            lo, hi = i * step, min(total, (i + 1) * step)
            if lo >= hi:
                break
            with torch.cuda.stream(st):
                dst[lo:hi].copy_(src[lo:hi], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
            evs.append(ev)
        return evs

    def _stage(self, fn):
        """A stage that fails on ONE rank (e.g. ALGA_E_CAPACITY) must not leave the others waiting at the next barrier:
        the error is kept, the rank walks through the remaining barriers without computing, and ``_raise_together``
        makes every rank raise after the last one."""
        if self._err is None:
            try:
                fn()
            except Exception as e:  # noqa: BLE001
                self._err = e

    def _raise_together(self):
        flag = torch.tensor([0 if self._err is None else 1], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=self.group)
        if int(flag.item()):
            err, self._err = self._err, None
            if err is not None:
                raise err
            raise RuntimeError("sharded build failed on another rank (the build may be run again: plans grow their buffers)")

    def stats(self) -> dict:
        names = ("gather+index", "slices+phase1", "barrier+pull+phase2", "barrier", "pull+csr")
        self._ev[-1].synchronize()
        ms = {k: self._ev[i].elapsed_time(self._ev[i + 1]) for i, k in enumerate(names)}
        st = self.plan.stats()
        ms["pull_rows_kernel"] = st["stage_ms"]["transpose"]
        ms["phase2_kernels"] = st["stage_ms"]["phase2"]
        return {"kernel_launches": self._launches, "stage_ms": ms, "n_spilled_targets": st["n_spilled_targets"],
                "n_row_overflow": st["n_row_overflow"], "n_hard_sources": st["n_hard_sources"]}

    def total_edges(self) -> int:
        t = torch.tensor([self.plan.n_edges()], dtype=torch.int64, device=self.device)
        dist.all_reduce(t, group=self.group)
        return int(t.item())

    def result_device(self):
        return self.plan.result_device()

    def e2e(self, host_shard: torch.Tensor, steps: int, n_nodes_total: int) -> dict:
        """Same build with HOST buffers: pinned shard in, this rank's CSR rows out (pinned), wall clock, max over ranks."""
        host_in = host_shard.cpu().pin_memory()
        self.load_shard(host_in)
        self.run()
        ro, nb, of = self.result_device()
        cap = int(nb.numel() * 1.1) + 1024
        h_ro = torch.empty(ro.numel(), dtype=ro.dtype).pin_memory()
        h_nb = torch.empty(cap, dtype=nb.dtype).pin_memory()
        h_of = torch.empty(cap, dtype=of.dtype).pin_memory()
        d2h = 0

        def one():
            nonlocal d2h
            self.load_shard(host_in)
            self.run()
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
                "call": "alga_b200.distributed.ShardedPrefSuf.load_shard(pinned host) + run() + pinned host CSR rows out"}

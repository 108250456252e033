#!/usr/bin/env python
"""Benchmark of the overlap-graph hot path (GraphCreatorPrefSuf + retainOnlySmallestOffset).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg4|cfg1|cfg2|cfg3|cfg5]

A "step" is one complete overlap-graph build over one synthetic read set: packed reads resident in HBM -> CSR adjacency
resident in HBM.  Metric: graph nodes (strand-reads) per second.

* workload: BASELINE.json configs[3] by default -- the configuration north_star states its targets on (100 Mbp genome,
  2x150 bp, 50x, error-free, seed 4: 56.7 M strand-reads, fits one GPU) -- at every N: N = 1 is the whole read set on one
  GPU, N > 1 the SAME read set split into contiguous read-id ranges ("scaling": "strong", north_star's 1/2/4/8 sweep).
  The read set comes from the counter-based generator alga_b200/synth_dev.py (SURVEY.md 8-d distributions, integer only,
  identical on CPU and GPU), run on each rank's own GPU in a few seconds.
* parity: after the timed loop the graph the LAST timed step left in HBM is digested on the device (alga_b200/edge_hash.py,
  order-independent 128-bit sum over (source, target, offset); per-rank digests of disjoint row ranges add up) and compared
  with the digest of the graph the UNMODIFIED reference built from the same read set (tests/golden/full_<workload>_<gen>.json,
  written by tests/golden/make_full_golden.py through oracle/_ref).  "parity": true | false | null (no golden for this
  workload / scale); false makes the run exit non-zero.
* --impl reference : the reference's own CPU GraphCreatorPrefSuf (oracle/_ref/alga_ref_harness = the unmodified
  reference sources behind a file-reading main; else the plain-C oracle port), all host threads, on a bounded sample
  (genome-scaled: same read length and coverage) of the same workload; the sample is named in cpu_baseline.sample.

One JSON line on stdout (rank 0).  See DESIGN.md "Measurement" for the definition of every key.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "reads/sec overlap-graph build"
UNIT = "reads/s"


# --------------------------------------------------------------------------------------------------------
def alg_bytes_per_node(len_nt: int, lmin: int, edges_per_node: float) -> float:
    """B_alg = 4 W + 64 n_L + 8 E/N + 8."""
    W = (len_nt + 15) // 16
    n_l = len_nt - lmin + 1
    return 4 * W + 64 * n_l + 8 * edges_per_node + 8


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks line of B200_PROFILING.md, sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(prefix="alga_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self) -> dict:
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], [], set(), []
        try:
            for line in open(self.path):
                f = [x.strip() for x in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                except ValueError:
                    continue
                for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons), samples=len(sm),
                       power_w_max=float(max(power)))
        return out


# --------------------------------------------------------------------------------------------------------
# workloads
def golden_for(workload: str, gen: str, scale: float):
    """The reference's own result for this read set (tests/golden/full_*.json), or None."""
    tag = workload if scale == 1.0 else f"{workload}_at_{scale:g}"
    try:
        with open(os.path.join(ROOT, "tests", "golden", f"full_{tag}_{gen}.json")) as f:
            return json.load(f)
    except Exception:
        return None


def host_workload(workload: str, gen: str, scale: float):
    """-> (ReadSet, params, name, records, genome_size) on the host (CPU legs, small scales)."""
    from alga_b200 import synth, synth_dev

    if gen == "np":
        w = synth.make_config(workload, scale)
        return w.reads, w.params, w.name, w.records, w.genome_size
    w = synth_dev.make_config(workload, scale)
    return w.to_readset(), w.params, w.name, w.records, w.genome_size


# --------------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline: the reference's CPU graph creator on a bounded sample
def _cpu_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def _cpu_build(reads, params, threads):
    """-> (seconds of the replaced region main.cpp:282-291, kind, threads used)."""
    from oracle import harness, oracle  # the CPU legs are the only place bench.py touches oracle/

    if harness.available():
        _, info = harness.run_prefsuf(reads, params.min_overlap, params.rs_min_overlap, params.min_offset,
                                      threads=threads, want_edges=False)
        return float(info["graph_s"]), "reference", threads
    t = time.perf_counter()
    oracle.prefsuf(reads, params.min_overlap, params.rs_min_overlap, params.min_offset)
    return time.perf_counter() - t, "port", 1


def cpu_sample(workload: str, gen: str, target_s: float):
    """Pick a genome scale so that one CPU build of the sample takes about target_s; -> (host workload tuple, calib)."""
    from alga_b200 import synth

    threads = _cpu_threads()
    gsize = synth.CONFIGS[workload]["genome_size"]
    s0 = min(1.0, 50_000 / gsize)
    probe = host_workload(workload, gen, s0)
    s, kind, used = _cpu_build(probe[0], probe[1], threads)
    rate = probe[0].n / max(s, 1e-3)  # nodes/s on the tiny probe (pessimistic for the threaded reference)
    full_nodes = probe[0].n / s0
    scale = min(1.0, max(s0, rate * target_s / full_nodes))
    return host_workload(workload, gen, scale), dict(kind=kind, threads=used, scale=scale)


def _sample_text(hw, cal) -> str:
    reads, _params, name, records, gsize = hw
    return (f"{name}: genome scale {cal['scale']:.4g} of the named workload = {reads.n} nodes ({records} records, genome {gsize} bp, "
            f"same read length and coverage), region main.cpp:282-291 timed by steady_clock inside the harness")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    total = args.steps + args.warmup
    target = min(10.0, max(1.0, 150.0 / max(total, 1)))
    hw, cal = cpu_sample(args.workload, args.gen, target)
    times = []
    for i in range(total):
        s, kind, used = _cpu_build(hw[0], hw[1], cal["threads"])
        if i >= args.warmup:
            times.append(s)
    ms = 1e3 * float(np.mean(times))
    value = hw[0].n / (ms / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64/u32 (polynomial hashes mod 1e18+3, 1e9+7)", "data": "synthetic",
        "config": workload_config(args.workload, args.gpus, args.scale, args.gen),
        "sample_scale": cal["scale"], "sample_nodes": hw[0].n,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cal["threads"], "kind": cal["kind"], "sample": _sample_text(hw, cal)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def workload_config(workload: str, n_gpus: int, scale: float = 1.0, gen: str = "dev") -> dict:
    from alga_b200 import synth

    kw = synth.CONFIGS[workload]
    desc = (f"{workload}{'' if scale == 1.0 else f' at genome scale {scale:g}'}: synthetic {kw['genome_size'] * scale / 1e6:g} Mbp random genome, "
            f"{'2x' if kw['paired'] else ''}{kw['read_len']} bp {'paired' if kw['paired'] else 'single-end'} reads at "
            f"{kw['coverage']}x, error {kw.get('error', 0.0):g}, seed {kw['seed']}, --error_rate=0 (GraphCreatorPrefSuf only); generator "
            f"{'alga_b200/synth_dev.py (counter-based)' if gen == 'dev' else 'alga_b200/synth.py (NumPy default_rng)'}")
    if n_gpus > 1:
        desc += f"; strong scaling: the same read set, contiguous read-id ranges over {n_gpus} ranks"
    return {"workload": desc, "l2": "explicit flush (256 MiB write) between timed steps; step inputs+index also exceed L2",
            "sharding": "1 GPU" if n_gpus == 1 else f"read-id ranges over {n_gpus} GPUs"}


# --------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from alga_b200 import _lib, edge_hash, synth, synth_dev
    from alga_b200.plan import READ_PAD_BYTES, DeviceReads, PrefSufPlan

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the overlap-graph path has no CPU fallback")
    _lib.load()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # ---- the read set, generated on this rank's GPU (every rank generates the same one) ----------------------
    t0 = time.time()
    if args.gen == "dev":
        dw = synth_dev.make_config(args.workload, args.scale, device=dev)
        words2d, len_nt, params, n_records = dw.words, dw.len_nt, dw.params, dw.records
        del dw
    else:
        w = synth.make_config(args.workload, args.scale)
        assert int(w.reads.len_nt.min()) == int(w.reads.len_nt.max()), "bench.py runs equal-length read sets"
        W_ = int(w.reads.word_off[1] - w.reads.word_off[0])
        words2d = torch.from_numpy(w.reads.words.view(np.int32).reshape(w.reads.n, W_)).to(dev)
        len_nt, params, n_records = int(w.reads.len_nt[0]), w.params, w.records
    torch.cuda.synchronize(dev)
    gen_s = time.time() - t0
    n_nodes_total, W = int(words2d.shape[0]), int(words2d.shape[1])
    torch.cuda.empty_cache()

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    clocks = ClockSampler(local)

    if world == 1:
        flat = torch.zeros(n_nodes_total * W + READ_PAD_BYTES // 4, dtype=torch.int32, device=dev)
        flat[: n_nodes_total * W].copy_(words2d.reshape(-1))
        lens = torch.full((n_nodes_total,), len_nt, dtype=torch.int32, device=dev)
        dreads = DeviceReads.from_tensors(flat, lens, stride=W, n=n_nodes_total, max_len=len_nt)
        host_words = words2d.cpu().numpy().view(np.uint32).reshape(-1) if args.e2e_steps != 0 else None  # for the e2e leg
        del words2d
        torch.cuda.empty_cache()
        plan = PrefSufPlan(params.min_overlap, params.rs_min_overlap, params.min_offset, params.max_len_cap, device=dev)
        plan.bind(dreads)
        row_lo = 0

        def step():
            plan.run()

        def step_stats():
            return plan.stats()

        def result_device():
            return plan.result_device()
    else:
        from alga_b200.distributed import ShardedPrefSuf

        per = ((n_nodes_total + world - 1) // world + 1) & ~1  # reads per rank (even: twins stay together); the last rank has fewer
        row_lo = min(rank * per, n_nodes_total)
        shard_words = words2d[row_lo:min(row_lo + per, n_nodes_total)].clone()
        del words2d
        torch.cuda.empty_cache()
        sp = ShardedPrefSuf(params.min_overlap, params.rs_min_overlap, params.min_offset, params.max_len_cap, dev,
                            rank, world, len_nt=len_nt, n_shard=per, words_per_read=W, n_total=n_nodes_total)
        sp.load_shard(shard_words)  # the rank's packed reads, resident in its peer-visible HBM buffer

        def step():
            sp.run()

        def step_stats():
            return sp.stats()

        def result_device():
            return sp.result_device()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    stage_acc, launches, diag = {}, 0, {}
    clocks.start()
    barrier()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(2)  # L2 flush, outside the per-step events
        if world > 1:
            dist.barrier()
        a.record()
        step()
        b.record()
        st = step_stats()
        launches += int(st["kernel_launches"])
        diag = {k: st[k] for k in ("n_spilled_targets", "n_row_overflow", "n_hard_sources") if k in st}
        for k, v in st["stage_ms"].items():
            stage_acc[k] = stage_acc.get(k, 0.0) + float(v)
    barrier()
    wall_s = time.perf_counter() - wall0
    clk = clocks.stop()
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = n_nodes_total / (ms_per_step / 1e3)

    # ---- parity gate: digest of the graph the last timed step built, against the reference's own graph --------
    ro, nb, of = result_device()
    dg = edge_hash.digest_csr(ro, nb, of, first_row=row_lo)
    n_edges = int(nb.numel())
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, (dg, n_edges))
        dg, n_edges = (0, 0), 0
        for d_, e_ in parts:
            dg, n_edges = edge_hash.add(dg, d_), n_edges + e_
    digest = [f"{dg[0]:016x}", f"{dg[1]:016x}"]
    gold = golden_for(args.workload, args.gen, args.scale)
    parity, parity_note = None, "no reference golden for this workload / scale (tests/golden/make_full_golden.py makes one)"
    selfcheck = None  # config 5: no reference run at that size; the digest of the generic kernels' graph instead (NOT a parity claim)
    if gold is None:
        try:
            tag = args.workload if args.scale == 1.0 else f"{args.workload}_at_{args.scale:g}"
            with open(os.path.join(ROOT, "tests", "golden", f"selfcheck_{tag}_{args.gen}.json")) as f:
                sc = json.load(f)
            selfcheck = bool(sc["nodes"] == n_nodes_total and sc["edges"] == n_edges and sc["digest"] == digest)
            parity_note += "; selfcheck = the same digest as the generic kernels' graph (tests/golden/selfcheck_*.json)"
        except Exception:
            pass
    if gold is not None:
        parity = bool(gold["nodes"] == n_nodes_total and gold["edges"] == n_edges and gold["digest"] == digest)
        parity_note = (f"edge-set digest of the last timed step vs the unmodified reference at --threads={gold['threads']} "
                       f"(tests/golden/full_*.json: {gold['nodes']} nodes, {gold['edges']} edges)")

    # ---- e2e: the reference-facing call (GraphCreatorPrefSuf over alga_gpu_prefsuf_build) with host buffers ----
    e2e = None
    if world == 1 and host_words is not None:
        from alga_b200.graph_creator import GraphCreatorPrefSuf
        from alga_b200.readset import ReadSet

        hreads = ReadSet(host_words, np.arange(n_nodes_total + 1, dtype=np.uint64) * np.uint64(W),
                         np.full(n_nodes_total, len_nt, np.uint32))
        # borrow=True: the result stays in the library's page-locked staging, as the C++ shim reads it
        gc = GraphCreatorPrefSuf(hreads, params.min_overlap, params.rs_min_overlap, params.min_offset,
                                 params.max_len_cap, device=local, pinned=True, borrow=True)
        del hreads, host_words
        for _ in range(2):
            gc.startAlignmentGraphCreation()
        torch.cuda.synchronize(dev)
        k_e2e = args.e2e_steps if args.e2e_steps > 0 else max(1, min(args.steps, 10 if n_nodes_total < 10_000_000 else 5))
        ts = time.perf_counter()
        for _ in range(k_e2e):
            g = gc.startAlignmentGraphCreation()
        torch.cuda.synchronize(dev)
        e2e_ms = 1e3 * (time.perf_counter() - ts) / k_e2e
        assert g.n_edges == n_edges, (g.n_edges, n_edges)
        r_ = gc.reads
        h2d = int(r_.words.nbytes + r_.len_nt.nbytes + gc.alignFrom.nbytes + gc.alignTo.nbytes)
        d2h = int(g.row_off.nbytes + g.nbr.nbytes + g.off.nbytes)
        e2e = {"value": n_nodes_total / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": k_e2e,
               "call": "alga_b200.GraphCreatorPrefSuf.startAlignmentGraphCreation -> alga_gpu_prefsuf_build (host buffers)",
               "timing": gc.timing}
    elif world > 1:
        e2e = sp.e2e(shard_words, steps=max(1, min(args.steps, 5)), n_nodes_total=n_nodes_total)

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return 0 if parity is not False else 3

    peak, peak_src = measured_peak_gbs()
    b_alg = alg_bytes_per_node(len_nt, params.min_overlap, n_edges / max(n_nodes_total, 1))
    nodes_per_gpu = n_nodes_total / world
    achieved = nodes_per_gpu * b_alg / (ms_per_step / 1e3) / 1e9
    stages = {k: v / args.steps for k, v in stage_acc.items()}
    traffic = measured_traffic(args.workload, args.scale, world)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (traffic or {}).get("pipeline_dram_bytes_per_step"), "peak_source": peak_src,
                "kernel": "whole device pipeline of one step (index + phase1 + phase2 + csr), per GPU",
                "alg_bytes_per_node": b_alg, "nodes_per_gpu": nodes_per_gpu, "stage_ms": stages, "diag": diag,
                "traffic_detail": traffic}
    if world > 1:
        # what crosses NVLink per rank and step, from the sizes of what is pulled (alga_b200/distributed.py): the other ranks' read
        # shards, their slices of the two seed tables, the phase-1 edges addressed to this rank (24 B) and the surviving edges
        # whose source it owns (12 B).  `achieved` divides by the WHOLE step (transfers overlap the kernels), so it is a lower
        # bound of the rate on the wire.
        rem = (world - 1) / world
        nv_bytes = (n_nodes_total * (W * 4 + (12 if sp.seed_keys else 0)) * rem + 2 * sp.table_bytes * rem + 3 * nodes_per_gpu * 24 * rem + (n_edges / world) * 12 * rem)
        roofline["nvlink"] = {"bytes_per_step_per_rank": nv_bytes, "achieved_gbs": nv_bytes / (ms_per_step / 1e3) / 1e9, "peak_gbs": 900.0,
                              "frac": nv_bytes / (ms_per_step / 1e3) / 1e9 / 900.0,
                              "note": "read shards + seed-table slices + exchanged edges pulled by one rank, over the whole step time"}
    if world == 1 and stages.get("phase2"):
        # the dominant kernels on their own: algorithmic bytes of their share of the overlap lengths (SURVEY.md 8-d:
        # 64 B per node and length, + the CSR output for phase 2) over their CUDA-event time inside this run
        rs_, lmin_ = params.rs_min_overlap, params.min_overlap
        n_l2 = max(0, len_nt - max(rs_, lmin_) + 1)
        n_l1 = max(0, min(rs_, len_nt + 1) - lmin_)
        b2 = 64 * n_l2 + 8 * n_edges / max(n_nodes_total, 1) + 8
        b1 = 64 * n_l1
        k2 = nodes_per_gpu * b2 / (stages["phase2"] / 1e3) / 1e9
        k1 = nodes_per_gpu * b1 / (stages["phase1"] / 1e3) / 1e9
        share2, share1 = stages["phase2"] / ms_per_step, stages["phase1"] / ms_per_step
        roofline["dominant_kernel"] = {"kernel": f"phase 2 (+ its spill kernels; {100 * share2:.0f} % of the step)", "bound": "hbm",
                                       "achieved": k2, "peak": peak, "unit": "GB/s", "frac": k2 / peak,
                                       "alg_bytes_per_node": b2, "ms": stages["phase2"],
                                       "traffic": (traffic or {}).get("phase2")}
        roofline["second_kernel"] = {"kernel": f"phase 1 (+ its queue kernel; {100 * share1:.0f} % of the step)", "bound": "hbm",
                                     "achieved": k1, "peak": peak, "unit": "GB/s", "frac": k1 / peak,
                                     "alg_bytes_per_node": b1, "ms": stages["phase1"],
                                     "traffic": (traffic or {}).get("phase1")}

    cpu = None
    if world == 1 and not args.no_cpu:
        hw, cal = cpu_sample(args.workload, args.gen, 12.0)
        s, kind, used = _cpu_build(hw[0], hw[1], cal["threads"])
        cpu = {"value": hw[0].n / s, "unit": UNIT, "cores": used, "kind": kind, "seconds": s, "sample": _sample_text(hw, cal)}

    legs_ok = world == 1 and args.gen == "np"
    if (args.with_input or args.with_preprocess or args.with_triangles or args.workload == "cfg3") and not legs_ok:
        print("bench.py: the optional legs (--with-*, supplement) need --gen np on one GPU; skipped", file=sys.stderr)
    supplement = None
    if legs_ok and args.workload == "cfg3":
        # BASELINE configs[2]: the error-rate supplement (main.cpp:300-355) on top of the graph just built, host to host
        from alga_b200.graph_creator import GraphCreatorLI
        from alga_b200.graph_creator import supplement_params

        sp = supplement_params(float(w.reads.len_nt.mean()))
        li = GraphCreatorLI(w.reads, g, **sp, device=local)
        ts = time.perf_counter()
        g2 = li.startAlignmentGraphCreation()
        supplement = {"ms": 1e3 * (time.perf_counter() - ts), "edges_before": int(g.n_edges), "edges_after": int(g2.n_edges),
                      "params": sp, **li.timing,
                      "note": "alga_gpu_supplement: LI k-mers, pair enumeration and canAlign on the GPU, k-mer sort on the GPU (tied buckets re-sorted on the host), ordered replay on the host"}

    triangles = None
    if legs_ok and args.with_triangles:
        # the first simplifier step (sortEdgesByIncreasingOffset + cutNonAndWeaklyMetricTriangles) on the
        # graph just built (after the supplement where it ran), host to host
        from alga_b200.graph_creator import GraphSimplifier

        gin = g2 if supplement else g
        LEN = int(float(w.reads.len_nt[w.reads.len_nt > 0].mean())) + 2 * synth.TRIM
        mopp = max(250, int(1.75 * LEN))  # Params::MAX_OFFSET_PARALLEL_PATHS, main.cpp:95
        edges_in = gin.edges().copy()
        n_in = int(gin.n_edges)
        GraphSimplifier(gin, mopp, device=local).cutNonAndWeaklyMetricTriangles()  # warm-up
        gs = GraphSimplifier(gin, mopp, device=local)
        ts = time.perf_counter()
        gout = gs.cutNonAndWeaklyMetricTriangles()
        triangles = {"ms": 1e3 * (time.perf_counter() - ts), "edges_before": n_in, "edges_after": int(gout.n_edges),
                     "max_offset_parallel_paths": mopp, **gs.timing, "call": "alga_gpu_cut_triangles (host buffers)"}
        from oracle import harness as _h
        if _h.available() and not args.no_cpu:
            t0c = time.perf_counter()
            ref = _h.run_cut_triangles(edges_in, w.reads.n, mopp, threads=_cpu_threads())
            triangles["cpu_reference"] = {"seconds_incl_io": time.perf_counter() - t0c, "cores": _cpu_threads(), "edges_after": int(ref.shape[0]),
                                          "note": "oracle/_ref harness: the reference's own GraphSimplifier step on the same graph, file IO included"}

    preprocess = None
    if legs_ok and args.with_preprocess:
        # ReadPreprocess::getPrefixReads on the strand-reads BEFORE duplicate removal, host to host
        from alga_b200 import readset as _rs
        from alga_b200.graph_creator import ReadPreprocess

        kwp = dict(synth.CONFIGS[args.workload])
        kwp["genome_size"] = max(20_000, int(kwp["genome_size"] * args.scale))
        rngp = np.random.default_rng(kwp["seed"])
        gen = synth.make_genome(kwp["genome_size"], rngp, repeats=kwp.get("repeats", 0))
        if kwp["paired"]:
            m1, m2 = synth.sample_paired_end(gen, kwp["read_len"], kwp["coverage"], rngp, kwp.get("error", 0.0))
            raw = _rs.from_code_matrix(synth.strand_nodes(m1, m2))
        else:
            raw = _rs.from_code_matrix(synth.strand_nodes(synth.sample_single_end(gen, kwp["read_len"], kwp["coverage"], rngp,
                                                                                   kwp.get("error", 0.0))))
        rp = ReadPreprocess(raw, device=local)
        rp.getPrefixReads(2)
        ts = time.perf_counter()
        mask = rp.getPrefixReads(2)
        ms = 1e3 * (time.perf_counter() - ts)
        preprocess = {"ms": ms, "reads": raw.n, "removed": int(mask.sum()), **rp.timing,
                      "reads_per_s": raw.n / (ms / 1e3), "call": "alga_gpu_prefix_reads (host buffers)"}
        from oracle import harness as _h
        if _h.available() and not args.no_cpu:
            sub = _rs.ReadSet(raw.words[: (raw.n // 8) * int(raw.word_off[1])], raw.word_off[: raw.n // 8 + 1], raw.len_nt[: raw.n // 8])
            t0c = time.perf_counter()
            _h.run_prefix_reads(sub, 2, threads=_cpu_threads())
            preprocess["cpu_reference"] = {"reads": sub.n, "seconds_incl_io": time.perf_counter() - t0c, "cores": _cpu_threads(),
                                           "note": "oracle/_ref harness: ReadPreprocess::getPrefixReads on 1/8 of the reads, file IO included"}

    input_leg = None
    if legs_ok and args.with_input:
        # the files of the workload through InputReader::readInput on the GPU (alga_gpu_read_input), and
        # the whole driver path main.cpp:82-291 (reader, prefix-read removal, renumbering, GraphCreatorPrefSuf), host to host
        from alga_b200.input_reader import FASTA, InputReader, build_overlap_graph

        kwi = dict(synth.CONFIGS[args.workload])
        kwi["genome_size"] = max(20_000, int(kwi["genome_size"] * args.scale))
        rngi = np.random.default_rng(kwi["seed"])
        gen = synth.make_genome(kwi["genome_size"], rngi, repeats=kwi.get("repeats", 0))
        if kwi["paired"]:
            m1, m2 = synth.sample_paired_end(gen, kwi["read_len"], kwi["coverage"], rngi, kwi.get("error", 0.0))
            t1, t2 = synth.fasta_text(m1), synth.fasta_text(m2)
        else:
            m1 = synth.sample_single_end(gen, kwi["read_len"], kwi["coverage"], rngi, kwi.get("error", 0.0))
            t1, t2 = synth.fasta_text(m1), None
        rd = InputReader(FASTA, device=local)
        rd.readInput(t1, t2)
        ts = time.perf_counter()
        rs_in = rd.readInput(t1, t2)
        ms = 1e3 * (time.perf_counter() - ts)
        n_rec = m1.shape[0] * (2 if t2 is not None else 1)
        nbytes = len(t1) + (len(t2) if t2 is not None else 0)
        input_leg = {"ms": ms, "records": n_rec, "reads": rs_in.n, "file_bytes": nbytes, **rd.timing, **rd.info,
                     "records_per_s": n_rec / (ms / 1e3), "call": "alga_gpu_read_input (host buffers)"}
        from alga_b200.input_reader import PinnedText
        p1, p2 = PinnedText(t1), (PinnedText(t2) if t2 is not None else None)  # the files read into page-locked memory
        for _ in range(2):
            build_overlap_graph(p1, p2, FASTA, device=local)
        best = None
        for _ in range(3):
            ts = time.perf_counter()
            og = build_overlap_graph(p1, p2, FASTA, device=local)
            msg = 1e3 * (time.perf_counter() - ts)
            if best is None or og.timing["total_ms"] < best[1].timing["total_ms"]:
                best = (msg, og)
        msg, og = best
        if args.workload != "cfg3" and args.scale == 1.0:  # same generator and seed as the main workload: the same graph must come out
            assert (og.reads.n, og.graph.n_edges) == (n_nodes_total, n_edges), (og.reads.n, og.graph.n_edges, n_nodes_total, n_edges)
        input_leg["files_to_graph"] = {"ms_python": msg, "nodes": og.reads.n, "edges": og.graph.n_edges, "params": og.params,
                                       **og.timing, "records_per_s": n_rec / (og.timing["total_ms"] / 1e3),
                                       "h2d_bytes": nbytes, "d2h_bytes": int(og.reads.words.nbytes + og.reads.len_nt.nbytes * 2 + og.reads.n
                                                                          + og.graph.n_edges * 8 + (og.reads.n + 1) * 8),
                                       "call": "alga_gpu_files_to_graph (main.cpp:82-291 in one call, file text in page-locked memory)"}
        og2 = build_overlap_graph(t1, t2, FASTA, device=local)
        input_leg["files_to_graph_pageable"] = {"total_ms": og2.timing["total_ms"], "h2d_ms": og2.timing["h2d_ms"],
                                                "note": "same call, file text in ordinary (pageable) memory: upload through page-locked chunks"}
        from oracle import harness as _h
        if _h.available() and not args.no_cpu:
            k = max(1, m1.shape[0] // 8)
            s1 = synth.fasta_text(m1[:k])
            s2 = synth.fasta_text(m2[:k]) if t2 is not None else None
            t0c = time.perf_counter()
            _h.run_read_input(s1, s2, 1, threads=_cpu_threads())
            input_leg["cpu_reference"] = {"records": k * (2 if t2 is not None else 1), "seconds_incl_io": time.perf_counter() - t0c,
                                          "cores": _cpu_threads(),
                                          "note": "oracle/_ref harness: InputReader::readInput on 1/8 of the records, file IO included"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u64 (2-bit packed words, exact compare)", "data": "synthetic",
        "config": workload_config(args.workload, world, args.scale, args.gen),
        "nodes": n_nodes_total, "records": n_records, "edges": n_edges, "parity": parity, "selfcheck": selfcheck, "parity_note": parity_note,
        "edge_digest": digest, "gen_s": gen_s, "wall_s_timed_region": wall_s,
        "clocks": clk, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
    }
    if supplement:
        line["supplement"] = supplement
    if triangles:
        line["triangles"] = triangles
    if preprocess:
        line["preprocess"] = preprocess
    if input_leg:
        line["input"] = input_leg
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if parity is False:
        print("bench.py: PARITY FAILED -- the edge set differs from the reference's", file=sys.stderr)
        return 3
    return 0


def measured_traffic(workload: str, scale: float, world: int):
    """ncu DRAM bytes of one step (profiles/traffic_r2.json), only if it was captured for this very workload and N."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic_r2.json")) as f:
            t = json.load(f)
        if t.get("workload") == workload and float(t.get("scale", 1.0)) == float(scale) and int(t.get("n_gpus", 1)) == world:
            return t
    except Exception:
        pass
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--gen", default="dev", choices=["dev", "np"],
                    help="read-set generator: dev = synth_dev (counter-based, on the GPU), np = synth (NumPy, host)")
    ap.add_argument("--scale", type=float, default=1.0, help="genome scale of the GPU workload (1.0 = the named config)")
    ap.add_argument("--e2e-steps", type=int, default=-1, help="timed steps of the host-buffer leg (0 = skip it)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--with-preprocess", action="store_true",
                    help="also time ReadPreprocess::getPrefixReads (alga_gpu_prefix_reads) on the reads before dedupe")
    ap.add_argument("--with-triangles", action="store_true",
                    help="also time the first simplifier step (alga_gpu_cut_triangles) on the graph just built")
    ap.add_argument("--with-input", action="store_true",
                    help="also time InputReader::readInput (alga_gpu_read_input) and the files-to-graph path on FASTA text of the workload")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = max(args.warmup, 1)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())

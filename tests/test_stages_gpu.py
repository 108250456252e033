"""GPU parity of the remaining entry points of the C ABI: fingerprints, 2-bit packing, candidate verification, the
device-resident plan, the staged (sharded) pipeline emulated rank by rank on one GPU, and -- at BASELINE.json's
full config-2 size -- size-independent properties of the result."""
import numpy as np
import pytest

from alga_b200 import readset, synth
from alga_b200.graph_creator import GraphCreatorPrefSuf, fingerprints, pack_reads, verify_pairs
from oracle import oracle
from tests.cases import build_case, verify_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,L", [("tiny", 1), ("tiny", 33), ("varlen", 55), ("varlen", 90), ("long_reads", 400),
                                    ("cfg1_small", 94)])
def test_fingerprints_match_oracle(gpu, name, L):
    """updatePrefixHash / updateSuffixHash (GraphCreatorPrefSuf.cpp:213-236): bit-exact u64 / u32 values."""
    rs, *_ = build_case(name)
    want = oracle.fingerprints(rs, L)
    got = fingerprints(rs, L)
    for g, w in zip(got, want):
        assert np.array_equal(g, w)


@pytest.mark.parametrize("n,ln", [(1, 1), (7, 16), (1000, 94), (513, 144), (33, 1000), (3, 16384)])
def test_pack_reads_matches_host_packing(gpu, n, ln):
    """Read::createSequence (Read.cpp:40-68) incl. non-ACGT -> 0 and lower case -> 0."""
    rng = np.random.default_rng(n * 1000 + ln)
    codes = rng.integers(0, 4, size=(n, ln), dtype=np.uint8)
    ascii_ = np.frombuffer(b"ACGT", np.uint8)[codes].copy()
    junk = rng.random(codes.shape) < 0.02
    ascii_[junk] = rng.choice(np.frombuffer(b"Nacgt-", np.uint8), size=int(junk.sum()))
    codes[junk] = 0
    assert np.array_equal(pack_reads(ascii_), readset.pack_matrix(codes))


def test_verify_pairs_banded_lcs_matches_reference(gpu):
    """alga_gpu_verify_pairs with lcs_rate_pct > 0 (USE_ACLER_INSTEAD_OF_ACLCS = 0): the banded LCS of AlignmentControllerLCS on
    the pairs the low-error test rejects, against the verdicts of the unmodified reference and the oracle."""
    import os

    from tests.cases import LCS_SETTINGS

    rs, pairs, vp = verify_case()
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "verify_pairs_lcs.npz"))
    for rate, band in LCS_SETTINGS:
        got = verify_pairs(rs, pairs, lcs_rate_pct=rate, lcs_band=band, **vp)
        assert np.array_equal(got, z[f"v_{rate}_{band}"]), (rate, band)
        assert np.array_equal(got, oracle.verify_pairs(rs, pairs, lcs_rate_pct=rate, lcs_band=band, **vp))
    with pytest.raises(Exception):
        verify_pairs(rs, pairs[:10], lcs_rate_pct=95, lcs_band=9, **vp)


def test_verify_pairs_matches_oracle(gpu):
    rs, pairs, vp = verify_case()
    want = oracle.verify_pairs(rs, pairs, **vp)
    got = verify_pairs(rs, pairs, **vp)
    assert np.array_equal(got, want)
    assert 0 < int(got.sum()) < got.shape[0]


@pytest.mark.parametrize("thr,max_off,area", [(99, 10, 130), (90, 60, 60), (100, 32, 111)])
def test_verify_pairs_parameter_sweep(gpu, thr, max_off, area):
    rs, pairs, _ = verify_case(seed=32, n_reads=1500)
    vp = dict(threshold_pct=thr, max_offset_pct=max_off, min_overlap_area=area, min_offset=0)
    assert np.array_equal(verify_pairs(rs, pairs, **vp), oracle.verify_pairs(rs, pairs, **vp))


@pytest.mark.parametrize("name", ["cfg2_small", "varlen_dups", "flags", "nulls", "long_reads_rs"])
def test_device_resident_plan_matches_oracle(gpu, name):
    import torch

    from alga_b200.plan import DeviceReads, PrefSufPlan

    rs, lmin, rsmin, mo = build_case(name)
    plan = PrefSufPlan(lmin, rsmin, mo, device=0)
    plan.bind(DeviceReads(rs, torch.device("cuda", 0)))
    for _ in range(2):  # the workspace is reused across runs
        plan.run()
        assert np.array_equal(plan.result_host().edges(), oracle.prefsuf(rs, lmin, rsmin, mo))
    ro, nb, of = plan.result_device()
    assert ro.shape[0] == rs.n + 1 and int(ro[-1].item()) == nb.shape[0] == of.shape[0] == plan.n_edges()
    plan.close()


@pytest.mark.parametrize("name", ["cfg2_small", "cfg3_small", "varlen_dups", "periodic_dups", "flags"])
@pytest.mark.parametrize("world", [2, 3])
def test_staged_pipeline_emulated_ranks(gpu, name, world):
    """The sharded build of alga_b200/distributed.py with the ranks emulated one after another on one GPU:
    phase 1 per source range -> route by target -> phase 2 per target range -> route by source -> CSR per range."""
    import torch

    from alga_b200.plan import DeviceReads, PrefSufPlan

    rs, lmin, rsmin, mo = build_case(name)
    dev = torch.device("cuda", 0)
    plan = PrefSufPlan(lmin, rsmin, mo, device=0)
    plan.bind(DeviceReads(rs, dev))
    plan.stage_index()
    bounds = [rs.n * r // world for r in range(world + 1)]
    t1 = torch.cat([plan.stage_phase1(bounds[r], bounds[r + 1]).clone() for r in range(world)])
    t2 = []
    for r in range(world):
        sel = (t1[:, 1] >= bounds[r]) & (t1[:, 1] < bounds[r + 1])
        t2.append(plan.stage_phase2(bounds[r], bounds[r + 1], t1[sel]).clone())
    t2 = torch.cat(t2)
    edges = []
    for r in range(world):
        sel = (t2[:, 0] >= bounds[r]) & (t2[:, 0] < bounds[r + 1])
        plan.stage_csr(bounds[r], bounds[r + 1], t2[sel])
        e = plan.result_host().edges()
        e[:, 0] += bounds[r]
        edges.append(e)
    got = np.concatenate(edges) if edges else np.zeros((0, 3), np.int32)
    assert np.array_equal(got, oracle.prefsuf(rs, lmin, rsmin, mo))
    plan.close()


def test_full_config2_properties(gpu):
    """BASELINE.json configs[1] at full size (2.6 M nodes): properties that do not need the CPU oracle.
    * rows sorted by target, one entry per target (retainOnlySmallestOffset, Graph.cpp:348-387)
    * every edge is an exact suffix/prefix overlap >= min overlap (Global::checkOLCGraphCorrectness, Global.cpp:121-145)
    * strand symmetry: b -> c with overlap L  <=>  c^1 -> b^1 with overlap L (both strands of every read are nodes)
    * run-to-run determinism (checksum of the CSR)"""
    w = synth.make_config("cfg2")
    rs, p = w.reads, w.params
    gc = GraphCreatorPrefSuf(rs, p.min_overlap, p.rs_min_overlap)
    g = gc.startAlignmentGraphCreation()
    e = g.edges()
    first = (g.row_off.copy(), g.nbr.copy(), g.off.copy())  # the result arrays are borrowed until the next build
    assert e.shape[0] > rs.n * 0.9
    same_row = e[1:, 0] == e[:-1, 0]
    assert np.all(e[1:, 1][same_row] > e[:-1, 1][same_row])
    ln = rs.len_nt.astype(np.int64)
    L = ln[e[:, 0]] - e[:, 2]
    assert np.all(L >= p.min_overlap) and np.all(e[:, 2] >= 0) and np.all(e[:, 0] != e[:, 1])
    rng = np.random.default_rng(0)
    W = int(rs.word_off[1])
    words = rs.words.reshape(rs.n, W)
    for k in rng.integers(0, e.shape[0], size=3000):
        b, c, o = (int(x) for x in e[k])
        sb = _codes(words[b], int(ln[b]))
        sc = _codes(words[c], int(ln[c]))
        assert np.array_equal(sb[o:], sc[: ln[b] - o])
    # equal-length reads: overlap L of (b, c) equals overlap of (c^1, b^1); phase-2 survivors are symmetric, the
    # phase-1 "last 3" rule is not, so only compare edges whose overlap is >= rs
    big = e[L >= p.rs_min_overlap]
    fwd = set(map(tuple, big.tolist()))
    mirrored = set((int(c) ^ 1, int(b) ^ 1, int(o)) for b, c, o in big.tolist())
    assert fwd == mirrored
    g2 = GraphCreatorPrefSuf(rs, p.min_overlap, p.rs_min_overlap).startAlignmentGraphCreation()
    assert all(np.array_equal(a, b) for a, b in zip((g2.row_off, g2.nbr, g2.off), first))


def _codes(words, ln):
    j = np.arange(ln)
    return (words[j >> 4] >> ((j & 15) * 2).astype(np.uint32)) & 3


@pytest.mark.parametrize("name", ["cfg2_small", "cfg5_small", "varlen_dups", "periodic_dups", "rs_above_maxl", "flags"])
@pytest.mark.parametrize("row_cap", [1, 3])
def test_row_overflow_paths_match_oracle(gpu, monkeypatch, name, row_cap):
    """Phase 1 writes into fixed-capacity rows of the transposed graph.  A tiny capacity (ALGA_PS_ROW_CAP, read when
    the plan is created) pushes edges into the overflow list: short lists are scanned by the generic phase-2 kernel,
    long ones (> 2048 entries) trigger the CSR rebuild on the device."""
    import torch

    from alga_b200.plan import DeviceReads, PrefSufPlan

    monkeypatch.setenv("ALGA_PS_ROW_CAP", str(row_cap))
    rs, lmin, rsmin, mo = build_case(name)
    plan = PrefSufPlan(lmin, rsmin, mo, device=0)
    plan.bind(DeviceReads(rs, torch.device("cuda", 0)))
    plan.run()
    st = plan.stats()
    assert st["n_row_overflow"] > 0
    assert np.array_equal(plan.result_host().edges(), oracle.prefsuf(rs, lmin, rsmin, mo))
    plan.close()


def test_row_overflow_list_regrows(gpu, monkeypatch):
    """More overflow entries than the overflow list holds: the build reruns with a worst-case list."""
    import torch

    from alga_b200.plan import DeviceReads, PrefSufPlan

    monkeypatch.setenv("ALGA_PS_ROW_CAP", "1")
    w = synth.make_config("cfg2", scale=0.05)  # ~130 k nodes, ~250 k overflow entries > n / 8 + 65536
    plan = PrefSufPlan(w.params.min_overlap, w.params.rs_min_overlap, 0, device=0)
    plan.bind(DeviceReads(w.reads, torch.device("cuda", 0)))
    plan.run()
    assert plan.stats()["n_row_overflow"] > w.reads.n // 8 + 65536
    assert np.array_equal(plan.result_host().edges(), oracle.prefsuf(w.reads, w.params.min_overlap, w.params.rs_min_overlap))
    plan.close()


@pytest.mark.parametrize("name", ["cfg2_small", "cfg1_small", "cfg5_small"])
@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("use_keys", [False, True, "split"])  # "split": prefix slices first, then the suffix slices in one pass
def test_sharded_peer_exchange_emulated_ranks(gpu, name, world, use_keys):
    """alga_ps_shard_* with the ranks emulated one after another on one GPU: every rank has its own exchange workspace
    (an ordinary device buffer here, peer-mapped symmetric memory in alga_b200/distributed.py), phase 1 appends to the
    owner's segment, phase 2 and the CSR stage read the segments addressed to them out of all workspaces."""
    import torch

    from alga_b200.plan import DeviceReads, PrefSufPlan

    rs, lmin, rsmin, mo = build_case(name)
    dev = torch.device("cuda", 0)
    n = rs.n
    n_shard = (n + world - 1) // world
    plan = PrefSufPlan(lmin, rsmin, mo, device=0)
    dr = DeviceReads(rs, dev)
    dr.align_from = dr.align_to = None
    plan.bind_uniform(dr, int(rs.len_nt[0]))
    ws = [torch.zeros(plan.shard_ws_bytes(n_shard, world), dtype=torch.uint8, device=dev) for _ in range(world)]
    tb = plan.shard_table_bytes(n, world)
    tp = [torch.zeros(tb, dtype=torch.uint8, device=dev) for _ in range(world)]
    ts = [torch.zeros(tb, dtype=torch.uint8, device=dev) for _ in range(world)]
    shards = [plan.shard_struct(r, world, n_shard, n, [w.data_ptr() for w in ws], tp[r].data_ptr(), ts[r].data_ptr())
              for r in range(world)]
    bounds = [min(n, r * n_shard) for r in range(world + 1)]
    if use_keys is True:  # the seeds of every read computed once, by its owner (12-byte records), and inserted from the records
        W = dr.stride
        keys = [torch.zeros(max(1, bounds[q + 1] - bounds[q]) * 3, dtype=torch.int32, device=dev) for q in range(world)]
        for q in range(world):
            plan.shard_seed_keys(shards[q], dr.words[bounds[q] * W:], W, bounds[q + 1] - bounds[q], keys[q])
    for r in range(world):  # every rank fills its slice of the bucket space, the reads arriving in pieces
        for q in range(world):
            if use_keys is True:
                plan.shard_index_keys(shards[r], keys[q], bounds[q], bounds[q + 1], first=(q == 0))
            elif use_keys == "split":
                plan.shard_index_range(shards[r], bounds[q], bounds[q + 1], first=(q == 0), which=1)
            else:
                plan.shard_index_range(shards[r], bounds[q], bounds[q + 1], first=(q == 0))
        if use_keys == "split":
            plan.shard_index_range(shards[r], 0, n, first=True, which=2)
    sb = tb // world
    for r in range(world):  # ... and copies the other ranks' slices
        for q in range(world):
            if q != r:
                tp[r][q * sb:(q + 1) * sb] = tp[q][q * sb:(q + 1) * sb]
                ts[r][q * sb:(q + 1) * sb] = ts[q][q * sb:(q + 1) * sb]
    for r in range(world):
        plan.shard_phase1(shards[r])
    for r in range(world):
        plan.shard_phase2(shards[r])
    edges = []
    for r in range(world):
        plan.shard_csr(shards[r])
        e = plan.result_host().edges()
        e[:, 0] += bounds[r]
        edges.append(e)
    assert np.array_equal(np.concatenate(edges), oracle.prefsuf(rs, lmin, rsmin, mo))
    plan.close()

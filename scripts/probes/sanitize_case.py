"""Small end-to-end runs for compute-sanitizer (memcheck / racecheck / initcheck): fast path, generic path, sharded
path with emulated ranks, supplement."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from alga_b200.graph_creator import GraphCreatorPrefSuf, GraphCreatorLI
from alga_b200.plan import DeviceReads, PrefSufPlan
from oracle import oracle
from tests.cases import build_case, supplement_case

for name in ("cfg2_small", "varlen_dups", "periodic_dups", "long_reads"):
    rs, lmin, rsmin, mo = build_case(name)
    want = oracle.prefsuf(rs, lmin, rsmin, mo)
    g = GraphCreatorPrefSuf(rs, lmin, rsmin, mo).startAlignmentGraphCreation()
    assert np.array_equal(g.edges(), want), name
    print("ok", name, flush=True)
rs, lmin, rsmin, mo = build_case("cfg2_small")
dev = torch.device("cuda", 0)
world, n = 3, rs.n
n_shard = (n + world - 1) // world
plan = PrefSufPlan(lmin, rsmin, mo, device=0)
dr = DeviceReads(rs, dev); dr.align_from = dr.align_to = None
plan.bind_uniform(dr, int(rs.len_nt[0]))
ws = [torch.zeros(plan.shard_ws_bytes(n_shard, world), dtype=torch.uint8, device=dev) for _ in range(world)]
tb = plan.shard_table_bytes(n, world)
tp = [torch.zeros(tb, dtype=torch.uint8, device=dev) for _ in range(world)]
ts = [torch.zeros(tb, dtype=torch.uint8, device=dev) for _ in range(world)]
sh = [plan.shard_struct(r, world, n_shard, n, [w.data_ptr() for w in ws], tp[r].data_ptr(), ts[r].data_ptr()) for r in range(world)]
bounds = [min(n, r * n_shard) for r in range(world + 1)]
for r in range(world):
    for q in range(world):
        plan.shard_index_range(sh[r], bounds[q], bounds[q + 1], first=(q == 0))
sb = tb // world
for r in range(world):
    for q in range(world):
        if q != r:
            tp[r][q * sb:(q + 1) * sb] = tp[q][q * sb:(q + 1) * sb]; ts[r][q * sb:(q + 1) * sb] = ts[q][q * sb:(q + 1) * sb]
for r in range(world): plan.shard_phase1(sh[r])
for r in range(world): plan.shard_phase2(sh[r])
edges = []
for r in range(world):
    plan.shard_csr(sh[r]); e = plan.result_host().edges(); e[:, 0] += bounds[r]; edges.append(e)
assert np.array_equal(np.concatenate(edges), oracle.prefsuf(rs, lmin, rsmin, mo))
print("ok sharded", flush=True)
rs, lmin, rsmin, sp = supplement_case("sup_varlen")
g0 = GraphCreatorPrefSuf(rs, lmin, rsmin).startAlignmentGraphCreation()
g1 = GraphCreatorLI(rs, g0, **sp).startAlignmentGraphCreation()
assert np.array_equal(g1.edges(), oracle.supplement(rs, g0.edges(), **sp))
print("ok supplement", flush=True)

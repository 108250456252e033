"""One GPU, the sharded code path with world = 1 (plain device buffers as 'peer' memory): stage timings of the
segment / pull kernels without any NVLink in the way.  python scripts/probes/shard_emul_bench.py [scale]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from alga_b200 import synth
from alga_b200.plan import DeviceReads, PrefSufPlan

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
w = synth.make_config("cfg2", scale=scale)
dev = torch.device("cuda", 0)
rs = w.reads
n = rs.n
plan = PrefSufPlan(w.params.min_overlap, w.params.rs_min_overlap, 0, device=0)
dr = DeviceReads(rs, dev); dr.align_from = dr.align_to = None
plan.bind_uniform(dr, int(rs.len_nt[0]))
ws = torch.zeros(plan.shard_ws_bytes(n, 1), dtype=torch.uint8, device=dev)
tb = plan.shard_table_bytes(n, 1)
tp = torch.zeros(tb, dtype=torch.uint8, device=dev); ts = torch.zeros(tb, dtype=torch.uint8, device=dev)
sh = plan.shard_struct(0, 1, n, n, [ws.data_ptr()], tp.data_ptr(), ts.data_ptr())
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
names = ["index", "phase1", "pull_rows+phase2", "pull+csr"]
acc = {k: 0.0 for k in names}; acc["pull_rows_kernel"] = 0.0; acc["phase2_kernels"] = 0.0
steps = 6
for it in range(steps + 2):
    flush.fill_(it)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record(); plan.shard_index_range(sh, 0, n, True)
    ev[1].record(); plan.shard_phase1(sh)
    ev[2].record(); plan.shard_phase2(sh)
    ev[3].record(); plan.shard_csr(sh)
    ev[4].record(); torch.cuda.synchronize()
    if it >= 2:
        for i, k in enumerate(names): acc[k] += ev[i].elapsed_time(ev[i + 1]) / steps
        st = plan.stats()["stage_ms"]; acc["pull_rows_kernel"] += st["transpose"] / steps; acc["phase2_kernels"] += st["phase2"] / steps
print({k: round(v, 3) for k, v in acc.items()}, "total", round(sum(acc[k] for k in names), 3), "edges", plan.n_edges())

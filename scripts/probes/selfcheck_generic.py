"""Full-size self-check where no reference golden exists (config 5: 250 Mbp with repeats, 2 x 100 bp, 40x -- the reference would need
the better part of an hour): the graph of the fast path (tile kernels, sorted index and CSR) against the graph of the GENERIC kernels,
which replay GraphCreatorPrefSuf.cpp:356-488 literally and are pinned against the oracle on every small case.  Order-independent digest
of (source, target, offset) + node and edge counts.

    python scripts/probes/selfcheck_generic.py [workload] [scale]
"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from alga_b200 import edge_hash, synth_dev  # noqa: E402
from alga_b200.plan import READ_PAD_BYTES, DeviceReads, PrefSufPlan  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
dev = torch.device("cuda", 0)
t0 = time.time()
dw = synth_dev.make_config(workload, scale, device=dev)
n, W, len_nt, params = int(dw.words.shape[0]), int(dw.words.shape[1]), dw.len_nt, dw.params
flat = torch.zeros(n * W + READ_PAD_BYTES // 4, dtype=torch.int32, device=dev)
flat[: n * W].copy_(dw.words.reshape(-1))
del dw
lens = torch.full((n,), len_nt, dtype=torch.int32, device=dev)
reads = DeviceReads.from_tensors(flat, lens, stride=W, n=n, max_len=len_nt)
torch.cuda.synchronize()
out = {"workload": workload, "scale": scale, "nodes": n, "gen_s": round(time.time() - t0, 2)}
for name, generic in (("fast", False), ("generic", True)):
    plan = PrefSufPlan(params.min_overlap, params.rs_min_overlap, params.min_offset, params.max_len_cap, device=dev, force_generic=generic)
    plan.bind(reads)
    plan.run()
    if not generic:
        plan.run()  # warm
    st = plan.stats()
    ro, nb, of = plan.result_device()
    d = edge_hash.digest_csr(ro, nb, of)
    out[name] = {"edges": int(nb.numel()), "digest": [f"{d[0]:016x}", f"{d[1]:016x}"], "device_ms": round(st["device_ms"], 2),
                 "stage_ms": {k: round(v, 2) for k, v in st["stage_ms"].items()}, "n_spilled_targets": st.get("n_spilled_targets"),
                 "n_hard_sources": st.get("n_hard_sources")}
    del ro, nb, of
    plan.close()
    torch.cuda.empty_cache()
out["match"] = out["fast"]["edges"] == out["generic"]["edges"] and out["fast"]["digest"] == out["generic"]["digest"]
print(json.dumps(out))
sys.exit(0 if out["match"] else 3)

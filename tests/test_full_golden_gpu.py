"""GPU parity at FULL size: the CUDA path against the graph the UNMODIFIED reference built for the whole BASELINE
configs (tests/golden/full_*.json, made by tests/golden/make_full_golden.py through oracle/_ref)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from alga_b200 import edge_hash, synth, synth_dev
from alga_b200.graph_creator import GraphCreatorPrefSuf
from alga_b200.plan import READ_PAD_BYTES, DeviceReads, PrefSufPlan

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _golden(tag):
    p = os.path.join(HERE, "golden", f"full_{tag}.json")
    if not os.path.exists(p):
        pytest.skip(f"no golden {tag}")
    return json.load(open(p))


@pytest.mark.parametrize("workload", ["cfg1", "cfg2", "cfg3"])
@pytest.mark.parametrize("gen", ["dev", "np"])
def test_host_call_matches_reference_golden(gpu, workload, gen):
    """alga_gpu_prefsuf_build (host buffers) on full configs 1-3: node count, edge count, digest and SHA-256 of the sorted
    (source, target, offset) array equal the reference's."""
    g = _golden(f"{workload}_{gen}")
    if gen == "dev":
        w = synth_dev.make_config(workload, device="cuda")
        rs, p = w.to_readset(), w.params
    else:
        w = synth.make_config(workload)
        rs, p = w.reads, w.params
    assert hashlib.sha256(np.ascontiguousarray(rs.words).tobytes()).hexdigest() == g["input_sha"], "generator drift"
    gr = GraphCreatorPrefSuf(rs, p.min_overlap, p.rs_min_overlap).startAlignmentGraphCreation()
    e = gr.edges()
    assert (rs.n, e.shape[0]) == (g["nodes"], g["edges"])
    assert [f"{x:016x}" for x in edge_hash.digest_edges(e)] == g["digest"]
    assert hashlib.sha256(np.ascontiguousarray(e).tobytes()).hexdigest() == g["sha256"]


def test_device_plan_matches_reference_golden_cfg4(gpu):
    """BASELINE config 4 (100 Mbp, 56.7 M strand-reads), device-resident plan: digest of the CSR in HBM against the
    reference's graph for the same read set."""
    g = _golden("cfg4_dev")
    dev = torch.device("cuda", 0)
    w = synth_dev.make_config("cfg4", device=dev)
    n, W = w.words.shape
    assert n == g["nodes"]
    flat = torch.zeros(n * W + READ_PAD_BYTES // 4, dtype=torch.int32, device=dev)
    flat[: n * W].copy_(w.words.reshape(-1))
    lens = torch.full((n,), w.len_nt, dtype=torch.int32, device=dev)
    plan = PrefSufPlan(w.params.min_overlap, w.params.rs_min_overlap, device=dev)
    plan.bind(DeviceReads.from_tensors(flat, lens, stride=W, n=n, max_len=w.len_nt))
    del w
    plan.run()
    ro, nb, of = plan.result_device()
    assert int(nb.numel()) == g["edges"]
    assert [f"{x:016x}" for x in edge_hash.digest_csr(ro, nb, of)] == g["digest"]
    plan.close()

"""alga_gpu_prefsuf_build_multi: the sharded build driven from ONE process (host thread per GPU, peer memory) -- what the
reference's single-process driver binds (shim: ALGA_GPU_DEVICES).  With one GPU on the box the call must fall back to it."""
import numpy as np
import pytest

from alga_b200 import _lib, synth
from alga_b200.graph_creator import GraphCreatorPrefSuf
from oracle import oracle
from tests.cases import build_case

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_gpus", [2, 8])
def test_multi_matches_oracle(gpu, n_gpus):
    w = synth.make_config("cfg2", scale=0.02)  # 52 k reads, not a multiple of the number of GPUs
    want = oracle.prefsuf(w.reads, w.params.min_overlap, w.params.rs_min_overlap)
    gc = GraphCreatorPrefSuf(w.reads, w.params.min_overlap, w.params.rs_min_overlap, n_gpus=n_gpus)
    for _ in range(2):
        got = gc.startAlignmentGraphCreation().edges()
        assert got.shape == want.shape and np.array_equal(got, want)
    used = int(gc.timing["n_gpus_used"]) if "n_gpus_used" in gc.timing else 0
    assert used in (0, min(n_gpus, _lib.load().alga_gpu_device_count()))


@pytest.mark.parametrize("name", ["varlen_dups", "flags", "nulls", "cfg1_small"])
def test_multi_falls_back_for_general_read_sets(gpu, name):
    """Ragged lengths, cleared flags, removed reads, small inputs: built on one GPU, same result."""
    rs, lmin, rsmin, mo = build_case(name)
    want = oracle.prefsuf(rs, lmin, rsmin, mo)
    got = GraphCreatorPrefSuf(rs, lmin, rsmin, mo, n_gpus=4).startAlignmentGraphCreation().edges()
    assert got.shape == want.shape and np.array_equal(got, want)

"""GPU parity of ReadPreprocess::getPrefixReads (the step right before the graph build) through the C ABI."""
import os

import numpy as np
import pytest

from alga_b200 import readset
from alga_b200.graph_creator import ReadPreprocess
from oracle import oracle
from tests.cases import PREPROCESS_CASES, build_case, preprocess_case

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.mark.parametrize("name", PREPROCESS_CASES)
@pytest.mark.parametrize("remove_type", [2, 1])
def test_prefix_reads_match_oracle_and_reference(gpu, name, remove_type):
    rs = preprocess_case(name)
    got = ReadPreprocess(rs).getPrefixReads(remove_type)
    assert np.array_equal(got, oracle.prefix_reads(rs, remove_type))
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    assert np.array_equal(got, z["mask_all" if remove_type == 2 else "mask_dup"])


@pytest.mark.parametrize("name", ["nulls", "long_reads", "tiny", "all_null", "periodic_dups", "short_lmin"])
def test_prefix_reads_edge_cases(gpu, name):
    """Removed (length 0) reads, reads of several thousand nucleotides, tiny and empty-ish sets."""
    rs, *_ = build_case(name)
    assert np.array_equal(ReadPreprocess(rs).getPrefixReads(2), oracle.prefix_reads(rs, 2))


def test_prefix_reads_idempotent_on_clean_set(gpu):
    """After removal (pairs dropped together, main.cpp:150-232) nothing is left to remove."""
    rs = preprocess_case("pre_equal")
    mask = ReadPreprocess(rs).getPrefixReads(2).astype(bool)
    drop = np.repeat(mask[0::2] | mask[1::2], 2)
    W = int(rs.word_off[1] - rs.word_off[0])
    words = rs.words.reshape(rs.n, W)[~drop]
    clean = readset.ReadSet(words.reshape(-1), np.arange(words.shape[0] + 1, dtype=np.uint64) * np.uint64(W),
                            rs.len_nt[~drop])
    assert ReadPreprocess(clean).getPrefixReads(2).sum() == 0

"""CPU tests: the oracle (plain-C restatement) against the golden fixtures produced by the UNMODIFIED
reference (tests/golden/make_golden.py), and -- when the compiled reference travels with the repo --
against the reference run live."""
import hashlib
import os

import numpy as np
import pytest

from oracle import harness, oracle
from tests.cases import (CASES, PREPROCESS_CASES, SUPPLEMENT_CASES, build_case, preprocess_case, supplement_case,
                         verify_case)

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def input_sha(rs) -> str:
    h = hashlib.sha256()
    for a in (rs.len_nt, rs.align_from, rs.align_to, rs.word_off, rs.words):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    return z


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    rs, lmin, rsmin, mo = build_case(name)
    z = load_golden(name)
    assert str(z["input_sha"]) == input_sha(rs), "seeded generator drifted: regenerate tests/golden"
    assert z["params"].tolist() == [lmin, rsmin, mo]
    got = oracle.prefsuf(rs, lmin, rsmin, mo)
    assert np.array_equal(got, z["edges"])


def test_oracle_verify_matches_reference_golden():
    rs, pairs, vp = verify_case()
    z = np.load(os.path.join(GOLDEN, "verify_pairs.npz"))
    assert str(z["input_sha"]) == input_sha(rs)
    assert np.array_equal(z["pairs"], pairs)
    got = oracle.verify_pairs(rs, pairs, **vp)
    assert np.array_equal(got, z["verdict"])
    assert 0 < got.sum() < got.shape[0]


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref/alga_ref_harness not present")
@pytest.mark.parametrize("name", ["cfg1_small", "cfg3_small", "varlen_dups", "periodic_dups", "long_reads"])
def test_oracle_matches_reference_live(name):
    rs, lmin, rsmin, mo = build_case(name)
    want, info = harness.run_prefsuf(rs, lmin, rsmin, mo, threads=1)
    assert np.array_equal(oracle.prefsuf(rs, lmin, rsmin, mo), want)
    assert info["edges"] == want.shape[0]


def test_fingerprint_definition():
    """oracle_fingerprints follows GraphCreatorPrefSuf.cpp:213-236: sum s_j 4^j mod (10^18+3, 10^9+7)."""
    rs, lmin, _, _ = build_case("tiny")
    for L in (1, 7, lmin, 60):
        p64, p32, s64, s32 = oracle.fingerprints(rs, L)
        for i in range(rs.n):
            if rs.len_nt[i] < L:
                continue
            c = [int(x) for x in rs.codes(i)]
            pre = sum(c[j] * 4 ** j for j in range(L))
            suf = sum(c[len(c) - L + j] * 4 ** j for j in range(L))
            assert int(p64[i]) == pre % (10 ** 18 + 3) and int(p32[i]) == pre % (10 ** 9 + 7)
            assert int(s64[i]) == suf % (10 ** 18 + 3) and int(s32[i]) == suf % (10 ** 9 + 7)


def test_edges_are_exact_overlaps():
    """Global::checkOLCGraphCorrectness (Global.cpp:121-145): every edge is an exact suffix/prefix overlap."""
    rs, lmin, rsmin, mo = build_case("varlen")
    e = oracle.prefsuf(rs, lmin, rsmin, mo)
    for b, c, o in e[:: max(1, e.shape[0] // 300)]:
        sb, sc = rs.sequence(int(b)), rs.sequence(int(c))
        L = len(sb) - int(o)
        assert L >= lmin and sb[int(o):] == sc[:L]


@pytest.mark.parametrize("name", SUPPLEMENT_CASES)
def test_oracle_supplement_matches_reference_golden(name):
    """Error-rate supplement (main.cpp:300-355): the C++ restatement against the graph the unmodified reference produced."""
    rs, lmin, rsmin, sp = supplement_case(name)
    z = load_golden(name)
    assert str(z["input_sha"]) == input_sha(rs), "seeded generator drifted: regenerate tests/golden"
    assert z["params"].tolist() == [lmin, rsmin, sp["threshold_pct"], sp["max_offset_pct"], sp["min_overlap_area"],
                                    sp["kmer_length_bucket"]]
    before = oracle.prefsuf(rs, lmin, rsmin, 0)
    assert np.array_equal(before, z["before"])
    after = oracle.supplement(rs, before, **sp)
    assert after.shape[0] > before.shape[0]
    assert np.array_equal(after, z["after"])


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref/alga_ref_harness not present")
def test_oracle_supplement_matches_reference_live():
    rs, lmin, rsmin, sp = supplement_case("sup_cfg3")
    before = oracle.prefsuf(rs, lmin, rsmin, 0)
    want, _ = harness.run_supplement(rs, before, sp["threshold_pct"], sp["max_offset_pct"], sp["min_overlap_area"],
                                     sp["kmer_length_bucket"], threads=1)
    assert np.array_equal(oracle.supplement(rs, before, **sp), want)


def test_li_kmer_definition():
    """oracle li_kmers follows Read.cpp:145-226: per interval the leftmost minimal K-mer under the priorities."""
    rs, *_ = supplement_case("sup_varlen")
    ids = np.flatnonzero(rs.len_nt >= 60)[:40].astype(np.uint32)
    K, IV = 35, 6
    for prio in ((0, 1, 2, 3), (1, 2, 3, 0), (3, 0, 1, 2)):
        h, ind = oracle.li_kmers(rs, ids, prio, K, IV)
        for q, i in enumerate(ids):
            c = [prio[int(x)] for x in rs.codes(int(i))]
            n_win = len(c) - K + 1
            ilen = -(-n_win // IV)
            vals = [sum(c[p + k] << (2 * (K - 1 - k)) for k in range(K)) for p in range(n_win)]
            for iv in range(IV):
                lo, hi = iv * ilen, min(n_win, (iv + 1) * ilen)
                if lo >= hi:
                    assert ind[q, iv] == -1
                    continue
                best = min(range(lo, hi), key=lambda p: (vals[p], p))
                assert ind[q, iv] == best and int(h[q, iv]) == vals[best] % (10 ** 18 + 3)


@pytest.mark.parametrize("name", PREPROCESS_CASES)
def test_oracle_prefix_reads_matches_reference_golden(name):
    """ReadPreprocess::getPrefixReads (ReadPreprocess.cpp:13-77), both removal types."""
    rs = preprocess_case(name)
    z = load_golden(name)
    assert str(z["input_sha"]) == input_sha(rs), "seeded generator drifted: regenerate tests/golden"
    assert np.array_equal(oracle.prefix_reads(rs, 2), z["mask_all"])
    assert np.array_equal(oracle.prefix_reads(rs, 1), z["mask_dup"])
    assert 0 < z["mask_dup"].sum() <= z["mask_all"].sum() < rs.n


@pytest.mark.skipif(not harness.available(), reason="oracle/_ref/alga_ref_harness not present")
def test_oracle_prefix_reads_matches_reference_live():
    rs = preprocess_case("pre_varlen")
    assert np.array_equal(oracle.prefix_reads(rs, 2), harness.run_prefix_reads(rs, 2))


def test_oracle_banded_lcs_matches_reference_golden():
    """USE_ACLER_INSTEAD_OF_ACLCS = 0: pairs the low-error test rejects go on to AlignmentControllerLCS (banded LCS)."""
    from tests.cases import LCS_SETTINGS

    rs, pairs, vp = verify_case()
    z = np.load(os.path.join(GOLDEN, "verify_pairs_lcs.npz"))
    base = oracle.verify_pairs(rs, pairs, **vp)
    for rate, band in LCS_SETTINGS:
        got = oracle.verify_pairs(rs, pairs, lcs_rate_pct=rate, lcs_band=band, **vp)
        assert np.array_equal(got, z[f"v_{rate}_{band}"]), (rate, band)
        assert (got >= base).all()  # the LCS only ever adds verdicts
    assert int(z["v_95_2"].sum()) > int(base.sum())

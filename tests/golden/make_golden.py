"""Generates the golden fixtures of tests/golden/ by running the UNMODIFIED reference.

Run in the dev container (needs oracle/_ref/alga_ref_harness, i.e. /root/reference compiled by
`make -C oracle ref`):   python tests/golden/make_golden.py

For every seeded case of tests/cases.py the reference's own GraphCreatorPrefSuf +
Graph::retainOnlySmallestOffset (main.cpp:282-291, --threads=1 order) is executed on the packed read
set and the edge set is stored as `<case>.npz`:
    edges      (E, 3) int32  (src, dst, offset) sorted
    input_sha  sha256 over len_nt | align_from | align_to | word_off | words  -- guards generator drift
    params     (min_overlap, rs_min_overlap, min_offset)
`verify_pairs.npz` holds (pairs, verdicts) of AlignmentControllerHybrid::canAlign evaluated by the
reference on candidate pairs of the cfg3_small read set; `sup_*.npz` hold the graph before and after the reference's
error-rate supplement (GraphCreatorLI, main.cpp:300-355) on the supplement cases of tests/cases.py; `pre_*.npz` hold the
removal masks of ReadPreprocess::getPrefixReads (both removal types) on the preprocessing cases; `in_*.npz` hold
Global::READS as the reference's InputReader::readInput leaves it (lengths, packed blocks) for the input files of the
input cases; `front_*.npz` hold the graph the STOCK binary serialises (--serialize=1, --threads=1) for the files of the
front cases, i.e. after its own reader, duplicate / prefix-read removal, renumbering and GraphCreatorPrefSuf, plus
Global::pairedReadOffset as the reference's own main() leaves it (harness mode `driver`); `tri_*.npz` hold
the edges that survive the reference's sortEdgesByIncreasingOffset + cutNonAndWeaklyMetricTriangles on graphs stored above.
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import harness  # noqa: E402
from tests.cases import (LCS_SETTINGS, CASES, FRONT_CASES, INPUT_CASES, PREPROCESS_CASES, SUPPLEMENT_CASES, TRIANGLE_CASES, build_case,  # noqa: E402
                         front_case, input_case, preprocess_case, supplement_case, triangle_case, verify_case)


def input_sha(rs) -> str:
    h = hashlib.sha256()
    for a in (rs.len_nt, rs.align_from, rs.align_to, rs.word_off, rs.words):
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def text_sha(t1, t2) -> str:
    h = hashlib.sha256(t1)
    if t2 is not None:
        h.update(b"|")
        h.update(t2)
    return h.hexdigest()


def main():
    assert harness.available(), "build the reference harness first: make -C oracle ref"
    for name in CASES:
        rs, lmin, rsmin, mo = build_case(name)
        if rs.n == 0:
            edges = np.zeros((0, 3), np.int32)  # the reference itself crashes on an empty read vector
        else:
            edges, _ = harness.run_prefsuf(rs, lmin, rsmin, mo, threads=1)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), edges=edges, input_sha=np.array(input_sha(rs)),
                            params=np.array([lmin, rsmin, mo], np.int32))
        print(f"{name}: n={rs.n} E={edges.shape[0]}")
    rs, pairs, vp = verify_case()
    verdict = harness.run_verify(rs, pairs, vp["threshold_pct"], vp["max_offset_pct"], vp["min_overlap_area"],
                                 vp["min_offset"])
    np.savez_compressed(os.path.join(HERE, "verify_pairs.npz"), pairs=pairs, verdict=verdict,
                        input_sha=np.array(input_sha(rs)))
    print(f"verify_pairs: {pairs.shape[0]} pairs, {int(verdict.sum())} accepted")
    # the same pairs with Params::USE_ACLER_INSTEAD_OF_ACLCS = 0: what the low-error test rejects goes on to the banded LCS
    # (AlignmentControllerLCS.cpp:30-150), for a few (MINIMAL_OVERLAP_RATE_FOR_LCS, MAX_ERROR_RATE_FOR_LCS) settings
    lcs = {}
    for rate, band in LCS_SETTINGS:
        lcs[f"v_{rate}_{band}"] = harness.run_verify(rs, pairs, vp["threshold_pct"], vp["max_offset_pct"], vp["min_overlap_area"],
                                                     vp["min_offset"], lcs_rate_pct=rate, lcs_band=band)
        print(f"verify_pairs_lcs rate {rate} band {band}: {int(lcs[f'v_{rate}_{band}'].sum())} accepted")
    np.savez_compressed(os.path.join(HERE, "verify_pairs_lcs.npz"), input_sha=np.array(input_sha(rs)), **lcs)
    # error-rate supplement (main.cpp:300-355, --threads=1) on top of the reference's own pre-supplement graph
    for name in SUPPLEMENT_CASES:
        rs, lmin, rsmin, sp = supplement_case(name)
        before, _ = harness.run_prefsuf(rs, lmin, rsmin, 0, threads=1)
        after, _ = harness.run_supplement(rs, before, sp["threshold_pct"], sp["max_offset_pct"], sp["min_overlap_area"],
                                          sp["kmer_length_bucket"], threads=1)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), before=before, after=after, input_sha=np.array(input_sha(rs)),
                            params=np.array([lmin, rsmin, sp["threshold_pct"], sp["max_offset_pct"], sp["min_overlap_area"],
                                             sp["kmer_length_bucket"]], np.int32))
        print(f"{name}: n={rs.n} E {before.shape[0]} -> {after.shape[0]}")
    # ReadPreprocess::getPrefixReads (main.cpp:132-134) on read sets that still hold duplicates and contained reads
    for name in PREPROCESS_CASES:
        rs = preprocess_case(name)
        m2 = harness.run_prefix_reads(rs, 2)
        m1 = harness.run_prefix_reads(rs, 1)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), mask_all=m2, mask_dup=m1, input_sha=np.array(input_sha(rs)))
        print(f"{name}: n={rs.n} removed {int(m2.sum())} (all prefix reads) / {int(m1.sum())} (duplicates only)")
    # InputReader::readInput (main.cpp:82) on file contents
    for name in INPUT_CASES:
        t1, t2, ft, extra = input_case(name)
        rs = harness.run_read_input(t1, t2, ft, extra=extra)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), len_nt=rs.len_nt, word_off=rs.word_off, words=rs.words,
                            input_sha=np.array(text_sha(t1, t2)))
        print(f"{name}: {rs.n} reads, {int((rs.len_nt == 0).sum())} nullptr")
    # the stock binary from the files to the serialised graph (main.cpp:57-293)
    for name in FRONT_CASES:
        t1, t2, ft = front_case(name)
        n, edges = harness.run_stock_graph(t1, t2, ft)
        po = harness.run_driver_paired_offsets(t1, t2, ft)  # Global::pairedReadOffset after the reference's own main()
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), n=np.array(n), edges=edges, paired_offset=po,
                            input_sha=np.array(text_sha(t1, t2)))
        print(f"{name}: n={n} E={edges.shape[0]}")
    # first simplifier step on the graphs stored above (the reference's own sortEdgesByIncreasingOffset + triangle cut)
    for name in TRIANGLE_CASES:
        e, n, mx = triangle_case(name, HERE)
        out = harness.run_cut_triangles(e, n, mx) if n else np.zeros((0, 3), np.int32)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), edges=out, n_in=np.array(e.shape[0]))
        print(f"{name}: {e.shape[0]} -> {out.shape[0]} edges")


if __name__ == "__main__":
    main()

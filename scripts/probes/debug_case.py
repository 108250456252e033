"""One parity case on the GPU with a diff against the oracle (development aid): python scripts/probes/debug_case.py <case> [generic]"""
import sys

import numpy as np

sys.path.insert(0, __file__.rsplit("/scripts/", 1)[0])
from alga_b200.graph_creator import GraphCreatorPrefSuf  # noqa: E402
from oracle import oracle  # noqa: E402
from tests.cases import build_case  # noqa: E402

name = sys.argv[1]
rs, lmin, rsmin, mo = build_case(name)
want = oracle.prefsuf(rs, lmin, rsmin, mo)
gc = GraphCreatorPrefSuf(rs, lmin, rsmin, mo, force_generic=len(sys.argv) > 2)
got = gc.startAlignmentGraphCreation().edges()
print(name, "n", rs.n, "lmin", lmin, "rs", rsmin, "want", want.shape[0], "got", got.shape[0], gc.timing)
ws = {tuple(x) for x in want.tolist()}
gs = {tuple(x) for x in got.tolist()}
miss, extra = sorted(ws - gs), sorted(gs - ws)
print("missing", len(miss), "extra", len(extra))
ln = rs.len_nt
for tag, lst in (("missing", miss), ("extra", extra)):
    for b, c, o in lst[:12]:
        L = int(ln[b]) - o
        print(f"  {tag}: {b} -> {c} offset {o}  L {L} ({'phase 1' if L < rsmin else 'phase 2'})  out-degree of {b}: want {sum(1 for e in want if e[0] == b)} got {sum(1 for e in got if e[0] == b)}")

"""Dev helper: run one config through the C ABI on the GPU and print timings (optionally check the oracle)."""
import argparse
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from alga_b200 import synth  # noqa: E402
from alga_b200.graph_creator import GraphCreatorPrefSuf  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cfg", default="cfg2")
ap.add_argument("--scale", type=float, default=1.0)
ap.add_argument("--check", action="store_true")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
t = time.time()
w = synth.make_config(a.cfg, a.scale)
print(f"{w.name}: nodes={w.reads.n} params={w.params} gen={time.time()-t:.1f}s", flush=True)
for r in range(a.reps):
    gc = GraphCreatorPrefSuf(w.reads, w.params.min_overlap, w.params.rs_min_overlap)
    t = time.time()
    g = gc.startAlignmentGraphCreation()
    print(f"rep {r}: E={g.n_edges} wall={1e3*(time.time()-t):.1f}ms timing={gc.timing}", flush=True)
if a.check:
    from oracle import oracle
    t = time.time()
    want = oracle.prefsuf(w.reads, w.params.min_overlap, w.params.rs_min_overlap)
    print(f"oracle: E={want.shape[0]} {time.time()-t:.1f}s equal={np.array_equal(want, g.edges())}")

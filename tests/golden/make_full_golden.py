"""Full-size golden results: the UNMODIFIED reference on the whole BASELINE configs.

    python tests/golden/make_full_golden.py <workload> <generator> [threads]      (dev container or GPU box)

workload  cfg1 .. cfg5, optionally `cfg4@0.25` (genome scale)
generator np  = alga_b200.synth      (NumPy default_rng; the fixtures and tests of round 1 use it)
          dev = alga_b200.synth_dev  (counter-based, identical on CPU and GPU; what bench.py runs)
threads   of the reference.  Error-free data gives the same graph for every thread count, so those
          configs may use all cores; data with substitution errors (cfg3) must use 1 (the canonical order).

Runs the reference's own GraphCreatorPrefSuf + Graph::retainOnlySmallestOffset (main.cpp:282-291) through
oracle/_ref/alga_ref_harness on the packed read set and writes tests/golden/full_<workload>_<generator>.json:
    nodes, edges, digest (alga_b200.edge_hash, two 64-bit sums), sha256 of the (E, 3) int32 array sorted by
    (source, target, offset), input_sha (sha256 of the packed words: guards generator drift), threads, seconds.
bench.py and tests/test_full_golden_gpu.py compare the graph the CUDA path builds with these.
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from alga_b200 import edge_hash, synth, synth_dev  # noqa: E402
from oracle import harness  # noqa: E402


def build(workload: str, gen: str):
    name, _, sc = workload.partition("@")
    scale = float(sc) if sc else 1.0
    if gen == "np":
        w = synth.make_config(name, scale)
        return w.reads, w.params, w.records
    w = synth_dev.make_config(name, scale)
    return w.to_readset(), w.params, w.records


def golden_path(workload: str, gen: str) -> str:
    return os.path.join(HERE, f"full_{workload.replace('@', '_at_')}_{gen}.json")


def main():
    workload, gen = sys.argv[1], sys.argv[2]
    threads = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    assert harness.available(), "build the reference harness first: make -C oracle ref"
    t0 = time.time()
    reads, params, records = build(workload, gen)
    gen_s = time.time() - t0
    t0 = time.time()
    edges, info = harness.run_prefsuf(reads, params.min_overlap, params.rs_min_overlap, params.min_offset, threads=threads)
    d = edge_hash.digest_edges(edges)
    out = {"workload": workload, "generator": gen, "nodes": reads.n, "records": records, "edges": int(edges.shape[0]),
           "digest": [f"{d[0]:016x}", f"{d[1]:016x}"], "sha256": hashlib.sha256(np.ascontiguousarray(edges).tobytes()).hexdigest(),
           "input_sha": hashlib.sha256(np.ascontiguousarray(reads.words).tobytes()).hexdigest(),
           "params": [params.min_overlap, params.rs_min_overlap, params.min_offset], "threads": threads,
           "reference_graph_s": info.get("graph_s"), "seconds_total": time.time() - t0, "gen_s": gen_s,
           "how": "oracle/_ref/alga_ref_harness prefsuf (the reference's GraphCreatorPrefSuf + retainOnlySmallestOffset)"}
    with open(golden_path(workload, gen), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()

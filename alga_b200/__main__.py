"""``python -m alga_b200 build-graph --file1 reads_1.fasta [--file2 reads_2.fasta] [--out-dir DIR]``

Builds the overlap graph of the input files on the GPU (``alga_gpu_files_to_graph``: reader, duplicate / prefix-read removal,
renumbering, GraphCreatorPrefSuf) and writes it where the UNMODIFIED ALGA binary looks for it when run with ``--serialize=1``
from DIR (main.cpp:242): ``<TEST_NAME>_beforeSimplifier.graph``.  ALGA then skips its own graph construction and carries on
with the supplement, the simplifier and the contigs.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="python -m alga_b200")
    sub = ap.add_subparsers(dest="cmd", required=True)
    b = sub.add_parser("build-graph", help="input files -> <TEST_NAME>_beforeSimplifier.graph for ALGA --serialize=1")
    b.add_argument("--file1", required=True)
    b.add_argument("--file2")
    b.add_argument("--out-dir", default=".")
    b.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)

    from .graph_file import graph_file_name, write_graph
    from .input_reader import PinnedText, build_overlap_graph, file_type_of

    t0 = time.perf_counter()
    t1 = PinnedText(open(args.file1, "rb").read())
    t2 = PinnedText(open(args.file2, "rb").read()) if args.file2 else None
    og = build_overlap_graph(t1, t2, file_type_of(args.file1), device=args.device)
    out = os.path.join(args.out_dir, graph_file_name(args.file1))
    stale = [f for f in os.listdir(args.out_dir) if f.endswith("_afterSimplifier.graph")]
    write_graph(out, og.graph)
    print(json.dumps({"graph": out, "nodes": og.reads.n, "edges": og.graph.n_edges, "params": og.params,
                      "gpu_call_ms": og.timing["total_ms"], "wall_s": time.perf_counter() - t0,
                      "warning": (f"delete {stale} first: ALGA --serialize=1 would load it instead of simplifying (main.cpp:385-390)"
                                  if stale else None)}))
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/bin/bash
# short bench only (no tests):  gpu_bench.sh <tag> [extra bench args]
tag=$1; shift
mkdir -p gpurun_out
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu "$@" > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
tail -3 gpurun_out/bench_${tag}.err
python - <<P
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}, d["roofline"].get("diag"), "e2e", round(d["e2e"]["ms_per_step"],2), "frac", round(d["roofline"]["frac"],3))
P

#!/bin/bash
# parity tests, then the short bench under a few values of a tuning env var:  gpu_sweep.sh <tag> <VAR> <v1> <v2> ...
tag=$1; var=$2; shift 2
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -3 gpurun_out/pytest_${tag}.log
for v in "$@"; do
  env $var=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$v.json 2> gpurun_out/bench_${tag}_$v.err
  python - <<P
import json
d=json.loads(open("gpurun_out/bench_${tag}_$v.json").read().strip().splitlines()[-1])
print("$var=$v", round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}, "e2e", round(d["e2e"]["ms_per_step"],2))
P
done

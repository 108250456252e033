#!/bin/bash
# A/B variants on one workload at a genome scale: gpu_ab3.sh <tag> <workload> <scale> <variant>...
tag=$1; wl=$2; sc=$3; shift 3
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != base ] && lib="$PWD/alga_b200/libalga_gpu_$v.so"
  ALGA_GPU_LIB=$lib timeout 600 python bench.py --workload $wl --scale $sc --steps 4 --warmup 2 --no-cpu > gpurun_out/bench_${tag}_$v.json 2> gpurun_out/bench_${tag}_$v.err
  python - "$v" "gpurun_out/bench_${tag}_$v.json" <<'P'
import json, sys
v, path = sys.argv[1:3]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(v, d["nodes"], round(d["ms_per_step"], 3), {k: round(x, 3) for k, x in d["roofline"]["stage_ms"].items()}, d["roofline"].get("diag"))
except Exception as e:
    print(v, "failed", e)
P
done

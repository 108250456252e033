#!/bin/bash
# bench A/B variants on several workloads: gpu_ab2.sh <tag> "<workloads>" <variant>...
tag=$1; wls=$2; shift 2
mkdir -p gpurun_out
for wl in $wls; do for v in "$@"; do
  lib=""; [ "$v" != base ] && lib="$PWD/alga_b200/libalga_gpu_$v.so"
  ALGA_GPU_LIB=$lib timeout 300 python bench.py --workload $wl --steps 8 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_${wl}_$v.json 2> gpurun_out/bench_${tag}_${wl}_$v.err
  python - "$wl" "$v" "gpurun_out/bench_${tag}_${wl}_$v.json" <<'P'
import json, sys
wl, v, path = sys.argv[1:4]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print(wl, v, round(d["ms_per_step"], 3), {k: round(x, 3) for k, x in d["roofline"]["stage_ms"].items()}, d["roofline"].get("diag"))
except Exception as e:
    print(wl, v, "failed", e)
P
done; done

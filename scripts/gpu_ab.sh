#!/bin/bash
# bench every A/B variant:  gpu_ab.sh <tag> <name> <name> ...   ("base" = the default library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != base ] && lib="$PWD/alga_b200/libalga_gpu_$v.so"
  ALGA_GPU_LIB=$lib timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_$v.json 2> gpurun_out/bench_${tag}_$v.err
  python - <<P
import json
try:
    d=json.loads(open("gpurun_out/bench_${tag}_$v.json").read().strip().splitlines()[-1])
    print("$v", round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}, d["roofline"].get("diag"))
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_${tag}_$v.err").read()[-500:])
P
done

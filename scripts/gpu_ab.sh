#!/bin/bash
# A/B of library variants (scripts/build_variant.sh) on config 4:  gpu_ab.sh <variant> ...   ("" = the default build)
mkdir -p gpurun_out
for v in "$@"; do
  lib=""; [ "$v" != "default" ] && lib="$PWD/alga_b200/libalga_gpu_$v.so"
  ALGA_GPU_LIB=$lib timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$v.json").read().strip().splitlines()[-1])
    print("$v", "ms", round(d["ms_per_step"], 3), {k: round(x, 3) for k, x in d["roofline"]["stage_ms"].items()}, "parity", d["parity"], d["roofline"]["diag"])
except Exception as e:
    print("$v: no bench line", e)
PY
  tail -1 gpurun_out/ab_$v.err
done

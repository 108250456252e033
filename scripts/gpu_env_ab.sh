#!/bin/bash
# A/B of run-time switches on config 4:  gpu_env_ab.sh name:VAR=val,VAR=val ...   (name "default" = no switches)
mkdir -p gpurun_out
for spec in "$@"; do
  name=${spec%%:*}; envs=""; [ "$spec" != "$name" ] && envs=$(echo "${spec#*:}" | tr ',' ' ')
  env $envs timeout 600 python bench.py --steps 3 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/ab_$name.json 2> gpurun_out/ab_$name.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/ab_$name.json").read().strip().splitlines()[-1])
    print("$name", "ms", round(d["ms_per_step"], 3), {k: round(x, 3) for k, x in d["roofline"]["stage_ms"].items()}, "parity", d["parity"], d["roofline"]["diag"])
except Exception as e:
    print("$name: no bench line", e)
PY
  tail -1 gpurun_out/ab_$name.err
done

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_prefsuf_gpu.py -x -q 2>&1 | tail -4
for wl in cfg2 cfg4; do
  timeout 600 python bench.py --workload $wl --steps 3 --warmup 2 --no-cpu --e2e-steps 0 > gpurun_out/dbg_$wl.json 2> gpurun_out/dbg_$wl.err; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/dbg_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"], 3),
          "parity", d["parity"], "edges", d["edges"], d["roofline"]["diag"])
except Exception as e:
    print("$wl: no bench line", e)
PY
  tail -2 gpurun_out/dbg_$wl.err
done

#!/bin/bash
# one development step on the GPU box: parity first (fast kernels, generic kernels, full-size reference goldens), then the
# bench lines of config 4 and config 2.    usage: gpurun --timeout 1500 -- 'bash scripts/gpu_step.sh <tag> [pytest files...]'
tag=${1:-step}; shift
tests=${@:-tests/test_prefsuf_gpu.py tests/test_stages_gpu.py tests/test_full_golden_gpu.py}
mkdir -p gpurun_out
timeout 900 python -m pytest $tests -x -q 2>&1 | tail -6
for wl in cfg4 cfg2; do
  timeout 600 python bench.py --workload $wl --steps 5 --warmup 3 --no-cpu > gpurun_out/${tag}_$wl.json 2> gpurun_out/${tag}_$wl.err; echo "bench $wl rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_$wl.json").read().strip().splitlines()[-1])
    print("$wl", "ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, "frac", round(d["roofline"]["frac"], 3),
          "parity", d["parity"], "edges", d["edges"], "e2e_ms", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["diag"])
except Exception as e:
    print("$wl: no bench line", e)
PY
  tail -2 gpurun_out/${tag}_$wl.err
done

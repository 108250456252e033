#!/bin/bash
mkdir -p gpurun_out
ALGA_PS_MINIMIZER=20 timeout 60 python -m pytest tests/test_prefsuf_gpu.py -x -q 2>&1 | tail -2
ALGA_PS_MINIMIZER=20 timeout 90 python bench.py --workload cfg4 --scale 0.25 --steps 3 --warmup 2 --no-cpu > gpurun_out/bench_mini.json 2> gpurun_out/bench_mini.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_mini.json").read().strip().splitlines()[-1])
    print("MINI cfg4q ms_per_step", d["ms_per_step"], d["roofline"]["stage_ms"], "edges", d["edges"], "nodes", d["nodes"])
except Exception as e:
    print("no bench line", e)
PY
tail -2 gpurun_out/bench_mini.err
timeout 40 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('default cfg2 ms_per_step', d['ms_per_step'])"

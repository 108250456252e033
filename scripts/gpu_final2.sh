#!/bin/bash
tag=${1:-r03b}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?"
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -6 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --steps 10 --warmup 3 --with-input --with-preprocess > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"])
print(d["input"]["files_to_graph"]["total_ms"], d["input"]["files_to_graph"]["stage_ms"], d["preprocess"]["device_ms"])
PY

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_supplement_gpu.py tests/test_dropin_gpu.py -q -x > gpurun_out/pytest_misc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_misc.log
tail -5 gpurun_out/pytest_misc.log
timeout 600 python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_cfg3.json 2> gpurun_out/bench_cfg3.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_cfg3.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("supplement"))); print("ms_per_step", d["ms_per_step"])
PY

#!/bin/bash
# GPU check of the input stage: parity tests, a sanitizer pass over the small ones, timing through bench.py --with-input
tag=${1:-in}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_input_gpu.py -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -25 gpurun_out/pytest_${tag}.log
if [ "${SAN:-1}" = "1" ]; then
  timeout 400 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_input_gpu.py -q -k "block_boundaries or remap_edge or errors or in_no_eol or in_long" \
      > gpurun_out/san_${tag}.log 2>&1; echo "sanitizer rc=$?" >> gpurun_out/san_${tag}.log
  tail -6 gpurun_out/san_${tag}.log
fi
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --with-input > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print(json.dumps(d.get("input"), indent=1))
print("ms_per_step", d["ms_per_step"], "frac", d["roofline"]["frac"])
PY
tail -5 gpurun_out/bench_${tag}.err

#!/bin/bash
# quick GPU iteration: parity tests + short bench
tag=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -15 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
cat gpurun_out/bench_${tag}.json; tail -5 gpurun_out/bench_${tag}.err

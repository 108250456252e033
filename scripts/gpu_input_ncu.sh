#!/bin/bash
tag=${1:-inp}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_input_gpu.py -q -x > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -5 gpurun_out/pytest_${tag}.log
timeout 300 python scripts/probes/input_profile.py 1.0 3 > gpurun_out/input_plain_${tag}.log 2>&1 && tail -1 gpurun_out/input_plain_${tag}.log &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_input_${tag}.csv \
    python scripts/probes/input_profile.py 1.0 2 > gpurun_out/ncu_input_${tag}.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_preprocess_gpu.py tests/test_input_gpu.py -q -x > gpurun_out/pytest_misc.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_misc.log
tail -3 gpurun_out/pytest_misc.log
timeout 300 python scripts/probes/input_profile.py 1.0 4 2>&1 | tail -2

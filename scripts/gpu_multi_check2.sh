#!/bin/bash
# 2-GPU sanity after changes elsewhere: real-peer-memory parity test + the driver's N = 2 bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_multi_gpu.py -q -m gpu 2>&1 | tail -3
bash scripts/gpu_multi_bench.sh 2 r03b_n2

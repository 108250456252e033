#!/bin/bash
# multi-GPU bench: usage  gpurun --gpus N -- 'bash scripts/gpu_multi.sh N tag'
n=${1:-2}; tag=${2:-m}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/gpus_${tag}.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $n --steps 5 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "rc=$?"; cat gpurun_out/bench_${tag}.json; tail -20 gpurun_out/bench_${tag}.err

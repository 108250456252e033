#!/bin/bash
# N-GPU: sharded parity on real peer memory, then the N-GPU bench line.  gpurun --gpus N -- 'bash scripts/gpu_multi_check.sh N tag'
n=${1:-2}; tag=${2:-m}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 \
   tests/sharded_worker.py 0.02 > gpurun_out/sharded_${tag}.log 2>&1; echo "sharded rc=$?"
grep -E "sharded world|Error|error" gpurun_out/sharded_${tag}.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; cut -c1-1800 gpurun_out/bench_${tag}.json; grep -v "^W1\|^\*\*\*\|OMP_NUM" gpurun_out/bench_${tag}.err | tail -15

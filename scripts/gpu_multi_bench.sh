#!/bin/bash
# N-GPU bench only:  gpurun --gpus N -- 'bash scripts/gpu_multi_bench.sh N tag'
n=${1:-2}; tag=${2:-m}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $n --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
echo "bench rc=$?"; grep -v "^W1\|^\*\*\*\|OMP_NUM\|^\[W" gpurun_out/bench_${tag}.err | tail -8
python - <<P
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("N=$n", round(d["ms_per_step"],3), {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}, d["roofline"].get("diag"), "e2e", round(d["e2e"]["ms_per_step"],2))
P

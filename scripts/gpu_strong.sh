#!/bin/bash
# strong scaling of one workload:  gpurun --gpus N -- 'bash scripts/gpu_strong.sh N tag workload scale'
n=${1:-2}; tag=${2:-s}; wl=${3:-cfg4}; sc=${4:-0.25}
mkdir -p gpurun_out
if [ "$n" = 1 ]; then
  timeout 1200 python bench.py --workload $wl --scale $sc --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
else
  timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
     bench.py --gpus $n --workload $wl --scale $sc --scaling strong --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err
fi
echo "bench rc=$?"; grep -v "^W1\|^\*\*\*\|OMP_NUM\|^\[W" gpurun_out/bench_${tag}.err | tail -8
python - <<P
import json
d=json.loads(open("gpurun_out/bench_${tag}.json").read().strip().splitlines()[-1])
print("N=$n", d["nodes"], "nodes", round(d["ms_per_step"],3), "ms", round(d["value"]/1e6,1), "M reads/s", {k: round(x,3) for k,x in d["roofline"]["stage_ms"].items()}, d["roofline"].get("diag"), "frac", round(d["roofline"]["frac"],3), "gen_s", round(d["gen_s"],1))
P

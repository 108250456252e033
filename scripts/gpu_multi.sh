#!/bin/bash
# multi-GPU step: parity on real peer memory, then the strong-scaling bench line(s).  usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [steps]'
n=${1:-2}; steps=${2:-5}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ $n -le 2 ]; then timeout 600 python -m pytest tests/test_sharded_multi_gpu.py -x -q 2>&1 | tail -5; fi
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $n --steps $steps --warmup 3 > gpurun_out/multi_n$n.json 2> gpurun_out/multi_n$n.err; echo "bench N=$n rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/multi_n$n.json").read().strip().splitlines()[-1])
    print("N=$n", "ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, "parity", d["parity"], "edges", d["edges"],
          "e2e_ms", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["diag"])
except Exception as e:
    print("no bench line", e)
PY
tail -5 gpurun_out/multi_n$n.err

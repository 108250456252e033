#!/bin/bash
# multi-GPU step: parity on real peer memory, then the strong-scaling bench line of config 4, once per setting of the seed-record
# exchange (ALGA_SHARD_SEED_KEYS).    usage: gpurun --gpus N -- 'bash scripts/gpu_multi.sh N [steps] ["0 1"]'
n=${1:-2}; steps=${2:-5}; variants=${3:-"0 1"}
mkdir -p gpurun_out
nvidia-smi -L | head -8
if [ $n -le 2 ]; then timeout 900 python -m pytest tests/test_sharded_multi_gpu.py tests/test_multi_abi_gpu.py -x -q 2>&1 | tail -5; fi
for k in $variants; do
ALGA_SHARD_SEED_KEYS=$k timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $n --steps $steps --warmup 3 > gpurun_out/multi_n${n}_k$k.json 2> gpurun_out/multi_n${n}_k$k.err; echo "bench N=$n keys=$k rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/multi_n${n}_k$k.json").read().strip().splitlines()[-1])
    print("N=$n keys=$k", "ms", round(d["ms_per_step"], 3), {k: round(v, 3) for k, v in d["roofline"]["stage_ms"].items()}, "parity", d["parity"], "edges", d["edges"],
          "e2e_ms", round(d["e2e"]["ms_per_step"], 2), d["roofline"]["diag"])
except Exception as e:
    print("no bench line", e)
PY
tail -3 gpurun_out/multi_n${n}_k$k.err
done

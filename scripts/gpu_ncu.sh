#!/bin/bash
# ncu only: launch list + full capture of kernels matching $2 (regex), using the short bench command
tag=${1:-n}; pat=${2:-phase1_tpr|phase2_tpr}
mkdir -p gpurun_out
timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${tag}.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${tag}.log 2>&1
echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$pat" -s 2 -c 2 \
    -f -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
echo "ncu full rc=$?"
cat gpurun_out/plain_${tag}.log | cut -c1-300

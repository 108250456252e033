#!/bin/bash
# One GPU-box round: parity tests, bench, ncu launch list and full captures of the two dominant kernels.
# usage (from the dev container):  gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh <tag>'
tag=${1:-r01}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_${tag}.log
true
tail -3 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
cat gpurun_out/bench_${tag}.json
if [ "${NCU:-1}" = "1" ]; then
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${tag}.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${tag}.log 2>&1
  echo "ncu launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'phase1_tpr|phase2_tpr' -s 2 -c 2 \
      -f -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
  echo "ncu full rc=$?"
fi

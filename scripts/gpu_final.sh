#!/bin/bash
# Round-end evidence in one GPU-box call: smoke, all GPU tests, the bench line (with the input / preprocess legs and the CPU arm),
# the reference arm, ncu launch lists (graph build; files-to-graph) and full captures of the dominant kernels.
# usage (from the dev container):  gpurun --timeout 1500 -- 'bash scripts/gpu_final.sh <tag>'
tag=${1:-r03}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?" | tee -a gpurun_out/smoke_${tag}.log
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_${tag}.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_${tag}.log
tail -4 gpurun_out/pytest_${tag}.log
timeout 600 python bench.py --steps 10 --warmup 3 --with-input --with-preprocess > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "bench ref rc=$?"
cat gpurun_out/bench_ref_${tag}.json | cut -c1-400
if [ "${NCU:-1}" = "1" ]; then
  timeout 300 python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/plain_${tag}.log 2>&1 &&
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_${tag}.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_launches_${tag}.log 2>&1
  echo "ncu launches rc=$?"
  timeout 300 python scripts/probes/input_profile.py 1.0 2 > gpurun_out/input_plain_${tag}.log 2>&1 &&
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_input_${tag}.csv \
      python scripts/probes/input_profile.py 1.0 2 > gpurun_out/ncu_input_${tag}.log 2>&1
  echo "ncu input launches rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'phase1_tpr|phase2_tpr' -s 2 -c 2 \
      -f -o gpurun_out/prof_${tag} python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/ncu_full_${tag}.log 2>&1
  echo "ncu full rc=$?"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'scan_records|pack_records|probe_prefixes' -s 3 -c 3 \
      -f -o gpurun_out/prof_input_${tag} python scripts/probes/input_profile.py 1.0 2 > gpurun_out/ncu_full_input_${tag}.log 2>&1
  echo "ncu full input rc=$?"
fi

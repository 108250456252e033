#!/bin/bash
# Round-end evidence on one GPU (config 4, the bench default): smoke, bench line incl. CPU baseline, reference arm, ncu launch
# list with DRAM bytes of one step (after the plain run exited 0), full captures of the two fast kernels.
#   gpurun --timeout 2400 -- 'bash scripts/gpu_evidence.sh <tag>'
tag=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/gpu_${tag}.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_${tag}.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke_${tag}.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_${tag}.json 2> gpurun_out/bench_ref_${tag}.err; echo "reference arm rc=$?"
timeout 600 python bench.py --workload cfg2 --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_${tag}_cfg2.json 2> gpurun_out/bench_${tag}_cfg2.err; echo "bench cfg2 rc=$?"
bash scripts/gpu_launches.sh ${tag}
bash scripts/gpu_ncu4.sh ${tag} both

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
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/plain_${tag}.log 2>&1 &&
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    -k regex:'read_stats|repack_reads|build_index|phase[12]_|rebuild_rows|over_to_csr|rows_to_csr|scan_|end_cursor|scatter_|split_pairs|sort_rows|sort_big|count_sources|peek' \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_launches_${tag}.log 2>&1
echo "ncu launches rc=$?"
bash scripts/gpu_ncu4.sh ${tag} both

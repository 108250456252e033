#!/bin/bash
# round-2 experiment 2: DRAM access granularity probe, new bench (device-generated cfg4, parity gate), ncu at cfg4 size
mkdir -p gpurun_out
nproc; free -g | head -2
./scripts/probes/random_granule > gpurun_out/random_granule.txt 2>&1; cat gpurun_out/random_granule.txt
for wl in "cfg1 dev" "cfg2 dev" "cfg2 np"; do set -- $wl
  timeout 600 python bench.py --workload $1 --gen $2 --steps 5 --warmup 3 --no-cpu > gpurun_out/exp2_$1_$2.json 2> gpurun_out/exp2_$1_$2.err; echo "bench $1 $2 rc=$?"
  python -c "
import json; d=json.loads(open('gpurun_out/exp2_$1_$2.json').read().strip().splitlines()[-1])
print('$1 $2', 'ms', round(d['ms_per_step'],3), 'parity', d['parity'], 'nodes', d['nodes'], 'edges', d['edges'], 'gen_s', round(d['gen_s'],1), 'e2e_ms', round(d['e2e']['ms_per_step'],2))"
  tail -2 gpurun_out/exp2_$1_$2.err
done
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/exp2_cfg4.json 2> gpurun_out/exp2_cfg4.err; echo "bench cfg4 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/exp2_cfg4.json').read().strip().splitlines()[-1])
print('cfg4', 'ms', round(d['ms_per_step'],3), d['roofline']['stage_ms'], 'frac', round(d['roofline']['frac'],3), 'parity', d['parity'], d['edge_digest'], 'nodes', d['nodes'], 'edges', d['edges'], 'gen_s', round(d['gen_s'],1), 'e2e', d['e2e'], 'cpu', d['cpu_baseline'])"
tail -3 gpurun_out/exp2_cfg4.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'phase1_tpr|phase2_tpr' -s 4 -c 4 \
   -f -o gpurun_out/prof_r2a_cfg4 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_r2a.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4

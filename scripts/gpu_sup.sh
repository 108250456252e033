#!/bin/bash
# Error-rate supplement: its tests, then config 3 end to end with the per-pass trace on stderr.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_supplement_gpu.py tests/test_dropin_gpu.py -x -q 2>&1 | tail -4
ALGA_SUP_TRACE=1 timeout 900 python bench.py --workload cfg3 --gen np --steps 3 --warmup 2 --no-cpu > gpurun_out/sup_cfg3.json 2> gpurun_out/sup_cfg3.err; echo "bench cfg3 rc=$?"
python -c "
import json; d=json.loads(open('gpurun_out/sup_cfg3.json').read().strip().splitlines()[-1]); print('cfg3 ms', d['ms_per_step'], 'parity', d['parity']); print(d.get('supplement'))"
grep "supplement pass" gpurun_out/sup_cfg3.err | tail -12

#!/bin/bash
# round-2 experiment 1: sliding-minimizer buckets -- parity, then config 4 at 25 Mbp and at full size
mkdir -p gpurun_out
export ALGA_SYNTH_CACHE=/tmp/alga_synth
echo "== parity with minimizer buckets"
ALGA_PS_MINIMIZER=20 ALGA_PS_MINIMIZER_SLIDE=1 ALGA_PS_BUCKET_LOAD=1 timeout 300 python -m pytest tests/test_prefsuf_gpu.py -x -q 2>&1 | tail -3
run() {  # tag scale envs...
  tag=$1; scale=$2; shift 2
  env "$@" timeout 600 python bench.py --workload cfg4 --scale $scale --steps 3 --warmup 2 --no-cpu > gpurun_out/exp1_$tag.json 2> gpurun_out/exp1_$tag.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/exp1_$tag.json").read().strip().splitlines()[-1])
    print("$tag", "ms", round(d["ms_per_step"],3), {k: round(v,3) for k,v in d["roofline"]["stage_ms"].items()}, "edges", d["edges"], "nodes", d["nodes"], "gen_s", round(d["gen_s"],1), d["roofline"]["diag"])
except Exception as e:
    print("$tag no bench line", e)
PY
  tail -2 gpurun_out/exp1_$tag.err
}
run q_default 0.25 A=1
run q_m20_l2 0.25 ALGA_PS_MINIMIZER=20 ALGA_PS_MINIMIZER_SLIDE=1
run q_m20_l1 0.25 ALGA_PS_MINIMIZER=20 ALGA_PS_MINIMIZER_SLIDE=1 ALGA_PS_BUCKET_LOAD=1
run q_m24_l1 0.25 ALGA_PS_MINIMIZER=24 ALGA_PS_MINIMIZER_SLIDE=1 ALGA_PS_BUCKET_LOAD=1
run q_m16_l1 0.25 ALGA_PS_MINIMIZER=16 ALGA_PS_MINIMIZER_SLIDE=1 ALGA_PS_BUCKET_LOAD=1
run f_default 1.0 A=1
run f_m20_l1 1.0 ALGA_PS_MINIMIZER=20 ALGA_PS_MINIMIZER_SLIDE=1 ALGA_PS_BUCKET_LOAD=1
run f_m24_l1 1.0 ALGA_PS_MINIMIZER=24 ALGA_PS_MINIMIZER_SLIDE=1 ALGA_PS_BUCKET_LOAD=1

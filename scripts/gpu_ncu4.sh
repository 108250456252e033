#!/bin/bash
# ncu --set full of the fast kernels at config-4 size (second build of a 2-build run), after the plain run exited 0
# usage: gpu_ncu4.sh <tag> [p1|p2|both]
mkdir -p gpurun_out
tag=${1:-r2d}; which=${2:-both}
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/plain_$tag.log 2>&1 || exit 1
if [ $which != p1 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase2_tpr -s 1 -c 1 -f -o gpurun_out/prof_${tag}_p2 \
   python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_${tag}_p2.log 2>&1; echo "ncu p2 rc=$?"
fi
if [ $which != p2 ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:phase1_tpr -s 2 -c 1 -f -o gpurun_out/prof_${tag}_p1 \
   python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_${tag}_p1.log 2>&1; echo "ncu p1 rc=$?"
fi

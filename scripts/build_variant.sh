#!/bin/bash
# A/B build of libalga_gpu with extra -D flags:  build_variant.sh <name> "<flags>"  -> alga_b200/libalga_gpu_<name>.so
name=$1; flags=$2
cd "$(dirname "$0")/../alga_b200/csrc" || exit 1
mkdir -p build_$name
for f in api prefsuf_kernels tpr_kernels misc_kernels supplement preprocess input simplify multi sorted_stages; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -Xptxas -v $flags -c $f.cu -o build_$name/$f.o 2> build_$name/$f.log &
done
wait
grep -hE "spill" build_$name/tpr_kernels.log | sort | uniq -c
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../libalga_gpu_$name.so build_$name/*.o -cudart static && echo built $name

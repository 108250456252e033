#!/bin/bash
# ncu launch list (time + DRAM bytes per launch) of the library's kernels in one bench.py run with two builds of config 4
tag=${1:-r2}
mkdir -p gpurun_out
timeout 600 python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/plain_${tag}.log 2>&1 || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    -k regex:'read_stats|repack_reads|build_index|phase[12]_|rebuild_rows|over_to_csr|rows_to_csr|scan_|end_cursor|scatter_|split_pairs|sort_rows|sort_big|count_sources|peek|seed_records|fill_buckets|csr_records|csr_rows|chain_overflow|DeviceRadixSort' \
    --log-file gpurun_out/launches_${tag}.csv python bench.py --steps 1 --warmup 1 --no-cpu --e2e-steps 0 > gpurun_out/ncu_launches_${tag}.log 2>&1
echo "ncu launches rc=$?"

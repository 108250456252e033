#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_input_gpu.py -q -x 2>&1 | tail -2
echo "--- DMA download + kernel peeks (default)"
timeout 300 python scripts/probes/input_profile.py 1.0 4 2>&1 | tail -2
echo "--- copy kernel download, 4 blocks"
ALGA_FE_COPY_KERNEL=1 ALGA_FE_COPY_BLOCKS=4 timeout 300 python scripts/probes/input_profile.py 1.0 4 2>&1 | tail -1

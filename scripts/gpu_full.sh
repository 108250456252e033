#!/bin/bash
# the whole GPU suite + smoke, as the driver runs them at round end
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8

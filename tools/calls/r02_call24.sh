#!/bin/bash
# round 2, last 2-GPU call: the 20-iteration half of SURVEY 8e's gate, and smoke() on the final build
set -u
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_multigpu.py -m gpu -v -k "after_20_iterations" > gpurun_out/r02z_pytest_2gpu_20it.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02z_pytest_2gpu_20it.log; grep -E "PASSED|FAILED|ERROR|passed|failed|assert|Error" gpurun_out/r02z_pytest_2gpu_20it.log | cut -c1-300 | tail -12
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3

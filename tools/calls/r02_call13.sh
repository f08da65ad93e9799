#!/bin/bash
# round 2, third 2-GPU call: multi-GPU tests on the final library
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_zz_elbo_gpu.py tests/test_cli.py -m gpu -v > gpurun_out/r02o_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02o_pytest_2gpu.log; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r02o_pytest_2gpu.log | cut -c1-200 | tail -12

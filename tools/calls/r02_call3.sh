#!/bin/bash
# round 2, third 1-GPU call: full suite on the rebuilt engine (pipelined sweep, group ctx code present), sweep variants timed
set -u
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r02d_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02d_pytest_gpu.log; tail -40 gpurun_out/r02d_pytest_gpu.log | cut -c1-250
echo "--- default library (pipe1, 3 blocks)"; timeout 300 python tools/sweep_experiments.py netflix '{"HPF_SWEEP_G": ["8", "16"]}' 2>&1 | tee gpurun_out/r02d_exp_default.log
for v in pipe0 pipe1_mb2 pipe1_mb4; do echo "--- $v"; HPF_LIB=$PWD/build/variants/lib_$v.so timeout 300 python tools/sweep_experiments.py netflix '{"HPF_SWEEP_G": ["8", "16"]}' 2>&1 | tee gpurun_out/r02d_exp_$v.log; done
echo "--- msd"; timeout 300 python tools/sweep_experiments.py msd '{"HPF_SWEEP_G": ["16"]}' 2>&1 | tee gpurun_out/r02d_exp_msd_default.log
HPF_LIB=$PWD/build/variants/lib_pipe0.so timeout 300 python tools/sweep_experiments.py msd '{"HPF_SWEEP_G": ["16"]}' 2>&1 | tee gpurun_out/r02d_exp_msd_pipe0.log

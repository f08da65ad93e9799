#!/bin/bash
# First GPU call of the next round (one GPU): what round 1 built after its GPU budget ran out, measured.
#   gpurun --timeout 1500 -- 'bash tools/calls/r02_first_call.sh'
# Two-GPU follow-up (overlap of the all-reduce, 2-GPU tests):
#   gpurun --gpus 2 --timeout 900 -- 'python -m pytest tests/test_multigpu.py tests/test_zz_elbo_gpu.py -m gpu -q > gpurun_out/pytest_2gpu.log 2>&1;
#     for o in 0 1; do HPF_AR_OVERLAP=$o python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
#       bench.py --gpus 2 --no-cpu-baseline > gpurun_out/bench_n2_overlap$o.json 2> gpurun_out/bench_n2_overlap$o.err; done'
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
HPF_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_zz_elbo_gpu.py -m gpu -q -k variants > gpurun_out/pytest_variants.log 2>&1; tail -3 gpurun_out/pytest_variants.log
timeout 600 python tools/sweep_experiments.py netflix '{"HPF_HEAD_VARIANT": ["0", "1", "2", "3", "4", "5", "7"]}' > gpurun_out/exp_head_variants.log 2>&1; cat gpurun_out/exp_head_variants.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; cat gpurun_out/bench.json
timeout 600 python tools/bench_extras.py elbo c4 > gpurun_out/extras.json 2> gpurun_out/extras.err; echo "extras exit $?"; cat gpurun_out/extras.json

#!/bin/bash
# round 2, first 1-GPU call: the suite on the library as round 1 left it, the head-kernel variants that were
# never run, the set-up phase trace, the topn bench and one bench line.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02b_pytest_gpu.log; tail -5 gpurun_out/r02b_pytest_gpu.log
HPF_TEST_EXPERIMENTS=1 timeout 300 python -m pytest tests/test_zz_elbo_gpu.py -m gpu -q -k variants > gpurun_out/r02b_pytest_variants.log 2>&1; tail -3 gpurun_out/r02b_pytest_variants.log
timeout 500 python tools/sweep_experiments.py netflix '{"HPF_HEAD_VARIANT": ["0", "1", "2", "3", "4", "5", "7"]}' > gpurun_out/r02b_exp_head_variants.log 2>&1; cat gpurun_out/r02b_exp_head_variants.log
HPF_TRACE=1 timeout 300 python bench.py --no-cpu-baseline --steps 5 > gpurun_out/r02b_bench_trace.json 2> gpurun_out/r02b_bench_trace.err; grep "hpf trace" gpurun_out/r02b_bench_trace.err | tail -24
timeout 300 python tools/bench_topn.py > gpurun_out/r02b_topn.json 2> gpurun_out/r02b_topn.err; cat gpurun_out/r02b_topn.json
timeout 600 python tools/bench_extras.py elbo c4 > gpurun_out/r02b_extras.json 2> gpurun_out/r02b_extras.err; echo "extras exit $?"; cat gpurun_out/r02b_extras.json

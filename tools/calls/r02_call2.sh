#!/bin/bash
# round 2, second 1-GPU call: the rebuilt engine (3-array dense update, device work lists, single-sort CSC, head variant 7):
# full GPU suite, the L2 gather micro-benchmark, the bench line with its secondary legs, the set-up trace.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02c_pytest_gpu.log; tail -25 gpurun_out/r02c_pytest_gpu.log
for mode in ldg8 ldg32 bulk gather4; do for rows in 17770 480189; do timeout 120 tools/gather_bench_bin $mode $rows 26 >> gpurun_out/r02c_gather_bench.jsonl 2>> gpurun_out/r02c_gather_bench.err; echo "gather $mode $rows exit $?"; done; done
cat gpurun_out/r02c_gather_bench.jsonl
HPF_TRACE=1 timeout 600 python bench.py > gpurun_out/r02c_bench.json 2> gpurun_out/r02c_bench.err; echo "bench exit $?"; grep "hpf trace" gpurun_out/r02c_bench.err | tail -6; tail -3 gpurun_out/r02c_bench.err; cat gpurun_out/r02c_bench.json

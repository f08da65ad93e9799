#!/bin/bash
# round 2, 1-GPU call 5: suite on the current library (topn v2: radix select, two CTAs/SM), C5 timings, head block share sweep, ncu captures
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02f_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02f_pytest_gpu.log; tail -30 gpurun_out/r02f_pytest_gpu.log | cut -c1-250
timeout 400 python tools/bench_topn.py > gpurun_out/r02f_topn.json 2> gpurun_out/r02f_topn.err; echo "topn exit $?"; cat gpurun_out/r02f_topn.json; tail -3 gpurun_out/r02f_topn.err
timeout 400 python tools/sweep_experiments.py netflix '{"HPF_DENSE_BLOCK_SHARE": ["0.06", "0.045", "0.035", "0.02"]}' 2>&1 | tee gpurun_out/r02f_exp_block_share.log
KREGEX='regex:sweep_kernel|update_kernel|combine_kernel|colsum|heldout|head_kernel|head_reduce|split_kernel|split_aux_kernel|derive_kernel|wl_|orient_|DeviceRadixSort|DeviceScan'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 600 --csv --log-file gpurun_out/r02f_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02f_bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sweep_kernel|update_kernel|head_kernel" --launch-skip 12 -c 8 -o gpurun_out/r02f_iter_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02f_ncu_full.log 2>&1; echo "ncu full exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:topn_kernel|rank_kernel" -c 3 -o gpurun_out/r02f_topn_full -f python tools/bench_topn.py 0.25 > gpurun_out/r02f_ncu_topn.log 2>&1; echo "ncu topn exit $?"
ls -la gpurun_out | tail -12

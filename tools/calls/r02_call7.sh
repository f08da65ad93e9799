#!/bin/bash
# round 2, 1-GPU call 7: suite (tensor-core item ranks, short-chunk fix, head cost model), C5 timings, ncu summaries as CSV
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02h_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02h_pytest_gpu.log; tail -30 gpurun_out/r02h_pytest_gpu.log | cut -c1-250
timeout 400 python tools/bench_topn.py > gpurun_out/r02h_topn.json 2> gpurun_out/r02h_topn.err; echo "topn exit $?"; cat gpurun_out/r02h_topn.json; tail -3 gpurun_out/r02h_topn.err
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02h_bench.json 2> gpurun_out/r02h_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02h_bench.err; cut -c1-400 gpurun_out/r02h_bench.json
KREGEX='regex:sweep_kernel|update_kernel|combine_kernel|colsum|heldout|head_kernel|head_reduce|split_kernel|split_aux_kernel|derive_kernel|wl_|orient_|DeviceRadixSort|DeviceScan|expand_rows|check_range|dense_y|head_flag|head_split|run_ptr'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 700 --csv --log-file gpurun_out/r02h_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02h_bench_under_ncu.log 2>&1; echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sweep_kernel|update_kernel|head_kernel" --launch-skip 14 -c 9 -o /tmp/r02h_iter_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02h_ncu_full.log 2>&1; echo "ncu full exit $?"
bash tools/ncu_export.sh /tmp/r02h_iter_full.ncu-rep gpurun_out/r02h_iter_full
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:topn_kernel|rank_mma_kernel" -c 2 -o /tmp/r02h_topn_full -f python tools/bench_topn.py 0.25 > gpurun_out/r02h_ncu_topn.log 2>&1; echo "ncu topn exit $?"
bash tools/ncu_export.sh /tmp/r02h_topn_full.ncu-rep gpurun_out/r02h_topn_full
ncu -i /tmp/r02h_topn_full.ncu-rep --page source --csv 2>/dev/null | cut -d, -f1-8 | gzip > gpurun_out/r02h_topn_source.csv.gz
ls -la gpurun_out | tail -14

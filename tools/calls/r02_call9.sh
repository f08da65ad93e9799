#!/bin/bash
# round 2, 1-GPU call 9: suite on the lazy-E[log v] build (fallback out of the hot loop), bench, C5 timings, ncu summaries as CSV
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02k_pytest_gpu.log; tail -30 gpurun_out/r02k_pytest_gpu.log | cut -c1-250
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02k_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02k_bench.json").read())
print(d["ms_per_step"], d["roofline"]["per_kernel_ms"], "e2e", d["e2e"]["ms_per_step"])
print("msd", d["hbm_bound_workload"]["ms_per_step"], d["hbm_bound_workload"]["roofline"]["per_kernel_ms"])
print("steady", d["steady_state"]["ms_per_step"], d["steady_state"]["slow_path_nnz_per_iteration"])
PY
timeout 400 python tools/bench_topn.py > gpurun_out/r02k_topn.json 2> gpurun_out/r02k_topn.err; echo "topn exit $?"; cat gpurun_out/r02k_topn.json; tail -3 gpurun_out/r02k_topn.err
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sweep_kernel|update_kernel|head_kernel" --launch-skip 18 -c 10 -o /tmp/r02k_iter_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02k_ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/ncu_export.py /tmp/r02k_iter_full.ncu-rep gpurun_out/r02k_iter_full
timeout 600 ncu --set full --clock-control none --import-source on -k "regex:topn_kernel|rank_mma_kernel" --launch-skip 2 -c 2 -o /tmp/r02k_topn_full -f python tools/bench_topn.py 0.25 > gpurun_out/r02k_ncu_topn.log 2>&1; echo "ncu topn exit $?"
python tools/ncu_export.py /tmp/r02k_topn_full.ncu-rep gpurun_out/r02k_topn_full
ls -la gpurun_out | tail -12

#!/bin/bash
# round 2, 1-GPU call 10: suite + bench + C5 timings on the build with merged head launches, __fdividef, rank kernel v2
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02l_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02l_pytest_gpu.log; tail -30 gpurun_out/r02l_pytest_gpu.log | cut -c1-250
timeout 600 python bench.py > gpurun_out/r02l_bench.json 2> gpurun_out/r02l_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02l_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02l_bench.json").read())
print(d["ms_per_step"], d["roofline"]["per_kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print("msd", d["hbm_bound_workload"]["ms_per_step"], d["hbm_bound_workload"]["roofline"]["per_kernel_ms"])
print("steady", d["steady_state"]["ms_per_step"], d["steady_state"]["slow_path_nnz_per_iteration"])
print("cpu", d["cpu_baseline"])
PY
timeout 400 python tools/bench_topn.py > gpurun_out/r02l_topn.json 2> gpurun_out/r02l_topn.err; echo "topn exit $?"; cat gpurun_out/r02l_topn.json; tail -3 gpurun_out/r02l_topn.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 3 > gpurun_out/r02l_bench_reference.json 2> gpurun_out/r02l_bench_reference.err; echo "reference arm exit $?"; cut -c1-600 gpurun_out/r02l_bench_reference.json

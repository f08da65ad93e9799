#!/bin/bash
# round 2, 1-GPU call 11: packed index|rating stream: suite, A/B timing (HPF_PACK=0/1) at Netflix and MSD scale, bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02m_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02m_pytest_gpu.log; tail -30 gpurun_out/r02m_pytest_gpu.log | cut -c1-250
timeout 400 python tools/sweep_experiments.py netflix '{"HPF_PACK": ["0", "1"]}' 2>&1 | tee gpurun_out/r02m_exp_pack_netflix.log
timeout 400 python tools/sweep_experiments.py msd '{"HPF_PACK": ["0", "1"]}' 2>&1 | tee gpurun_out/r02m_exp_pack_msd.log
timeout 600 python bench.py > gpurun_out/r02m_bench.json 2> gpurun_out/r02m_bench.err; echo "bench exit $?"; tail -2 gpurun_out/r02m_bench.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02m_bench.json").read())
print(d["ms_per_step"], d["roofline"]["per_kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print("msd", d["hbm_bound_workload"]["ms_per_step"], d["hbm_bound_workload"]["roofline"]["per_kernel_ms"])
PY

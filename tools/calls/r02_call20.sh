#!/bin/bash
# round 2, 1-GPU call: update kernel with the warp-vote digamma -- parity subset, bench without the CPU leg
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/r02v_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02v_pytest_gpu.log; tail -5 gpurun_out/r02v_pytest_gpu.log | cut -c1-250
timeout 900 python bench.py --no-cpu-baseline > gpurun_out/r02v_bench.json 2> gpurun_out/r02v_bench.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02v_bench.json") if l.startswith("{")][-1])
r=lambda k:{a:round(b,3) for a,b in k.items()}
print(d["ms_per_step"], r(d["roofline"]["per_kernel_ms"]), "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
h=d["hbm_bound_workload"]; print("msd", h["ms_per_step"], r(h["roofline"]["per_kernel_ms"]))
print("steady", d["steady_state"]["ms_per_step"])
PY

#!/bin/bash
# round 2, 1-GPU call: update kernel on the special-function unit's own approximations -- whole GPU suite, then the bench line
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02u_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02u_pytest_gpu.log; tail -30 gpurun_out/r02u_pytest_gpu.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r02u_bench.json 2> gpurun_out/r02u_bench.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02u_bench.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["roofline"]["per_kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
print({k: (v.get("ms_per_step"), v.get("roofline", {}).get("per_kernel_ms")) for k, v in d.items() if isinstance(v, dict) and "ms_per_step" in v})
PY

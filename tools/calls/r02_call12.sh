#!/bin/bash
# round 2, 1-GPU call 12: ratings uploaded under the first device sort: suite, bench (e2e), set-up trace
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02n_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02n_pytest_gpu.log; tail -30 gpurun_out/r02n_pytest_gpu.log | cut -c1-250
HPF_TRACE=1 timeout 600 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err; echo "bench exit $?"; grep "hpf trace" gpurun_out/r02n_bench.err | tail -3; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02n_bench.json").read())
print(d["ms_per_step"], d["roofline"]["per_kernel_ms"], "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"])
PY

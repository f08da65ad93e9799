#!/bin/bash
# round 2, final 1-GPU call: whole GPU suite, the bench line with all its legs, C5 timings, ncu launch list and full captures
set -u
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r02x_pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02x_pytest_gpu.log; tail -5 gpurun_out/r02x_pytest_gpu.log | cut -c1-250
timeout 900 python bench.py > gpurun_out/r02x_bench.json 2> gpurun_out/r02x_bench.err; echo "bench exit $?"; python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r02x_bench.json") if l.startswith("{")][-1])
r=lambda k:{a:round(b,3) for a,b in k.items()}
print(d["ms_per_step"], r(d["roofline"]["per_kernel_ms"]), "e2e", d["e2e"]["ms_per_step"], "frac", d["roofline"]["frac"], "cpu", d["cpu_baseline"])
h=d["hbm_bound_workload"]; print("msd", h["ms_per_step"], r(h["roofline"]["per_kernel_ms"]), h["roofline"]["frac"])
print("steady", d["steady_state"]["ms_per_step"], d["steady_state"]["slow_path_nnz_per_iteration"])
PY
timeout 400 python tools/bench_topn.py > gpurun_out/r02x_topn.json 2> gpurun_out/r02x_topn.err; echo "topn exit $?"; cat gpurun_out/r02x_topn.json | cut -c1-600
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02x_ncu_launches.log 2>&1; echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:sweep_kernel|update_kernel|head_kernel" --launch-skip 18 -c 10 -o /tmp/r02x_iter_full -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02x_ncu_full.log 2>&1; echo "ncu full exit $?"
python tools/ncu_export.py /tmp/r02x_iter_full.ncu-rep gpurun_out/r02x_iter_full
ls -la gpurun_out | tail -8

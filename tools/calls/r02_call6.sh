#!/bin/bash
# round 2, the 8-GPU call: strong scaling of the three BASELINE workloads on one 8 x B200 box (one fixed problem each,
# users sharded, items replicated, chunked all-reduce under the sweeps), plus Netflix-scale at N=4.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02g_smi.txt
run() { # nproc workload extra...
  local n=$1 w=$2; shift 2
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --no-cpu-baseline "$@" > gpurun_out/r02g_bench_n${n}_$w.json 2> gpurun_out/r02g_bench_n${n}_$w.err
  echo "bench n=$n $w exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02g_bench_n${n}_$w.err | tail -3
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r02g_bench_n${n}_$w.json").read())
    print("  value %.4g nnz/s  %.3f ms/step  e2e %.1f ms  chunks %s  per-kernel %s" % (d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["config"]["item_chunks"], {k: round(v, 3) for k, v in d["roofline"]["per_kernel_ms"].items()}))
    if "weak" in d: print("  weak: %.4g nnz/s %.3f ms/step" % (d["weak"]["value"], d["weak"]["ms_per_step"]))
except Exception as e:
    print("  no line:", e)
PY
}
run 8 netflix
run 8 msd
run 8 bpf-1b --e2e-steps 1
run 4 netflix --no-extras
run 4 msd

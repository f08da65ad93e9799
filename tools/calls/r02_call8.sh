#!/bin/bash
# round 2, second 2-GPU call: the multi-GPU tests on the current library (exact-fallback test fixed), N=2 lines
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py tests/test_zz_elbo_gpu.py tests/test_cli.py -m gpu -v > gpurun_out/r02i_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02i_pytest_2gpu.log; grep -E "FAILED|ERROR|passed|failed" gpurun_out/r02i_pytest_2gpu.log | cut -c1-200 | tail -12
for w in netflix msd; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --no-cpu-baseline --no-extras > gpurun_out/r02i_bench_n2_$w.json 2> gpurun_out/r02i_bench_n2_$w.err; echo "bench $w exit $?"; grep '^{' gpurun_out/r02i_bench_n2_$w.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done

#!/bin/bash
# 2-GPU: multi-GPU tests + MSD N=2 chunk variants on the build that issues its collectives after the compute launches
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r02q_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02q_pytest_2gpu.log; tail -4 gpurun_out/r02q_pytest_2gpu.log | cut -c1-200
for ch in 1 6; do
  HPF_AR_CHUNKS=$ch timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload msd --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02q_bench_n2_msd_chunks$ch.json 2> gpurun_out/r02q_bench_n2_msd_chunks$ch.err; echo "bench chunks=$ch exit $?"; grep '^{' gpurun_out/r02q_bench_n2_msd_chunks$ch.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['roofline']['per_kernel_ms'])"
done

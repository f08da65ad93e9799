#!/bin/bash
# round 2, first 2-GPU call: every multi-GPU test (two ctxs + NCCL, chunked all-reduce, exact-fallback rollback, one ctx
# driving two GPUs, hgaprec -gpus 2, 2-GPU ELBO), then the strong-scaling bench line at N=2 for the three workloads.
set -u
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/r02e_smi.txt
timeout 900 python -m pytest tests/test_multigpu.py tests/test_zz_elbo_gpu.py tests/test_cli.py -m gpu -v > gpurun_out/r02e_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02e_pytest_2gpu.log; grep -E "PASSED|FAILED|SKIPPED|ERROR|passed|failed" gpurun_out/r02e_pytest_2gpu.log | cut -c1-200 | tail -60
for w in netflix msd; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload $w --no-cpu-baseline > gpurun_out/r02e_bench_n2_$w.json 2> gpurun_out/r02e_bench_n2_$w.err; echo "bench $w exit $?"; tail -2 gpurun_out/r02e_bench_n2_$w.err; cut -c1-1800 gpurun_out/r02e_bench_n2_$w.json
done

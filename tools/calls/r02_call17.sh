#!/bin/bash
# 2-GPU: sharded beta update -- tests, then MSD N=2 with and without
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_multigpu.py -m gpu -q > gpurun_out/r02s_pytest_2gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/r02s_pytest_2gpu.log; tail -30 gpurun_out/r02s_pytest_2gpu.log | cut -c1-300
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload msd --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02s_n2_msd_$name.json 2> gpurun_out/r02s_n2_msd_$name.err
  echo "$name exit $? $(grep '^{' gpurun_out/r02s_n2_msd_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel_ms']; print(d['ms_per_step'], {a:round(b,3) for a,b in k.items()})")"
}
run shard0 HPF_SHARD_BETA=0
run shard1 HPF_SHARD_BETA=1

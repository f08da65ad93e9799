#!/bin/bash
# round 2, last call: MSD scale on 4 GPUs with the final library (the N=4 cell of the strong-scaling table)
set -u
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --workload msd --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02zz_bench_n4_msd.json 2> gpurun_out/r02zz_bench_n4_msd.err; echo "exit $?"
grep '^{' gpurun_out/r02zz_bench_n4_msd.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], {k: round(v,3) for k,v in d['roofline']['per_kernel_ms'].items()})"

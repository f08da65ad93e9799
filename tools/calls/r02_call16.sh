#!/bin/bash
# 2-GPU: where does the chunked all-reduce lose its time? (MSD scale)
set -u
mkdir -p gpurun_out
run() { # name, env...
  local name=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --workload msd --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/r02r_n2_msd_$name.json 2> gpurun_out/r02r_n2_msd_$name.err
  echo "$name exit $? $(grep '^{' gpurun_out/r02r_n2_msd_$name.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); k=d['roofline']['per_kernel_ms']; print(d['ms_per_step'], {a:round(b,3) for a,b in k.items()})")"
}
run c1_hi HPF_AR_CHUNKS=1
run c6_hi HPF_AR_CHUNKS=6
run c6_lo HPF_AR_CHUNKS=6 HPF_COMM_PRIO=0
run c6_defer HPF_AR_CHUNKS=6 HPF_AR_DEFER=1
run c3_hi HPF_AR_CHUNKS=3
run c6_hi_ch8 HPF_AR_CHUNKS=6 NCCL_MAX_NCHANNELS=8
run c6_hi_ch4 HPF_AR_CHUNKS=6 NCCL_MAX_NCHANNELS=4

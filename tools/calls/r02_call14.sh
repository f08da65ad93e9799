#!/bin/bash
# round 2, second 8-GPU call: strong scaling of the three BASELINE workloads on the final library, chunk variants at MSD scale
set -u
mkdir -p gpurun_out
run() { # tag nproc workload extra...
  local tag=$1 n=$2 w=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --no-cpu-baseline "$@" > gpurun_out/r02p_bench_n${n}_${w}${tag}.json 2> gpurun_out/r02p_bench_n${n}_${w}${tag}.err
  echo "bench n=$n $w $tag exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02p_bench_n${n}_${w}${tag}.err | tail -3
  grep '^{' gpurun_out/r02p_bench_n${n}_${w}${tag}.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('  value %.4g nnz/s  %.3f ms/step  e2e %.1f ms  chunks %s  per-kernel %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['item_chunks'], {k: round(v, 3) for k, v in d['roofline']['per_kernel_ms'].items()}))
if 'weak' in d: print('  weak: %.4g nnz/s %.3f ms/step' % (d['weak']['value'], d['weak']['ms_per_step']))
"
}
run "" 8 netflix
run "" 8 msd --no-extras
HPF_AR_CHUNKS=1 run _chunks1 8 msd --no-extras --e2e-steps 1
HPF_AR_CHUNKS=3 run _chunks3 8 msd --no-extras --e2e-steps 1
run "" 8 bpf-1b --e2e-steps 1
run "" 4 netflix --no-extras
run "" 4 msd --no-extras

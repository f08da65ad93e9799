#!/bin/bash
# round 2, last 8-GPU call: MSD scale strong scaling on the final library (faster update kernel, replicated beta update)
set -u
mkdir -p gpurun_out
run() { # tag nproc workload extra...
  local tag=$1 n=$2 w=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $n --workload $w --no-cpu-baseline "$@" > gpurun_out/r02y_bench_n${n}_${w}${tag}.json 2> gpurun_out/r02y_bench_n${n}_${w}${tag}.err
  echo "bench n=$n $w $tag exit $?"; grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/r02y_bench_n${n}_${w}${tag}.err | tail -3
  grep '^{' gpurun_out/r02y_bench_n${n}_${w}${tag}.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('  value %.4g nnz/s  %.3f ms/step  e2e %.1f ms  chunks %s sharded %s per-kernel %s' % (d['value'], d['ms_per_step'], d['e2e']['ms_per_step'], d['config']['item_chunks'], d['config'].get('beta_sharded'), {k: round(v, 3) for k, v in d['roofline']['per_kernel_ms'].items()}))
if 'weak' in d: print('  weak: %.4g nnz/s %.3f ms/step' % (d['weak']['value'], d['weak']['ms_per_step']))
"
}
run "" 8 msd --no-extras --e2e-steps 1

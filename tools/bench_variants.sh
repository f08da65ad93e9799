#!/bin/bash
# time the sweep kernels of every build/variants/lib_*.so (device-resident, Netflix scale), both plans
for lib in build/variants/lib_*.so; do
  for plan in "HPF_ITEM_TILE=0 HPF_HEAD_TILE=0" "HPF_PLAN=auto"; do
    env HPF_LIB=$PWD/$lib $plan timeout 600 python bench.py --no-cpu-baseline --e2e-steps 1 --steps 10 2>/dev/null | tail -1 | \
      python -c "import sys,json; d=json.loads(sys.stdin.read()); p=d['roofline']['per_kernel_ms']; print('%-28s %-34s step %.2f user %.2f (head %.2f) item %.2f' % ('$lib'.split('/')[-1], '$plan', d['ms_per_step'], p['sweep_user_ms'], p['sweep_user_head_ms'], p['sweep_item_ms']))"
  done
done

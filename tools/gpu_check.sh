#!/bin/bash
# Run on the B200 box through gpurun: GPU parity tests, one bench line, the ncu
# launch list of the same bench command and one --set full capture of the sweep.
# Everything lands in gpurun_out/ (scratch); summaries are copied to profiles/ by hand.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/smi.txt 2>&1
if [ "${SKIP_TESTS:-0}" != "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -15 gpurun_out/pytest_gpu.log
fi
timeout 900 python bench.py ${BENCH_ARGS:-} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "${SKIP_NCU:-0}" != "1" ]; then
  KREGEX='regex:sweep_kernel|update_kernel|combine_kernel|colsum|heldout|topn|head_kernel|head_reduce|split_kernel|split_aux_kernel|nnz_kernel|gamma_'
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KREGEX" -c 400 --csv \
      --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 \
      > gpurun_out/bench_under_ncu.log 2>&1
  echo "ncu list exit $?"
  timeout 1200 ncu --set full --clock-control none --import-source on -k "${NCU_FULL_K:-regex:sweep_kernel|update_kernel|head_kernel}" \
      --launch-skip ${NCU_SKIP:-14} -c ${NCU_COUNT:-7} -o gpurun_out/sweep_full -f \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full.log 2>&1
  echo "ncu full exit $?"
fi
ls -la gpurun_out

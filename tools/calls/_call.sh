mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -15 gpurun_out/pytest_gpu.log
python tools/sweep_experiments.py netflix '{"HPF_L2_TILE_MB": ["0","16","32","48","64"], "HPF_SWEEP_G": ["8"]}' > gpurun_out/exp_tiles.log 2>&1
cat gpurun_out/exp_tiles.log
python tools/sweep_experiments.py netflix '{"HPF_SEG_LEN": ["256","1024"], "HPF_SWEEP_G": ["4","8"]}' > gpurun_out/exp_seg.log 2>&1
cat gpurun_out/exp_seg.log

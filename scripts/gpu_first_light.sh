#!/bin/bash
# First GPU session: kernel tests (all, no -x), e2e small + full, logs into gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
python -m pytest tests/test_kernels_gpu.py -m gpu -q -rA --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1
tail -n 60 gpurun_out/kernels.log
python -m pytest tests/test_e2e_gpu.py -m gpu -q -rA -s --tb=short -p no:cacheprovider > gpurun_out/e2e.log 2>&1
tail -n 40 gpurun_out/e2e.log

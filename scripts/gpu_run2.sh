#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_kernels_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1; tail -n 8 gpurun_out/kernels.log
python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -rA -s --tb=short -p no:cacheprovider > gpurun_out/tc.log 2>&1; tail -n 60 gpurun_out/tc.log
python -m pytest tests/test_e2e_gpu.py -m gpu -q -rA -s --tb=short -p no:cacheprovider -k "fast_mode or sample" > gpurun_out/e2e_fast.log 2>&1; tail -n 30 gpurun_out/e2e_fast.log
python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; tail -c 3000 gpurun_out/bench_fp32.json; tail -n 5 gpurun_out/bench_fp32.err

#!/bin/bash
# cheaper CosKernel epilogue: kernel tests, e2e parity, all-pairs graph timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py tests/test_split_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -k "cos or gp or kde" > gpurun_out/pytest_cos.log 2>&1; tail -n 8 gpurun_out/pytest_cos.log
timeout 300 python scripts/allpairs_bench.py > gpurun_out/allpairs_bench.txt 2>&1; tail -n 2 gpurun_out/allpairs_bench.txt
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_e2e.log 2>&1; tail -n 6 gpurun_out/pytest_e2e.log

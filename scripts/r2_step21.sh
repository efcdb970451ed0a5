#!/bin/bash
# per-pixel prologue kernel on the tile grid after a tile pass: tests + flow sweep
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 -k "prologue or local_corr" > gpurun_out/pytest_lc.log 2>&1; tail -n 3 gpurun_out/pytest_lc.log
timeout 300 python scripts/lc_sweep.py > gpurun_out/lc_sweep.txt 2>&1; tail -n 6 gpurun_out/lc_sweep.txt
timeout 600 python -m pytest tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 500 -k "small_vs_reference or full_vs_reference" > gpurun_out/pytest_e2e_lc.log 2>&1; tail -n 2 gpurun_out/pytest_e2e_lc.log

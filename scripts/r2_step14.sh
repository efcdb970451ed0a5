#!/bin/bash
# tile-cooperative prologue, pitch rule + 8x2 tiles: tests, flow sweep A/B, ncu full capture of the tile kernels on smooth flow
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 -k "prologue or local_corr" > gpurun_out/pytest_lc.log 2>&1; tail -n 6 gpurun_out/pytest_lc.log
timeout 300 python scripts/lc_sweep.py > gpurun_out/lc_sweep.txt 2>&1; tail -n 4 gpurun_out/lc_sweep.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"refiner_prologue_tile" --launch-skip 10 --launch-count 5 -o gpurun_out/ncu_lc_tile -f python scripts/lc_sweep.py > gpurun_out/ncu_lc_tile.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_lc_tile.ncu-rep > gpurun_out/ncu_lc_tile.txt; grep -c . gpurun_out/ncu_lc_tile.txt

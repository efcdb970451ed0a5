#!/bin/bash
mkdir -p gpurun_out
python scripts/gemm_bench.py all > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_fc1 -f python scripts/gemm_bench.py fc1 > gpurun_out/ncu_fc1.log 2>&1; tail -n 3 gpurun_out/ncu_fc1.log
ls -la gpurun_out/*.ncu-rep

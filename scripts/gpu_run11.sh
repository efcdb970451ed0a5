#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/gemm_bench.py ref > gpurun_out/gemm_bench_ref.log 2>&1; cat gpurun_out/gemm_bench_ref.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/prof_ref2 -f python scripts/gemm_bench.py ref2 > gpurun_out/ncu_ref2.log 2>&1; tail -n 2 gpurun_out/ncu_ref2.log

#!/bin/bash
mkdir -p gpurun_out
for epi in 0 1; do
ROMAB200_GEMM_EPI=$epi ROMAB200_GEMM_PAIR=0 timeout 300 python scripts/gemm_clk.py 2>&1 | tee gpurun_out/gemm_clk_epi${epi}.log
done
ROMAB200_GEMM_EPI=0 ROMAB200_GEMM_PAIR=0 timeout 300 python scripts/gemm_bench2.py > gpurun_out/gemm_bench_epi0.log 2>&1
ROMAB200_GEMM_EPI=1 ROMAB200_GEMM_PAIR=0 timeout 300 python scripts/gemm_bench2.py > gpurun_out/gemm_bench_epi1.log 2>&1
paste -d'|' gpurun_out/gemm_bench_epi0.log gpurun_out/gemm_bench_epi1.log | cut -c1-36,53-84,137-168

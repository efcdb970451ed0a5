#!/bin/bash
mkdir -p gpurun_out
ROMAB200_GEMM_PAIR=0 timeout 300 python scripts/gemm_clk.py 2>&1 | tee gpurun_out/gemm_clk_pair0.log
ROMAB200_GEMM_PAIR=1 timeout 300 python scripts/gemm_clk.py 2>&1 | tee gpurun_out/gemm_clk_pair1.log

#!/bin/bash
mkdir -p gpurun_out
for mode in split fp16; do for pair in 0 1; do
ROMAB200_GEMM_PAIR=$pair timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc --launch-skip 4 --launch-count 2 -o gpurun_out/ncu_gemm_${mode}_pair${pair} -f python scripts/gemm_prof.py $mode > gpurun_out/ncu_gemm_${mode}_pair${pair}.log 2>&1
done; done
ls -la gpurun_out/*.ncu-rep

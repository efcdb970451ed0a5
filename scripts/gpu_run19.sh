#!/bin/bash
mkdir -p gpurun_out
# (a) ncu --set full on three representative kernels: ViT fc1 GEMM, all-pairs CosKernel GEMM, local-correlation prologue
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 2 -c 1 -o gpurun_out/r01_ncu_gemm_fc1 -f python scripts/gemm_bench.py fc1 > gpurun_out/ncu_a.log 2>&1; tail -n 1 gpurun_out/ncu_a.log
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"refiner_prologue_kernel|local_corr" -c 4 -o gpurun_out/r01_ncu_localcorr -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_b.log 2>&1; tail -n 1 gpurun_out/ncu_b.log
# (b) bench without sample to size sample()'s share
timeout 600 python bench.py --precision fp16 --steps 5 --warmup 3 --no-cpu-baseline --no-sample > gpurun_out/bench_fp16_nosample.json 2> gpurun_out/bench_ns.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16_nosample.json')); print('no-sample', d['value'], d['ms_per_step'], d['e2e']['value'])"
ls -la gpurun_out/*.ncu-rep

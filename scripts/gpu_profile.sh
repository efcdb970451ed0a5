#!/bin/bash
# ncu evidence for profiles/ (one GPU; never a multi-rank command).  Keep gpurun_out/ under the runner's 64 MiB copy-back
# limit: a full-set capture with sources is ~1 MB per kernel launch.
mkdir -p gpurun_out
# 1. launch list of one eager match() (serialised, cold-cache kernel times)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file gpurun_out/launches_fp16_one_pass.csv python scripts/profile_one_pass.py fp16 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/launches_fp16_one_pass.csv > gpurun_out/launches_fp16_one_pass.txt
# 2. launch list of the bench command itself (first launches only: ncu serialises every kernel, ~0.3 s each)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_bench_cmd.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
python scripts/launch_table.py gpurun_out/launches_bench_cmd.csv > gpurun_out/launches_bench_cmd.txt
# 3. full-set captures: last ViT block's GEMMs + the CosKernel GEMMs, then one launch each of the CUDA-core kernels
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"gemm_tc_kernel" \
    --launch-skip 92 --launch-count 10 -o gpurun_out/ncu_gemm -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_gemm.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off \
    -k regex:"refiner_prologue_kernel|refiner_block_c144|dwconv5x5_relu_tma|chol_block128|refiner_block_small|flash_attn" \
    --launch-skip 24 --launch-count 24 -o gpurun_out/ncu_others -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_others.log 2>&1
for r in gpurun_out/ncu_gemm.ncu-rep gpurun_out/ncu_others.ncu-rep; do python scripts/ncu_summary.py $r --unique > ${r%.ncu-rep}.txt; done

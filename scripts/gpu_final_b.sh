#!/bin/bash
# profiles: ncu launch list of the bench command, one-pass launch list, ncu --set full captures of the kernels DESIGN.md discusses
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r01_launches_bench_cmd.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
python scripts/launch_table.py gpurun_out/r01_launches_bench_cmd.csv > gpurun_out/r01_launches_bench_cmd.txt 2>/dev/null; head -n 12 gpurun_out/r01_launches_bench_cmd.txt
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r01_launches_fp16_one_pass.csv python scripts/profile_one_pass.py fp16 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/r01_launches_fp16_one_pass.csv > gpurun_out/r01_launches_fp16_one_pass.txt 2>/dev/null; head -n 30 gpurun_out/r01_launches_fp16_one_pass.txt
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"gemm_tc_kernel" --launch-skip 88 --launch-count 22 -o gpurun_out/r01_gemm -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_gemm.log 2>&1; tail -n 1 gpurun_out/ncu_gemm.log
timeout 600 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"refiner_prologue_kernel|refiner_block_c144|dwconv5x5_relu_tma|chol_block128|refiner_block_small|flash_attn" --launch-count 110 -o gpurun_out/r01_others -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_others.log 2>&1; tail -n 1 gpurun_out/ncu_others.log

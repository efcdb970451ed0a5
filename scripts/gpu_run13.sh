#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/tc.log 2>&1; echo "tc rc=$?"; tail -n 3 gpurun_out/tc.log
timeout 300 python scripts/gemm_bench.py all > gpurun_out/gemm_bench.log 2>&1; cat gpurun_out/gemm_bench.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_fp16.csv python scripts/profile_one_pass.py fp16 > gpurun_out/prof_pass.log 2>&1; tail -n 1 gpurun_out/prof_pass.log
python scripts/launch_table.py gpurun_out/launches_r1_fp16.csv | head -28
python scripts/launch_table.py gpurun_out/launches_r1_fp16.csv grid | grep -E "dwconv|block_small|prologue|tail" | head -30
timeout 600 python bench.py --precision fp16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(d['stage_ms_per_step']); print(d['roofline']['achieved'], d['roofline']['frac'])"; tail -n 3 gpurun_out/bench_fp16.err

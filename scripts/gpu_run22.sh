#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/kernels.log 2>&1; tail -n 4 gpurun_out/kernels.log
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/e2e.log 2>&1; grep -E "^\[|passed|failed|Error" gpurun_out/e2e.log | tail -n 12
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_fp16.csv python scripts/profile_one_pass.py fp16 > gpurun_out/prof_pass.log 2>&1; tail -n 1 gpurun_out/prof_pass.log
python scripts/launch_table.py gpurun_out/launches_r1_fp16.csv 2>/dev/null | head -14
timeout 600 python bench.py --precision fp16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -n 3 gpurun_out/bench_fp16.err
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32.json 2> gpurun_out/bench_fp32.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp32.json')); print('fp32', d['value'], d['ms_per_step'], d['e2e']['value'], d['gemm_backends'])"

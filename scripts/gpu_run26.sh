#!/bin/bash
mkdir -p gpurun_out
timeout 120 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "c144" > gpurun_out/c144.log 2>&1; echo "c144 rc=$?"; tail -n 8 gpurun_out/c144.log
nvidia-smi --query-gpu=name,memory.used --format=csv,noheader
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -s --tb=short -p no:cacheprovider > gpurun_out/e2e.log 2>&1; grep -E "^\[|passed|failed|Error" gpurun_out/e2e.log | tail -n 12
timeout 600 python bench.py --precision fp16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v for k,v in d['stage_ms_per_step'].items() if k.startswith('refine')})"; tail -n 3 gpurun_out/bench_fp16.err

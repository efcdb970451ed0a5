#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py tests/test_e2e_gpu.py -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/e2e.log 2>&1; grep -E "^\[|passed|failed|Error" gpurun_out/e2e.log | tail -n 6
timeout 600 python bench.py --precision fp16 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:v for k,v in d['stage_ms_per_step'].items() if v>0.25})
for t in d['top_gemm_shapes'][:10]: print(t)"

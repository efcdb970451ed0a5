#!/bin/bash
# last validation of the round: full GPU suite + headline bench (no CPU baseline), one small ncu capture of the fc1 GEMM
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -n 2 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('fp16', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']); print(d['roofline']['achieved'], d['roofline']['frac'])"

#!/bin/bash
# final validation: full GPU test suite, smoke(), bench lines for profiles/
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/r01_bench_fp16_n1.json 2> gpurun_out/bench_fp16.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_fp16_n1.json')); print('fp16', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['clocks']); print(d['roofline']['achieved'], d['roofline']['frac']); print(d['roofline_kernels']); print(d['cpu_baseline'])"
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_fp32_n1.json 2> gpurun_out/bench_fp32.err; python -c "
import json; d=json.load(open('gpurun_out/r01_bench_fp32_n1.json')); print('fp32', d['value'], d['ms_per_step'], d['e2e']['value'])"

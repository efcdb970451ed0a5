#!/bin/bash
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k gp_solve > gpurun_out/gp.log 2>&1; echo "gp rc=$?"; tail -n 3 gpurun_out/gp.log
timeout 300 python scripts/overlap_test.py 2>&1 | tail -8
timeout 600 python bench.py --precision fp16 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['eager_ms_per_step'])"; tail -n 3 gpurun_out/bench_fp16.err

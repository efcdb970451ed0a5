#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_e2e_gpu.py -m gpu -q -rA -s --tb=short -p no:cacheprovider -k "fast_mode_full" > gpurun_out/e2e_fast_full.log 2>&1; tail -n 12 gpurun_out/e2e_fast_full.log
python bench.py --precision fp16 --steps 5 --warmup 3 > gpurun_out/bench_fp16.json 2> gpurun_out/bench_fp16.err; tail -c 4000 gpurun_out/bench_fp16.json; tail -n 5 gpurun_out/bench_fp16.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp16.csv python bench.py --precision fp16 --steps 1 --warmup 1 --no-cpu-baseline --no-sample > gpurun_out/ncu_bench.log 2>&1; tail -n 3 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches_fp16.csv

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_fp16_n2.json 2> gpurun_out/bench_n2.err; tail -c 1500 gpurun_out/bench_fp16_n2.json; tail -n 5 gpurun_out/bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 1 --warmup 0 --impl reference > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err & 
timeout 600 python bench.py --precision fp16 --steps 5 --warmup 3 --no-cpu-baseline --pairs-per-gpu 8 > gpurun_out/bench_fp16_p8.json 2> gpurun_out/bench_p8.err; python -c "
import json; d=json.load(open('gpurun_out/bench_fp16_p8.json')); print('P=8', d['value'], d['ms_per_step'], d['e2e']['value'])"; tail -n 3 gpurun_out/bench_p8.err
timeout 400 python scripts/cpu_threads_probe.py 2>&1 | tail -6
wait
tail -c 600 gpurun_out/bench_ref_n2.json

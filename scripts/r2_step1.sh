#!/bin/bash
# round-2 step 1: split-fp16 (RB_F16S) tensor-core parity mode — kernel tests, e2e parity, first bench numbers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt
timeout 600 python -m pytest tests/test_split_gpu.py -q -x -p no:cacheprovider --timeout 180 > gpurun_out/split.log 2>&1; tail -n 15 gpurun_out/split.log
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider --timeout 400 -s > gpurun_out/e2e.log 2>&1; grep -E "max-abs|passed|failed|Error|error|\{'proj" gpurun_out/e2e.log | tail -n 40
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py -q -p no:cacheprovider --timeout 180 > gpurun_out/kernels.log 2>&1; tail -n 5 gpurun_out/kernels.log
timeout 600 python bench.py --precision fp32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split_n1.json 2> gpurun_out/bench_split.err; tail -n 3 gpurun_out/bench_split.err
python - <<'PY'
import json
try:
    d = json.load(open("gpurun_out/bench_split_n1.json"))
    print("split", d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["roofline"]["frac"] if d["roofline"] else None)
    print({k: v for k, v in list(d["stage_ms_per_step"].items())[:24]})
    print(d["gemm_backends"])
    for s in d["top_gemm_shapes"][:14]: print(s)
except Exception as e:
    print("bench parse failed", e)
PY

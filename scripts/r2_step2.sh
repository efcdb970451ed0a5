#!/bin/bash
# round-2 step 2: split flash attention, fp32 TMA dwconv, fp32 fused small block; batch-8 and fp16 comparison lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_split_gpu.py -q -p no:cacheprovider --timeout 180 -s > gpurun_out/split.log 2>&1; grep -E "flash split|passed|failed|Error" gpurun_out/split.log | tail -n 15
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider --timeout 400 -s > gpurun_out/e2e.log 2>&1; grep -E "max-abs|passed|failed|Error|error|\{'proj" gpurun_out/e2e.log | tail -n 30
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_gemm_tc_gpu.py -q -p no:cacheprovider --timeout 180 > gpurun_out/kernels.log 2>&1; tail -n 3 gpurun_out/kernels.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_split_n1.json 2> gpurun_out/bench_split.err; tail -n 3 gpurun_out/bench_split.err
timeout 600 python bench.py --steps 3 --warmup 3 --pairs-per-gpu 8 --no-cpu-baseline --no-fast-mode > gpurun_out/bench_split_p8.json 2> gpurun_out/bench_split_p8.err; tail -n 3 gpurun_out/bench_split_p8.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_split_n1.json", "gpurun_out/bench_split_p8.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "frac", d["roofline"]["frac"] if d["roofline"] else None, "parity", d.get("parity"), "cpu", d.get("cpu_baseline", {}) and d["cpu_baseline"]["value"])
        if d.get("fast_mode"): print("  fast", d["fast_mode"]["value"], d["fast_mode"]["parity"])
        print("  ", {k: v for k, v in list(d["stage_ms_per_step"].items())[:26]})
        print("  ", d["gemm_backends"])
        for s in d["top_gemm_shapes"][:10]: print("  ", s)
    except Exception as e:
        print(f, "parse failed", e)
PY

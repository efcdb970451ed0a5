#!/bin/bash
# What a round's GPU validation runs on the B200 box (through `gpurun -- 'bash scripts/gpu_validate.sh'`):
# the full GPU test suite, smoke(), and the bench lines kept under profiles/.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/bench_fp16_n1.json 2> gpurun_out/bench_fp16.err
timeout 600 python bench.py --precision fp32 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_fp32_n1.json 2> gpurun_out/bench_fp32.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_fp16_n1.json", "gpurun_out/bench_fp32_n1.json"):
    d = json.load(open(f))
    print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d["gpu_launches"], d["clocks"], d["roofline"]["frac"] if d["roofline"] else None)
PY

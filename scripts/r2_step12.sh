#!/bin/bash
# re-entry validation of HEAD: whole GPU suite, smoke(), default bench line
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -n 5 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err
tail -c 600 gpurun_out/bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["frac"], d["roofline"].get("frac_executed"))
for k in d.get("roofline_kernels", []): print(k)
print(d.get("stage_ms_per_step"))
PY

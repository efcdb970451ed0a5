#!/bin/bash
# the default bench line of the final tree (kept as profiles/r02_bench_fp32split_n1_final.json)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err; tail -c 200 gpurun_out/bench_default_n1.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "parity", d["parity"]["warp"], d["parity"]["certainty"], "launches", d["gpu_launches"], d["clocks"])
print("roofline", d["roofline"]["frac"], [(k["kernel"][:30], round(k["frac"], 4)) for k in d["roofline_kernels"]], "cpu", d["cpu_baseline"]["value"], "fast", d["fast_mode"]["value"])
PY

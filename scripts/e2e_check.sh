#!/bin/bash
# short bench: e2e with asynchronous read-back
mkdir -p gpurun_out
timeout 500 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_e2e_async.json 2> gpurun_out/bench_e2e_async.err; tail -c 300 gpurun_out/bench_e2e_async.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_e2e_async.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"], "parity", d["parity"]["warp"], d["parity"]["certainty"])
PY

#!/bin/bash
# 8x4 row-per-thread tile kernel, stride-16 table path, all-pairs graph micro-benchmark, e2e parity, short bench
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 -k "prologue or local_corr" > gpurun_out/pytest_lc.log 2>&1; tail -n 6 gpurun_out/pytest_lc.log
timeout 300 python scripts/lc_sweep.py > gpurun_out/lc_sweep.txt 2>&1; tail -n 4 gpurun_out/lc_sweep.txt
timeout 300 python scripts/allpairs_bench.py > gpurun_out/allpairs_bench.txt 2>&1; tail -n 3 gpurun_out/allpairs_bench.txt
timeout 900 python -m pytest tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_e2e.log 2>&1; tail -n 4 gpurun_out/pytest_e2e.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["warp"], d["parity"]["certainty"])
print({k: v for k, v in d["stage_ms_per_step"].items() if "prologue" in k or "gp" in k})
PY

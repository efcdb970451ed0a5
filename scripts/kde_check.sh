#!/bin/bash
# symmetric KDE: kernel test, timing, sample() statistics, short bench
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --timeout 150 -k "kde" > gpurun_out/pytest_kde.log 2>&1; tail -n 4 gpurun_out/pytest_kde.log
timeout 100 python scripts/kde_time.py > gpurun_out/kde_time.txt 2>&1; tail -n 4 gpurun_out/kde_time.txt
timeout 200 python -m pytest tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 150 -k "sample_statistics or sample_distribution" > gpurun_out/pytest_sample.log 2>&1; tail -n 3 gpurun_out/pytest_sample.log
timeout 200 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_kde_sym.json 2> gpurun_out/bench_kde_sym.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_kde_sym.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["warp"], d["parity"]["certainty"])
PY

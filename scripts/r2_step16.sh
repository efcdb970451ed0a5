#!/bin/bash
# KDE splits, engine-mode flow sweep, bench with the graph-timed all-pairs leg
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -k "kde or sample or prologue" > gpurun_out/pytest_kde.log 2>&1; tail -n 6 gpurun_out/pytest_kde.log
timeout 300 python scripts/lc_sweep.py > gpurun_out/lc_sweep.txt 2>&1; tail -n 6 gpurun_out/lc_sweep.txt
timeout 300 python scripts/sample_profile.py > gpurun_out/sample_profile.txt 2>&1; tail -n 12 gpurun_out/sample_profile.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -c 400 gpurun_out/bench_quick.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_quick.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"]["warp"], d["parity"]["certainty"])
print(json.dumps(d["roofline"])[:900])
for k in d["roofline_kernels"]:
    k = dict(k); k.pop("flow_sweep_other_kernels", None)
    print(json.dumps(k)[:1800])
PY

#!/bin/bash
# What a round's GPU validation runs on the B200 box (through `gpurun -- 'bash scripts/gpu_validate.sh'`): the whole GPU suite, smoke(),
# the default bench line (parity mode) and the stand-alone micro-benchmarks whose numbers DESIGN.md quotes.  scripts/r2_final.sh adds the
# ncu launch lists and full captures kept under profiles/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err
timeout 300 python scripts/lc_sweep.py > gpurun_out/lc_sweep.txt 2>&1
timeout 300 python scripts/allpairs_bench.py > gpurun_out/allpairs_bench.txt 2>&1
timeout 300 python scripts/kde_time.py > gpurun_out/kde_time.txt 2>&1
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"], "launches", d["gpu_launches"], d["clocks"])
print("roofline", d["roofline"]["frac"], [(k["kernel"][:30], k["frac"]) for k in d["roofline_kernels"]])
PY

#!/bin/bash
# split GEMMs on 128-wide CTA-pair tiles (double-buffered accumulators): tests + bench A/B
mkdir -p gpurun_out
ROMAB200_GEMM_SPLIT_PAIR_BN=128 ROMAB200_GEMM_PAIR=2 timeout 600 python -m pytest tests/test_split_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 > gpurun_out/pytest_pair128.log 2>&1; tail -n 4 gpurun_out/pytest_pair128.log
for v in 0 128; do
ROMAB200_GEMM_SPLIT_PAIR_BN=$v timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_pairbn$v.json 2> gpurun_out/bench_pairbn$v.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_pairbn$v.json").read().strip().splitlines()[-1])
    st = d["stage_ms_per_step"]
    print("split pair bn $v: value", round(d["value"], 3), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3), "parity", d["parity"]["warp"], d["parity"]["certainty"],
          "| roofline frac", round(d["roofline"]["frac"], 4), "dinov2", st.get("dinov2"), "dec.blocks", st.get("  dec.blocks"))
except Exception as e:
    print("$v failed", e); print(open("gpurun_out/bench_pairbn$v.err").read()[-800:])
PY
done

#!/bin/bash
# persistent-grid cap of the side stream's (VGG) GEMMs: sweep of ROMAB200_SIDE_CTAS
mkdir -p gpurun_out
for c in 0 132 116 100 84; do
ROMAB200_SIDE_CTAS=$c timeout 400 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_side$c.json 2> gpurun_out/bench_side$c.err
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/bench_side$c.json").read().strip().splitlines()[-1])
    st = d["stage_ms_per_step"]
    print("side_ctas $c: value", round(d["value"], 3), "ms", round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 3), "parity", d["parity"]["certainty"],
          "| eager: gp.solve", st.get("  gp.solve"), "gp+decoder", st.get("gp+decoder"), "vgg.lo", st.get("vgg.lo"), "vgg.up", st.get("vgg.up"))
except Exception as e:
    print("side_ctas $c failed", e); print(open("gpurun_out/bench_side$c.err").read()[-500:])
PY
done

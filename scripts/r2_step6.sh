#!/bin/bash
# round-2 step 6: GP solve on tensor cores, 16-bit per-kernel tests, validation shim, rect golden, library baseline, batch-8, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider --timeout 400 > gpurun_out/pytest_gpu.log 2>&1; tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_split_n1.json 2> gpurun_out/bench_split.err; tail -n 3 gpurun_out/bench_split.err
timeout 600 python bench.py --steps 3 --warmup 3 --pairs-per-gpu 8 --global-pairs 8 --no-cpu-baseline --no-fast-mode --no-library-baseline > gpurun_out/bench_split_p8.json 2> gpurun_out/bench_split_p8.err; tail -n 3 gpurun_out/bench_split_p8.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_split_n1.json", "gpurun_out/bench_split_p8.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "frac", d["roofline"]["frac"] if d["roofline"] else None, "parity", d.get("parity", {}).get("warp"), d.get("parity", {}).get("certainty"))
        if d.get("fast_mode"): print("  fast", d["fast_mode"]["value"])
        if d.get("gpu_library_baseline"):
            lb = d["gpu_library_baseline"]
            print("  library", {k: (round(v["value"], 2), v["stage_ms_per_step"]) for k, v in lb.items() if isinstance(v, dict)} if "error" not in lb else lb)
        if d.get("cpu_baseline"): print("  cpu", d["cpu_baseline"]["value"])
        print("  ", {k: v for k, v in list(d["stage_ms_per_step"].items())[:28]})
        print("  ", d["gemm_backends"])
    except Exception as e:
        print(f, "parse failed", e)
PY
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_split_one_pass.csv python scripts/profile_one_pass.py fp32 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/launches_split_one_pass.csv > gpurun_out/launches_split_one_pass.txt; head -n 45 gpurun_out/launches_split_one_pass.txt

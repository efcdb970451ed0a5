#!/bin/bash
# round-2 step 5: TMA-store epilogue (UTMASTG / reduce-add) — correctness in both tile schedules, phase clocks, micro-bench, bench
mkdir -p gpurun_out
ROMAB200_GEMM_PAIR=2 timeout 600 python -m pytest tests/test_split_gpu.py tests/test_gemm_tc_gpu.py -q -x -p no:cacheprovider --timeout 120 > gpurun_out/pair_forced.log 2>&1; tail -n 6 gpurun_out/pair_forced.log
ROMAB200_GEMM_PAIR=0 timeout 600 python -m pytest tests/test_split_gpu.py tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py -q -x -p no:cacheprovider --timeout 120 > gpurun_out/kernels.log 2>&1; tail -n 6 gpurun_out/kernels.log
ROMAB200_GEMM_EPI=2 ROMAB200_GEMM_PAIR=0 timeout 300 python scripts/gemm_clk.py 2>&1 | tee gpurun_out/gemm_clk_epi2.log
ROMAB200_GEMM_EPI=2 ROMAB200_GEMM_PAIR=1 timeout 300 python scripts/gemm_clk.py 2>&1 | tee gpurun_out/gemm_clk_epi2_pair.log
ROMAB200_GEMM_EPI=0 ROMAB200_GEMM_PAIR=1 timeout 300 python scripts/gemm_bench2.py > gpurun_out/gemm_bench_epi0.log 2>&1
ROMAB200_GEMM_EPI=2 ROMAB200_GEMM_PAIR=1 timeout 300 python scripts/gemm_bench2.py > gpurun_out/gemm_bench_epi2.log 2>&1
paste -d'|' gpurun_out/gemm_bench_epi0.log gpurun_out/gemm_bench_epi2.log | cut -c1-36,53-84,137-168
timeout 900 python -m pytest tests/test_e2e_gpu.py -q -p no:cacheprovider --timeout 400 -s > gpurun_out/e2e.log 2>&1; grep -E "full|passed|failed|Error|error" gpurun_out/e2e.log | tail -n 8
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_split_n1.json 2> gpurun_out/bench_split.err; tail -n 3 gpurun_out/bench_split.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_split_n1.json",):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "frac", d["roofline"]["frac"] if d["roofline"] else None, "parity", d.get("parity", {}).get("warp"), d.get("parity", {}).get("certainty"))
        if d.get("fast_mode"): print("  fast", d["fast_mode"]["value"])
        print("  ", {k: v for k, v in list(d["stage_ms_per_step"].items())[:26]})
        print("  ", d["gemm_backends"])
    except Exception as e:
        print(f, "parse failed", e)
PY

#!/bin/bash
# round-2 validation call: full GPU suite, smoke, default bench (all legs), A/B of the two new knobs, launch list and targeted ncu captures
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 > gpurun_out/pytest_gpu_full.log 2>&1; tail -n 12 gpurun_out/pytest_gpu_full.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 4 gpurun_out/smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err; tail -n 3 gpurun_out/bench_final_n1.err
ROMAB200_FA_HALVES=2 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_halves2.json 2> gpurun_out/bench_halves2.err
ROMAB200_GEMM_WAVE=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_wave0.json 2> gpurun_out/bench_wave0.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_final_n1.json", "gpurun_out/bench_halves2.json", "gpurun_out/bench_wave0.json"):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "frac", d["roofline"]["frac"] if d["roofline"] else None, "parity", d.get("parity", {}).get("warp"), d.get("parity", {}).get("certainty"))
        if d.get("fast_mode"): print("  fast", d["fast_mode"].get("value"))
        if d.get("preprocess"): print("  preprocess", d["preprocess"])
        st = d.get("stage_ms_per_step", {})
        print("  stages", {k: v for k, v in list(st.items())[:14]})
    except Exception as e:
        print(f, "parse failed", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_final_one_pass.csv python scripts/profile_one_pass.py fp32 > gpurun_out/launches_final.log 2>&1
ROMAB200_FA_HALVES=2 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:flash_attn --csv --log-file gpurun_out/launches_flash_halves2.csv python scripts/profile_one_pass.py fp32 > gpurun_out/launches_flash2.log 2>&1
for spec in "flash_attn_split:3" "dwconv5x5_relu_tma:40" "refiner_block_small_f32:12" "refiner_prologue_small:1" "refiner_prologue_kernel:4"; do
  k=${spec%%:*}; skip=${spec##*:}
  timeout 400 ncu --set full --clock-control none --profile-from-start off -k regex:$k --launch-skip $skip --launch-count 1 -o gpurun_out/ncu_final_$k -f python scripts/profile_one_pass.py fp32 > gpurun_out/ncu_final_$k.log 2>&1
done
ls -la gpurun_out/*.ncu-rep; du -sh gpurun_out

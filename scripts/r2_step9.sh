#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 600 -k "sample or weighted" > gpurun_out/pytest_sample.log 2>&1; tail -n 8 gpurun_out/pytest_sample.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_split_n1.json 2> gpurun_out/bench_split.err; tail -n 3 gpurun_out/bench_split.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_split_n1.json",):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "frac", d["roofline"]["frac"] if d["roofline"] else None, "parity", d.get("parity", {}).get("warp"), d.get("parity", {}).get("certainty"))
        if d.get("fast_mode"): print("  fast", d["fast_mode"]["value"])
    except Exception as e:
        print(f, "parse failed", e)
PY
for mode in split fp16; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc --launch-skip 4 --launch-count 2 -o gpurun_out/ncu_gemm_${mode}_final -f python scripts/gemm_prof.py $mode > gpurun_out/ncu_gemm_${mode}_final.log 2>&1
done
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:"flash_attn_split|dwconv5x5_relu_tma|refiner_prologue|refiner_block_small_f32|chol_block128" --launch-skip 24 --launch-count 14 -o gpurun_out/ncu_others_final -f python scripts/profile_one_pass.py fp32 > gpurun_out/ncu_others_final.log 2>&1
ls -la gpurun_out/*final*.ncu-rep; du -sh gpurun_out

#!/bin/bash
# round-2 step 8: device sampler + full suite, bench with / without sample(), ncu --set full of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -n 12 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_split_n1.json 2> gpurun_out/bench_split.err; tail -n 3 gpurun_out/bench_split.err
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline --no-fast-mode --no-sample > gpurun_out/bench_split_n1_nosample.json 2> gpurun_out/bench_split_ns.err; tail -n 3 gpurun_out/bench_split_ns.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_split_n1.json", "gpurun_out/bench_split_n1_nosample.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "launches", d["gpu_launches"],
              "frac", d["roofline"]["frac"] if d["roofline"] else None, "parity", d.get("parity", {}).get("warp"), d.get("parity", {}).get("certainty"))
        if d.get("fast_mode"): print("  fast", d["fast_mode"]["value"], d["fast_mode"]["parity"])
    except Exception as e:
        print(f, "parse failed", e)
PY
for mode in split fp16; do
timeout 300 ncu --set full --import-source on --clock-control none -k regex:gemm_tc --launch-skip 4 --launch-count 2 -o gpurun_out/ncu_gemm_${mode}_final -f python scripts/gemm_prof.py $mode > gpurun_out/ncu_gemm_${mode}_final.log 2>&1
done
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"flash_attn_split|dwconv5x5_relu_tma|refiner_prologue|refiner_block_small_f32|chol_block128|gemm_tc_pair_kernel<192" --launch-skip 20 --launch-count 44 -o gpurun_out/ncu_others_final -f python scripts/profile_one_pass.py fp32 > gpurun_out/ncu_others_final.log 2>&1
ls -la gpurun_out/*final*.ncu-rep

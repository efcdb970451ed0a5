#!/bin/bash
# flash attention with double-buffered probabilities: kernel tests, e2e parity, bench A/B of ROMAB200_FA_HALVES
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_split_gpu.py tests/test_gemm_tc_gpu.py -m gpu -q -p no:cacheprovider --timeout 300 -k "flash or attention" > gpurun_out/pytest_flash.log 2>&1; tail -n 6 gpurun_out/pytest_flash.log
timeout 600 python -m pytest tests/test_e2e_gpu.py -m gpu -q -p no:cacheprovider --timeout 500 -k "small_vs_reference or full_vs_reference or fast_mode" > gpurun_out/pytest_e2e_flash.log 2>&1; tail -n 4 gpurun_out/pytest_e2e_flash.log
for h in 1 2; do
ROMAB200_FA_HALVES=$h timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-library-baseline > gpurun_out/bench_fa_halves$h.json 2> gpurun_out/bench_fa_halves$h.err
ROMAB200_FA_HALVES=$h timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:flash_attn --csv --log-file gpurun_out/launches_fa_halves$h.csv python scripts/profile_one_pass.py fp32 > /dev/null 2>&1
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -k regex:flash_attn --csv --log-file gpurun_out/launches_fa_fp16.csv python scripts/profile_one_pass.py fp16 > /dev/null 2>&1
python - <<'PY'
import json, csv
for h in (1, 2):
    f = f"gpurun_out/bench_fa_halves{h}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "parity", d["parity"]["warp"], d["parity"]["certainty"], "fast", (d.get("fast_mode") or {}).get("value"))
        print("   attn.vit", d["stage_ms_per_step"].get("  attn.vit"), "dinov2", d["stage_ms_per_step"].get("dinov2"))
    except Exception as e:
        print(f, "parse failed", e); print(open(f.replace(".json", ".err")).read()[-600:])
for name in ("fa_halves1", "fa_halves2", "fa_fp16"):
    try:
        rows = list(csv.reader(open(f"gpurun_out/launches_{name}.csv")))
        hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
        mi = rows[hdr].index("Metric Value")
        v = [float(r[mi].replace(",", "")) for r in rows[hdr + 1:] if len(r) > mi]
        print(name, "flash launches", len(v), "avg us", sum(v) / len(v) / 1e3)
    except Exception as e:
        print(name, "failed", e)
PY

#!/bin/bash
# round-2 final state on one B200: whole GPU suite, smoke(), the default bench line, ncu launch lists (one eager pass; the bench command itself),
# ncu --set full of the local-correlation kernels and of the all-pairs launches
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; tail -n 3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 3
timeout 900 python bench.py > gpurun_out/bench_default_n1.json 2> gpurun_out/bench_default_n1.err; tail -c 300 gpurun_out/bench_default_n1.err
timeout 300 python scripts/lc_sweep.py > gpurun_out/lc_sweep.txt 2>&1
timeout 300 python scripts/allpairs_bench.py > gpurun_out/allpairs_bench.txt 2>&1; tail -n 1 gpurun_out/allpairs_bench.txt
timeout 300 python scripts/kde_time.py > gpurun_out/kde_time.txt 2>&1
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fp32split_one_pass.csv python scripts/profile_one_pass.py fp32 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/launches_fp32split_one_pass.csv > gpurun_out/launches_fp32split_one_pass.txt; head -n 12 gpurun_out/launches_fp32split_one_pass.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches_bench_cmd.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-library-baseline --no-fast-mode > gpurun_out/bench_under_ncu.json 2> gpurun_out/bench_under_ncu.err
python scripts/launch_table.py gpurun_out/launches_bench_cmd.csv > gpurun_out/launches_bench_cmd.txt; head -n 6 gpurun_out/launches_bench_cmd.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"refiner_prologue" --launch-skip 10 --launch-count 22 -o gpurun_out/ncu_lc_final -f python scripts/lc_sweep.py > gpurun_out/ncu_lc_final.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_lc_final.ncu-rep --unique > gpurun_out/ncu_lc_final.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc" --launch-skip 9 --launch-count 3 -o gpurun_out/ncu_allpairs_final -f python scripts/allpairs_bench.py > gpurun_out/ncu_allpairs_final.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_allpairs_final.ncu-rep > gpurun_out/ncu_allpairs_final.txt
rm -f gpurun_out/launches_bench_cmd.csv gpurun_out/launches_fp32split_one_pass.csv
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_default_n1.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "parity", d["parity"], "launches", d["gpu_launches"])
print("roofline", d["roofline"]["frac"], "all-pairs", d["roofline_kernels"][0]["frac"], d["roofline_kernels"][0]["ms_per_step"], "lc", d["roofline_kernels"][1]["frac"], d["roofline_kernels"][1].get("flow_sweep_hbm_frac"))
print("cpu", d.get("cpu_baseline"), "fast", (d.get("fast_mode") or {}).get("value"))
PY

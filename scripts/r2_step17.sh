#!/bin/bash
# ncu of the all-pairs launches (duration + tensor pipe), KDE split timing
mkdir -p gpurun_out
timeout 300 python scripts/kde_time.py > gpurun_out/kde_time.txt 2>&1; cat gpurun_out/kde_time.txt | tail -6
timeout 600 ncu --set full --import-source on --clock-control none -k regex:"gemm_tc" --launch-skip 9 --launch-count 4 -o gpurun_out/ncu_allpairs -f python scripts/allpairs_bench.py > gpurun_out/ncu_allpairs.log 2>&1
python scripts/ncu_summary.py gpurun_out/ncu_allpairs.ncu-rep > gpurun_out/ncu_allpairs.txt; grep -E "Kernel Name|Grid Size|time_duration|tensor_cycles_active.avg.pct_of_peak_sustained_elapsed|registers|warps_active" gpurun_out/ncu_allpairs.txt

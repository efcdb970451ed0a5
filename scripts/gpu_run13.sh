#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "gp_solve" 2>&1 | grep -E "Error|passed|failed|^FAILED" | head
timeout 120 python scripts/gp_profile.py 0
timeout 120 python scripts/gp_profile.py 2
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/gp2_launches.csv python scripts/gp_profile.py 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/gp2_launches.csv 2>/dev/null | head -12

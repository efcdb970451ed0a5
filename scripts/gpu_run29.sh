#!/bin/bash
mkdir -p gpurun_out
for a in 0 2; do timeout 120 python scripts/gp_profile.py $a 2>&1 | tail -1; done
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/gp_launches.csv python scripts/gp_profile.py 2 > /dev/null 2>&1
python scripts/launch_table.py gpurun_out/gp_launches.csv grid 2>/dev/null | head -30

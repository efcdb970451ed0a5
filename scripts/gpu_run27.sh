#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1_fp16.csv python scripts/profile_one_pass.py fp16 > gpurun_out/prof_pass.log 2>&1; tail -n 1 gpurun_out/prof_pass.log
python scripts/launch_table.py gpurun_out/launches_r1_fp16.csv 2>/dev/null | head -34

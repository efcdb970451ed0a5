#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"refiner_prologue_kernel" --launch-count 6 -o gpurun_out/prol -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_prol.log 2>&1
tail -2 gpurun_out/ncu_prol.log

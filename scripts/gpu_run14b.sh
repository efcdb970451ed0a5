#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:chol_block128 --launch-skip 3 --launch-count 1 -o gpurun_out/chol128 -f python scripts/gp_profile.py 2 > gpurun_out/ncu_chol.log 2>&1
tail -3 gpurun_out/ncu_chol.log

#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:"dwconv5x5_relu_h2|refiner_block_c144|refiner_block_small" --launch-skip 61 --launch-count 12 -o gpurun_out/refk -f python scripts/profile_one_pass.py fp16 > gpurun_out/ncu_refk.log 2>&1
tail -3 gpurun_out/ncu_refk.log

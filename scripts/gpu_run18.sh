#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "gp_solve" 2>&1 | grep -E "Error|passed|failed|^FAILED" | head
timeout 60 python scripts/gp_clk.py
timeout 60 python scripts/gp_profile.py 2

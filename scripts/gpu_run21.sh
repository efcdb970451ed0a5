#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --tb=short -p no:cacheprovider -k "kde" 2>&1 | tail -5
timeout 300 python scripts/sample_profile.py 2>&1 | grep -E "kde|multinomial 40000"

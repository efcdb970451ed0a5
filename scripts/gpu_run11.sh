#!/bin/bash
mkdir -p gpurun_out
run() {  # name, env...
  name=$1; shift
  for i in 1 2; do env "$@" timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider -k "coskernel" 2>&1 | grep -E "AssertionError: max|passed|failed" | tr '\n' ' '; done; echo
  env "$@" timeout 600 python bench.py --precision fp16 --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
  python -c "
import json; d=json.load(open('gpurun_out/bench_$name.json')); print('$name', d['value'], d['ms_per_step'], d['e2e']['value'], {k:v for k,v in d['stage_ms_per_step'].items() if k in ('dinov2','gp+decoder','  gp.solve')})"
}
run nopdl ROMAB200_NO_PDL=1
run gemmonly ROMAB200_NO_PDL=2
run trig X=1
run notrig ROMAB200_LIB=$PWD/roma_b200/lib/libromab200_notrig.so
run trig2 X=1

#!/bin/bash
# 8-GPU box: BASELINE configs[2] (64 pairs over 8 GPUs) and configs[3] (roma_indoor, 32 pairs over 4 GPUs) with the NCCL scatter / gather in
# the timed region (the default weak-scaling lines at N = 1, 2, 4, 8 are the driver's SCALE run)
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() { # name nproc port args...
  name=$1; np=$2; port=$3; shift 3
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $np --master-addr 127.0.0.1 --master-port $port bench.py --gpus $np "$@" > gpurun_out/$name.json 2> gpurun_out/$name.err
  tail -n 2 gpurun_out/$name.err
}
run bench_cfg2_n8_g64 8 29521 --steps 4 --warmup 3 --global-pairs 64 --pairs-per-gpu 8
run bench_cfg3_indoor_n4_g32 4 29522 --steps 4 --warmup 3 --global-pairs 32 --pairs-per-gpu 8 --model indoor
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q -p no:cacheprovider --timeout 200 -k "prologue" > gpurun_out/pytest_prologue_n8box.log 2>&1; tail -n 2 gpurun_out/pytest_prologue_n8box.log
python - <<'PY'
import json
for n in ("bench_cfg2_n8_g64", "bench_cfg3_indoor_n4_g32"):
    f = f"gpurun_out/{n}.json"
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(n, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["value"], 2), "scaling", d["scaling"], "|", d["config"]["workload"],
              d["config"].get("nccl_bytes_per_step"), "parity", (d.get("parity") or {}).get("certainty"))
    except Exception as e:
        print(f, "parse failed", e); print(open(f).read()[-400:]); print(open(f.replace(".json", ".err")).read()[-800:])
PY

#!/bin/bash
# 2-GPU validation: NCCL sharding test, default N=2 bench (scatter/gather inside the timed region), a strong-scaling batch
mkdir -p gpurun_out
nvidia-smi -L | head -4
timeout 900 python -m pytest tests/test_sharding_gpu.py -m gpu -q -p no:cacheprovider --timeout 800 > gpurun_out/pytest_sharding_n2.log 2>&1; tail -n 5 gpurun_out/pytest_sharding_n2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_split_n2.json 2> gpurun_out/bench_split_n2.err; tail -n 3 gpurun_out/bench_split_n2.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 3 --global-pairs 16 --pairs-per-gpu 8 > gpurun_out/bench_split_n2_g16.json 2> gpurun_out/bench_split_n2_g16.err; tail -n 3 gpurun_out/bench_split_n2_g16.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_split_n2.json", "gpurun_out/bench_split_n2_g16.json"):
    try:
        d = json.load(open(f))
        print(f, "value", round(d["value"], 2), "ms", round(d["ms_per_step"], 2), "e2e", d["e2e"], "scaling", d["scaling"], d["config"]["workload"], d["config"].get("nccl_bytes_per_step"), "parity", d.get("parity", {}).get("certainty"))
    except Exception as e:
        print(f, "parse failed", e); print(open(f).read()[-500:])
PY

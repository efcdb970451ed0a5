"""How many host threads does the CPU oracle want on this box? (coarse 560 pass only, bounded)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle.roma_oracle import RomaOracle
from roma_b200 import synthetic
mw, dw = synthetic.make_weights(0)
orc = RomaOracle(mw, dw, 560, 864, upsample_preds=False)
A, B, _, _ = synthetic.make_pair(1, 560, None, 1)
print("cpu_count", os.cpu_count())
for th in (16, 32, 64, 128):
    if th > (os.cpu_count() or 1): break
    torch.set_num_threads(th)
    t0 = time.perf_counter(); orc.match(A, B); print(th, "threads: %.1f s" % (time.perf_counter() - t0), flush=True)

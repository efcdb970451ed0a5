"""One eager match() (fp16 mode, 560->864, 1 pair) inside a cudaProfilerStart/Stop range, for
`ncu --profile-from-start off --metrics gpu__time_duration.sum ...` launch lists."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200 import roma_outdoor, synthetic

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
amp = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[prec]
mw, dw = synthetic.make_weights(0)
model = roma_outdoor("cuda", weights=mw, dinov2_weights=dw, amp_dtype=amp)
model.use_cuda_graph = False
A, B, Ah, Bh = [t.cuda() for t in synthetic.make_pair(1, 560, 864, 1)]
for _ in range(2):
    model.match(A, B, im_A_high_res=Ah, im_B_high_res=Bh)
torch.cuda.synchronize()
torch.cuda.profiler.start()
warp, cert = model.match(A, B, im_A_high_res=Ah, im_B_high_res=Bh)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done", warp.shape)

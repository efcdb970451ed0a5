import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200.cabi import call
dev = "cuda"
n, nrhs, batch = 1600, 512, 2
g = torch.Generator().manual_seed(0)
feats = torch.randn(batch, n, 48, generator=g); feats = feats / feats.norm(dim=-1, keepdim=True)
Kyy = ((feats @ feats.transpose(1, 2) - 1) / 0.2).exp() + 0.1 * torch.eye(n)
W0 = torch.zeros(batch, n + nrhs, n); W0[:, :n] = Kyy; W0[:, n:] = torch.randn(nrhs, n, generator=g)
W0 = W0.to(dev); W = W0.clone()
ws_floats = max(batch * ((n + 31) // 32) * 1024 + 1, batch * ((n + 127) // 128) * 16384)
ws = torch.empty(ws_floats, device=dev)
algo = int(sys.argv[1]) if len(sys.argv) > 1 else 2
def gp():
    W.copy_(W0)
    call("romab200_gp_solve", "rb_gp_solve_args", W=W, n=n, nrhs=nrhs, batch=batch, ldw=n, stride=(n + nrhs) * n,
         workspace=ws if algo else None, workspace_bytes=ws_floats * 4 if algo else 0, algo=algo)
for _ in range(2): gp()
torch.cuda.synchronize()
ts = []
for _ in range(5):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); s.record(); gp(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
print("algo", algo, "gp solve median %.3f ms" % sorted(ts)[2])
torch.cuda.profiler.start(); gp(); torch.cuda.synchronize(); torch.cuda.profiler.stop()

# In-kernel phase clocks of chol_block128_kernel.  Needs a debug build of the library with -DRB_CB_CLK linked as
# gpurun_clk.so at the repo root (all objects of roma_b200/lib/obj, gp.cu recompiled with the define).
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import roma_b200.cabi as cabi
lib = cabi.load_library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_clk.so"))
from roma_b200.cabi import call
dev = "cuda"
n, nrhs, batch = 1600, 512, 2
g = torch.Generator().manual_seed(0)
feats = torch.randn(batch, n, 48, generator=g); feats = feats / feats.norm(dim=-1, keepdim=True)
Kyy = ((feats @ feats.transpose(1, 2) - 1) / 0.2).exp() + 0.1 * torch.eye(n)
W0 = torch.zeros(batch, n + nrhs, n); W0[:, :n] = Kyy; W0[:, n:] = torch.randn(nrhs, n, generator=g)
W0 = W0.to(dev); W = W0.clone()
ws_floats = batch * ((n + 127) // 128) * 16384
ws = torch.empty(ws_floats, device=dev)
for _ in range(3):
    W.copy_(W0)
    call("romab200_gp_solve", "rb_gp_solve_args", W=W, n=n, nrhs=nrhs, batch=batch, ldw=n, stride=(n + nrhs) * n,
         workspace=ws, workspace_bytes=ws_floats * 4, algo=2)
torch.cuda.synchronize()
out = (ctypes.c_longlong * 32)()
lib.romab200_debug_clk(out)
c = list(out)
names = {0: "start", 1: "loaded", 2: "panel0", 3: "tiles0", 4: "panel1", 5: "tiles1", 6: "panel2", 7: "tiles2", 8: "panel3", 9: "tiles3", 10: "inv diag", 11: "inv offdiag", 12: "stored"}
prev = c[0]
for i in range(1, 13):
    print("%-12s %7d cycles" % (names[i], c[i] - prev)); prev = c[i]
print("total", c[12] - c[0])

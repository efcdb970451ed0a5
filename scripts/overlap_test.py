"""Does the persistent GP solve co-run with tcgen05 GEMM work on another stream?"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200 import cabi
from roma_b200.cabi import call

dev = "cuda"
n, nrhs, batch = 1600, 512, 2
g = torch.Generator().manual_seed(0)
feats = torch.randn(batch, n, 48, generator=g); feats = feats / feats.norm(dim=-1, keepdim=True)
Kyy = ((feats @ feats.transpose(1, 2) - 1) / 0.2).exp() + 0.1 * torch.eye(n)
W0 = torch.zeros(batch, n + nrhs, n); W0[:, :n] = Kyy; W0[:, n:] = torch.randn(nrhs, n, generator=g)
W0 = W0.to(dev); W = W0.clone()
ws_floats = batch * ((n + 31) // 32) * 1024 + 1
ws = torch.empty(ws_floats, device=dev)

def gp(persistent=True):
    W.copy_(W0)
    call("romab200_gp_solve", "rb_gp_solve_args", W=W, n=n, nrhs=nrhs, batch=batch, ldw=n, stride=(n + nrhs) * n,
         workspace=ws if persistent else None, workspace_bytes=ws_floats * 4 if persistent else 0, algo=1 if persistent else 0)

M, N, K = 95048, 256, 2304
A = torch.randn(M, K, device=dev).half(); B = torch.randn(N, K, device=dev).half(); C = torch.empty(M, N, device=dev, dtype=torch.float16)
def gemms(reps=8):
    for _ in range(reps):
        call("romab200_gemm", "rb_gemm_args", A=A, B=B, C=C, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, dtype_ab=1, dtype_c=1, batch0=1, batch1=1, ntaps=1, alpha=1.0)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); s.record(); fn(); e.record(); torch.cuda.synchronize(); ts.append(s.elapsed_time(e))
    return sorted(ts)[len(ts) // 2]

side = torch.cuda.Stream()
def both(persistent=True):
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    with torch.cuda.stream(side):
        gemms()
    gp(persistent)
    main.wait_stream(side)

print("gp persistent alone  %.3f ms" % timeit(lambda: gp(True)))
print("gp multi-kernel alone %.3f ms" % timeit(lambda: gp(False)))
print("8 gemms alone        %.3f ms" % timeit(gemms))
print("both (persistent)    %.3f ms" % timeit(lambda: both(True)))
print("both (multi-kernel)  %.3f ms" % timeit(lambda: both(False)))

"""A few launches of the ViT fc1 / qkv GEMM shapes for `ncu --set full` (scripts/r2_ncu_gemm.sh): argv[1] = split|fp16."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200 import cabi
from roma_b200.cabi import call

dev = "cuda"
split = (sys.argv[1] if len(sys.argv) > 1 else "split") == "split"


def linear(M, N, K, act=0):
    mk = lambda r, c: torch.randn(r, c, device=dev).to(torch.float16)
    A, Al, B, Bl = mk(M, K), mk(M, K), mk(N, K), mk(N, K)
    bias = torch.randn(N, device=dev)
    args = dict(A=A, B=B, M=M, N=N, K=K, lda=K, ldb=K, ldc=N, ntaps=1, dtype_ab=cabi.RB_F16S if split else cabi.RB_F16, batch0=1, batch1=1, alpha=1.0, bias=bias, act=act,
                C=torch.empty(M, N, device=dev, dtype=torch.float16))
    if split:
        args.update(A_lo=Al, B_lo=Bl, C_lo=torch.empty(M, N, device=dev, dtype=torch.float16), dtype_c=cabi.RB_F16S)
    else:
        args.update(dtype_c=cabi.RB_F16)
    return args


fc1 = linear(3202, 4096, 1024, cabi.ACT_GELU)
qkv = linear(3202, 3072, 1024)
for _ in range(3):
    call("romab200_gemm", "rb_gemm_args", **fc1)
    call("romab200_gemm", "rb_gemm_args", **qkv)
torch.cuda.synchronize()

"""All-pairs CosKernel launches alone, replayed from a CUDA graph (no host launch latency in the timed region):
(a) the engine's three launches per pair (K_AA|K_BB batched, K_AB, K_BA), (b) one launch of all four matrices (batch 2x2)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200 import arch, cabi
from roma_b200.cabi import call

dev = torch.device("cuda:0")
torch.cuda.set_device(dev)
n, cf, E = 1600, 512, 2
ldw = (n + 7) // 8 * 8
g = torch.Generator().manual_seed(0)
x = torch.randn(E * n, cf, generator=g).to(dev)
norms = torch.empty(E * n, device=dev)
call("romab200_row_norms", "rb_rownorm_args", x=x, out=norms, rows=E * n, cols=cf, ldx=cf, dtype=cabi.RB_F32)
hi, lo = torch.empty(E * n, cf, dtype=torch.float16, device=dev), torch.empty(E * n, cf, dtype=torch.float16, device=dev)
call("romab200_split_f16s", "rb_split_pair_args", x=x, hi=hi, lo=lo, rows=E * n, cols=cf, ldx=cf, ldd=cf, row_norm=norms)
stride_w = (n + arch.GP_DIM) * ldw
Wk = torch.zeros(E, n + arch.GP_DIM, ldw, device=dev)
kxy_hi, kxy_lo = torch.zeros(E, n, ldw, dtype=torch.float16, device=dev), torch.zeros(E, n, ldw, dtype=torch.float16, device=dev)
K4 = torch.zeros(2, 2, n, ldw, device=dev)
common = dict(M=n, N=n, K=cf, lda=cf, ldb=cf, ldc=ldw, dtype_ab=cabi.RB_F16S, ntaps=1, alpha=1.0, epi=cabi.EPI_COSKERNEL,
              eps=arch.GP_COS_EPS, inv_t=1.0 / arch.GP_TEMPERATURE, cos_normalized=1)


def three():
    call("romab200_gemm", "rb_gemm_args", A=hi, A_lo=lo, B=hi, B_lo=lo, C=Wk, dtype_c=cabi.RB_F32, batch0=E, batch1=1, sa0=n * cf, sb0=n * cf, sc0=stride_w,
         norm_a=norms, norm_b=norms, sna0=n, snb0=n, diag_add=arch.GP_SIGMA_NOISE, **common)
    for i0, y0 in ((0, 1), (1, 0)):
        call("romab200_gemm", "rb_gemm_args", A=hi[i0 * n:], A_lo=lo[i0 * n:], B=hi[y0 * n:], B_lo=lo[y0 * n:], C=kxy_hi[i0], C_lo=kxy_lo[i0], dtype_c=cabi.RB_F16S,
             batch0=1, batch1=1, sa0=n * cf, sb0=n * cf, sc0=n * ldw, norm_a=norms[i0 * n:], norm_b=norms[y0 * n:], sna0=n, snb0=n, diag_add=0.0, **common)


def four():
    call("romab200_gemm", "rb_gemm_args", A=hi, A_lo=lo, B=hi, B_lo=lo, C=K4, dtype_c=cabi.RB_F32, batch0=2, batch1=2, sa0=n * cf, sa1=0, sb0=0, sb1=n * cf,
         sc0=2 * n * ldw, sc1=n * ldw, norm_a=norms, norm_b=norms, sna0=n, snb0=0, diag_add=0.0, **common)      # timing only: the ABI has no batch1 stride for the norms


def timed(fn, rep=20):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(rep):
            fn()
    ts = []
    for _ in range(7):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); gr.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / rep)
    return sorted(ts)[len(ts) // 2]


out = {}
for name, fn in (("three_launches", three), ("one_launch_2x2", four)):
    try:
        ms = timed(fn)
        out[name] = {"ms_per_pair": ms, "tflops_algorithmic_4_matrices": 4 * 2 * n * n * cf / ms / 1e9}
    except Exception as exc:
        out[name] = {"error": f"{type(exc).__name__}: {exc}"[:300]}
print(json.dumps(out))

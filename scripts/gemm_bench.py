"""Micro-benchmark of romab200_gemm on the shapes of the 560->864 path (CUDA events, L2 flushed between reps)."""
import math
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200 import cabi
from roma_b200.cabi import call

dev = "cuda"
dt = torch.float16
CODE = cabi.DTYPE_CODE
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, fn, flops, reps=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:44s} {t*1e3:9.1f} us  {flops/t/1e9:8.1f} TFLOP/s")


def linear(M, N, K, residual=False, **kw):
    ldk = (K + 7) // 8 * 8
    ldn = (N + 7) // 8 * 8
    A = torch.randn(M, ldk, device=dev).to(dt); B = torch.randn(N, ldk, device=dev).to(dt)
    bias = torch.randn(N, device=dev)
    if residual:
        C = torch.randn(M, ldn, device=dev)
        gamma = torch.rand(N, device=dev)
        args = dict(A=A, B=B, C=C, M=M, N=N, K=K, lda=ldk, ldb=ldk, ldc=ldn, dtype_ab=CODE[dt], dtype_c=cabi.RB_F32, batch0=1, batch1=1,
                    ntaps=1, alpha=1.0, bias=bias, col_scale=gamma, R=C, ldr=ldn, dtype_r=cabi.RB_F32)
    else:
        C = torch.empty(M, ldn, device=dev, dtype=dt)
        args = dict(A=A, B=B, C=C, M=M, N=N, K=K, lda=ldk, ldb=ldk, ldc=ldn, dtype_ab=CODE[dt], dtype_c=CODE[dt], batch0=1, batch1=1, ntaps=1,
                    alpha=1.0, bias=bias)
    args.update(kw)
    return lambda: call("romab200_gemm", "rb_gemm_args", **args)


def attn_qk(Bn, H, N, d):
    dim = H * d
    qkv = torch.randn(Bn, N, 3 * dim, device=dev).to(dt)
    npad = (N + 7) // 8 * 8
    S = torch.empty(Bn, H, N, npad, device=dev, dtype=dt)
    return lambda: call("romab200_gemm", "rb_gemm_args", A=qkv.data_ptr(), B=qkv.data_ptr() + dim * 2, C=S, M=N, N=N, K=d, lda=3 * dim,
                        ldb=3 * dim, ldc=npad, dtype_ab=CODE[dt], dtype_c=CODE[dt], batch0=Bn, batch1=H, ntaps=1, alpha=1 / math.sqrt(d),
                        sa0=N * 3 * dim, sa1=d, sb0=N * 3 * dim, sb1=d, sc0=H * N * npad, sc1=N * npad), (qkv, S)


def conv(E, H, W, cin, cout):
    xp = torch.randn(E, H + 2, W + 2, cin, device=dev).to(dt)
    wm = torch.randn(cout, 9 * cin, device=dev).to(dt)
    out = torch.zeros(E, H + 2, W + 2, cout, device=dev, dtype=dt)
    b = torch.randn(cout, device=dev)
    rows = E * (H + 2) * (W + 2)
    taps = [(ky - 1) * (W + 2) + (kx - 1) for ky in range(3) for kx in range(3)]
    return lambda: call("romab200_gemm", "rb_gemm_args", A=xp, B=wm, C=out, M=rows, N=cout, K=9 * cin, lda=cin, ldb=9 * cin, ldc=cout,
                        dtype_ab=CODE[dt], dtype_c=CODE[dt], batch0=1, batch1=1, ntaps=9, tap_rows=taps, a_rows=rows, alpha=1.0, bias=b,
                        act=cabi.ACT_RELU, rowmap=cabi.ROWMAP_PAD_KEEP, pad_h=H + 2, pad_w=W + 2), (xp, wm, out, b)


which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "fc1"):
    bench("vit fc1  3202x4096x1024 gelu", linear(3202, 4096, 1024, act=cabi.ACT_GELU), 2 * 3202 * 4096 * 1024)
if which in ("all", "qkv"):
    bench("vit qkv  3202x3072x1024", linear(3202, 3072, 1024), 2 * 3202 * 3072 * 1024)
    bench("vit fc2  3202x1024x4096", linear(3202, 1024, 4096), 2 * 3202 * 1024 * 4096)
    bench("vit proj 3202x1024x1024", linear(3202, 1024, 1024), 2 * 3202 * 1024 * 1024)
    bench("big      8192x8192x8192", linear(8192, 8192, 8192), 2 * 8192 ** 3)
    bench("ref8 pw  23328x1137x1137(pad1144)", linear(23328, 1137, 1144), 2 * 23328 * 1137 * 1137)
    bench("ref1 pw  1492992x24x24", linear(1492992, 24, 24), 2 * 1492992 * 24 * 24)
if which in ("all", "ref"):
    bench("ref2 pw  373248x144x144", linear(373248, 144, 144), 2 * 373248 * 144 * 144)
    bench("ref4 pw  93312x569x569", linear(93312, 569, 569), 2 * 93312 * 569 * 569)
    bench("ref8 pw  23328x1137x1137", linear(23328, 1137, 1137), 2 * 23328 * 1137 * 1137)
    bench("vit proj+res 3202x1024x1024 f32 inplace", linear(3202, 1024, 1024, residual=True), 2 * 3202 * 1024 * 1024)
if which in ("ref2",):
    bench("ref2 pw  373248x144x144", linear(373248, 144, 144), 2 * 373248 * 144 * 144)
if which in ("all", "attn"):
    f, keep = attn_qk(2, 16, 1601, 64)
    bench("vit QK^T 2x16x1601x1601x64", f, 2 * 32 * 1601 * 1601 * 64)
if which in ("all", "conv"):
    f, keep2 = conv(2, 864, 864, 64, 64)
    bench("vgg conv 864^2 64->64", f, 2 * 2 * 864 * 864 * 64 * 576)
    f, keep3 = conv(2, 216, 216, 256, 256)
    bench("vgg conv 216^2 256->256", f, 2 * 2 * 216 * 216 * 256 * 2304)

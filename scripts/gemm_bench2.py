"""Micro-benchmark of romab200_gemm (tcgen05 back-end) on the GEMM shapes of the 560->864 path: fp16 operands and RB_F16S
(split-fp16 parity mode) operands, CUDA events, L2 flushed between reps.  Run twice to compare tile schedules:
    ROMAB200_GEMM_PAIR=0 python scripts/gemm_bench2.py      (single-CTA 128 x BN tiles)
    ROMAB200_GEMM_PAIR=1 python scripts/gemm_bench2.py      (CTA-pair 256 x BN tiles where profitable)
TFLOP/s are ALGORITHMIC (2MNK); the split mode executes three MMAs per algorithmic MMA."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200 import cabi
from roma_b200.cabi import call

dev = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, fn, flops, reps=7):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record(); fn(); e.record(); torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    t = sorted(ts)[len(ts) // 2]
    print(f"{name:52s} {t*1e3:9.1f} us  {flops/t/1e9:8.1f} TFLOP/s", flush=True)


def linear(M, N, K, split, out="f32", taps=None, **kw):
    ldk, ldn = (K + 7) // 8 * 8, (N + 7) // 8 * 8
    mk = lambda r, c: torch.randn(r, c, device=dev).to(torch.float16)
    if taps:
        E, H, W, cin = taps
        rows = E * (H + 2) * (W + 2)
        A, Al, B, Bl = mk(rows, cin), mk(rows, cin), mk(N, 9 * cin), mk(N, 9 * cin)
        geo = dict(M=rows, N=N, K=9 * cin, lda=cin, ldb=9 * cin, ntaps=9, a_rows=rows, tap_rows=[(ky - 1) * (W + 2) + (kx - 1) for ky in range(3) for kx in range(3)],
                   rowmap=cabi.ROWMAP_PAD_KEEP, pad_h=H + 2, pad_w=W + 2, act=cabi.ACT_RELU)
        M = rows
    else:
        A, Al, B, Bl = mk(M, ldk), mk(M, ldk), mk(N, ldk), mk(N, ldk)
        geo = dict(M=M, N=N, K=K, lda=ldk, ldb=ldk, ntaps=1)
    bias = torch.randn(N, device=dev)
    args = dict(A=A, B=B, ldc=ldn, dtype_ab=cabi.RB_F16S if split else cabi.RB_F16, batch0=1, batch1=1, alpha=1.0, bias=bias, **geo)
    if split:
        args.update(A_lo=Al, B_lo=Bl)
    if out == "f32":
        args.update(C=torch.empty(M, ldn, device=dev), dtype_c=cabi.RB_F32)
    elif out == "pair":
        args.update(C=torch.empty(M, ldn, device=dev, dtype=torch.float16), C_lo=torch.empty(M, ldn, device=dev, dtype=torch.float16), dtype_c=cabi.RB_F16S)
    else:
        args.update(C=torch.empty(M, ldn, device=dev, dtype=torch.float16), dtype_c=cabi.RB_F16)
    args.update(kw)
    keep.append(args)
    return lambda: call("romab200_gemm", "rb_gemm_args", **args)


keep = []
print("ROMAB200_GEMM_PAIR =", os.environ.get("ROMAB200_GEMM_PAIR", "(default 1)"))
for split in (False, True):
    tag = "split" if split else "fp16 "
    o = "pair" if split else "f16"
    bench(f"{tag} big      8192x8192x8192", linear(8192, 8192, 8192, split, o), 2 * 8192 ** 3, reps=3)
    bench(f"{tag} vit fc1  3202x4096x1024 gelu", linear(3202, 4096, 1024, split, o, act=cabi.ACT_GELU), 2 * 3202 * 4096 * 1024)
    bench(f"{tag} vit qkv  3202x3072x1024", linear(3202, 3072, 1024, split, o), 2 * 3202 * 3072 * 1024)
    bench(f"{tag} vit fc2  3202x1024x4096 ->f32", linear(3202, 1024, 4096, split), 2 * 3202 * 1024 * 4096)
    bench(f"{tag} vit proj 3202x1024x1024 ->f32", linear(3202, 1024, 1024, split), 2 * 3202 * 1024 * 1024)
    bench(f"{tag} fc1 x8   25616x4096x1024 gelu", linear(25616, 4096, 1024, split, o, act=cabi.ACT_GELU), 2 * 25616 * 4096 * 1024)
    bench(f"{tag} ref16 pw 3200x1377x1377", linear(3200, 1377, 1377, split, "f32" if split else "f16"), 2 * 3200 * 1377 * 1377)
    bench(f"{tag} ref8 pw  23328x1137x1137", linear(23328, 1137, 1137, split, "f32" if split else "f16"), 2 * 23328 * 1137 * 1137)
    bench(f"{tag} ref4 pw  93312x569x569", linear(93312, 569, 569, split, "f32" if split else "f16"), 2 * 93312 * 569 * 569)
    bench(f"{tag} ref2 pw  373248x144x144", linear(373248, 144, 144, split, "f32" if split else "f16"), 2 * 373248 * 144 * 144)
    bench(f"{tag} vgg conv 864^2 64->64", linear(0, 64, 0, split, o, taps=(2, 864, 864, 64)), 2 * 2 * 864 * 864 * 64 * 576)
    bench(f"{tag} vgg conv 216^2 256->256", linear(0, 256, 0, split, o, taps=(2, 216, 216, 256)), 2 * 2 * 216 * 216 * 256 * 2304)
    bench(f"{tag} vgg conv 108^2 512->512", linear(0, 512, 0, split, o, taps=(2, 108, 108, 512)), 2 * 2 * 108 * 108 * 512 * 4608)
    keep.clear()
    torch.cuda.empty_cache()

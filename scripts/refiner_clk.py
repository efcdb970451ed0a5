# In-kernel phase clocks of the fused C=144 block and the TMA depthwise kernel.  Needs a debug build of the library with
# -DRB_FZ_CLK linked as gpurun_clk.so at the repo root.
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import roma_b200.cabi as cabi
lib = cabi.load_library(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_clk.so"))
from roma_b200.cabi import call
dev = "cuda"
dt = torch.float16
def ev(fn, n=5):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(n): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / n * 1e3
# fused c144, 2 x 432 x 432
B, H, W, C = 2, 432, 432, 144
x = torch.randn(B, H, W, C, device=dev).to(dt); out = torch.empty_like(x)
dwt = torch.randn(25, C, device=dev) * 0.3; db = torch.randn(C, device=dev); pw = (torch.randn(C, C, device=dev) * 0.1).to(dt); pb = torch.randn(C, device=dev)
f = lambda: call("romab200_refiner_block_c144", "rb_refiner_block_c144_args", **{"in": x}, out=out, ld=C, dw_weight=dwt, ldw=C, dw_bias=db, pw_weight=pw, ld_pw=C, pw_bias=pb, batch=B, h=H, w=W, c=C, dtype=cabi.DTYPE_CODE[dt])
print("c144 2x432x432: %.1f us" % ev(f))
buf = (ctypes.c_longlong * 64)()
lib.romab200_debug_fzclk(buf, 1); f(); torch.cuda.synchronize(); lib.romab200_debug_fzclk(buf, 0)
for w in range(9):
    g = list(buf)[w * 5:w * 5 + 5]
    n = max(g[4], 1)
    print("  dw warp %d: tiles %d  wait_in %6d  compute %6d  wait_a %6d  store %6d cycles/tile" % (w + 5, g[4], g[0] // n, g[1] // n, g[2] // n, g[3] // n))
# dwconv 2 x 216 x 216 x 569
B, H, W, C = 2, 216, 216, 569
ld = 576
x = torch.zeros(B, H, W, ld, device=dev, dtype=dt); x[..., :C] = torch.randn(B, H, W, C, device=dev).to(dt); out = torch.empty_like(x)
wt = torch.zeros(25, ld, device=dev); wt[:, :C] = torch.randn(25, C, device=dev) * 0.3; b = torch.randn(ld, device=dev)
f2 = lambda: call("romab200_dwconv5x5_relu", "rb_dwconv_args", **{"in": x}, out=out, ldi=ld, ldo=ld, weight=wt, ldw=ld, bias=b, batch=B, h=H, w=W, c=C, dtype=cabi.DTYPE_CODE[dt])
print("dwconv 2x216x216x569: %.1f us" % ev(f2))
lib.romab200_debug_dwclk(buf, 1); f2(); torch.cuda.synchronize(); lib.romab200_debug_dwclk(buf, 0)
g = list(buf)[:4]; n = max(g[3], 1)
print("  warp 0: tiles %d  wait_full %6d  compute %6d  store %6d cycles/tile" % (g[3], g[0] // n, g[1] // n, g[2] // n))
for (B, H, W, C) in [(2, 108, 108, 1137), (2, 70, 70, 1137), (2, 140, 140, 569), (2, 40, 40, 1377)]:
    ld = (C + 7) // 8 * 8
    x = torch.zeros(B, H, W, ld, device=dev, dtype=dt); out = torch.empty_like(x)
    wt = torch.zeros(25, ld, device=dev); b = torch.randn(ld, device=dev)
    f3 = lambda: call("romab200_dwconv5x5_relu", "rb_dwconv_args", **{"in": x}, out=out, ldi=ld, ldo=ld, weight=wt, ldw=ld, bias=b, batch=B, h=H, w=W, c=C, dtype=cabi.DTYPE_CODE[dt])
    print("dwconv %dx%dx%dx%d: %.1f us" % (B, H, W, C, ev(f3)))

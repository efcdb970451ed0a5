"""KDE of sample(): 40000 points, one pass vs the j-split version the engine uses (CUDA events, graph replay of 10 calls)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200.cabi import call
dev = "cuda"
n = 40000
x = (torch.rand(n, 4, generator=torch.Generator().manual_seed(0)) * 2 - 1).to(dev)
out = torch.empty(n, device=dev)
for splits in (1, 8, 16, -16):          # negative: the symmetric (upper-triangle) schedule with that many splits
    sym = splits < 0
    splits = abs(splits)
    nws = (splits + (n + 255) // 256) * n if sym else max(splits, 1) * n
    ws = torch.empty(nws, device=dev)
    fn = lambda: call("romab200_kde_density", "rb_kde_args", x=x, density=out, n=n, std=0.1, half=1, workspace=ws if splits > 1 else None, splits=splits,
                      symmetric=int(sym), workspace_floats=nws)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn(); fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(10):
            fn()
    ts = []
    for _ in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); g.replay(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) / 10)
    print(f"splits {splits:2d}{" symmetric" if sym else ""}: {sorted(ts)[2]:.4f} ms, checksum {out.double().sum().item():.3f}")

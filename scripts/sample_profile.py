import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from roma_b200.cabi import call
dev = "cuda"
g = torch.Generator().manual_seed(0)
cert = torch.rand(864, 1728, generator=g).to(dev) * 0.3
warp = torch.rand(864, 1728, 4, generator=g).to(dev) * 2 - 1
def t(fn, name, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    print(f"{name:40s} {sorted(ts)[len(ts)//2]:8.3f} ms")
    return r
def thresh():
    c = cert.clone(); c[c > 0.05] = 1; return c
c = t(thresh, "clone + masked assign")
cf = c.reshape(-1); m = warp.reshape(-1, 4)
good = t(lambda: torch.multinomial(cf, 40000, replacement=False), "multinomial 40000 of 1.49M")
gm = t(lambda: (m[good], cf[good]), "gather")[0]
def kde():
    out = torch.empty(40000, device=dev)
    call("romab200_kde_density", "rb_kde_args", x=gm.contiguous(), density=out, n=40000, std=0.1, half=1)
    return out
d = t(kde, "kde 40000")
def pp():
    dd = d.half(); p = 1 / (dd + 1); p[dd < 10] = 1e-7; return p
p = t(pp, "p = 1/(d+1), mask")
t(lambda: torch.multinomial(p, 10000, replacement=False), "multinomial 10000 of 40000 (half)")
t(lambda: torch.multinomial(p.float(), 10000, replacement=False), "multinomial 10000 of 40000 (float)")
def topk_way():
    q = torch.empty_like(cf).exponential_(1); return torch.topk(cf / q, 40000).indices
t(topk_way, "exponential + topk 40000 of 1.49M")

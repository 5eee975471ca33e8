import sys, time
from pathlib import Path
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import gemlab_b200 as g
from gemlab_b200 import Block, GeoKind, Coef, CommonArgs, Context, integ
ctx = Context(0)
sub = Block.new_cube(1.0).set_ndiv([200, 200, 25]).subdivide(GeoKind.Hex8)
n = sub.ncell
keep = []
for name in ("coords", "kinds", "markers", "conn_off", "conn"):
    tp = torch.from_numpy(getattr(sub, name)).clone().pin_memory(); keep.append(tp); setattr(sub, name, tp.numpy())
host = torch.empty(n * 576, dtype=torch.float64).pin_memory()
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1/3, False))
for packed in (True, False):
    per = 300 if packed else 576
    buf = host[:n*per]
    a = CommonArgs(packed=packed)
    def step():
        t0 = time.perf_counter(); dm = ctx.upload(sub); torch.cuda.synchronize(); t1 = time.perf_counter()
        integ.mat_10_bdb(ctx, dm, None, a, Coef.uniform(dd), out=buf); t2 = time.perf_counter()
        dm.free(); t3 = time.perf_counter(); return t1-t0, t2-t1, t3-t2
    step()
    for _ in range(3):
        u, i, f = step()
        print("packed=%s upload(sync) %.2f ms, integ->host %.2f ms (%.1f GB/s), free %.2f ms -> %.2f M el/s" % (packed, 1e3*u, 1e3*i, n*per*8/i/1e9, 1e3*f, n/(u+i+f)/1e6))
dev = torch.empty(n*300, dtype=torch.float64, device="cuda")
torch.cuda.synchronize(); t0=time.perf_counter()
for _ in range(3): host[:n*300].copy_(dev, non_blocking=True)
torch.cuda.synchronize(); print("plain D2H packed size: %.2f ms each" % (1e3*(time.perf_counter()-t0)/3))

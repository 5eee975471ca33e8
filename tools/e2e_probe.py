import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import gemlab_b200 as g
from gemlab_b200 import Block, GeoKind, Coef, CommonArgs, Context, integ
ctx = Context(0)
sub = Block.new_cube(1.0).set_ndiv([200, 200, 50]).subdivide(GeoKind.Hex8)
n = sub.ncell
host = torch.empty(n * 576, dtype=torch.float64).pin_memory()
dev = torch.empty(n * 576, dtype=torch.float64, device="cuda")
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1/3, False))
def t(f, reps=3):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
print("plain D2H 9.2GB pinned: %.1f ms" % (1e3 * t(lambda: host.copy_(dev, non_blocking=True))))
def step():
    t0 = time.perf_counter()
    dm = ctx.upload(sub)
    t1 = time.perf_counter()
    integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=host)
    t2 = time.perf_counter()
    dm.free()
    t3 = time.perf_counter()
    return t1 - t0, t2 - t1, t3 - t2
def pin(mesh):
    keep = []
    for name in ("coords", "kinds", "markers", "conn_off", "conn"):
        tp = torch.from_numpy(getattr(mesh, name)).clone().pin_memory()
        keep.append(tp)
        setattr(mesh, name, tp.numpy())
    return keep
step()
for _ in range(3):
    a, b, c = step()
    print("pageable: upload %.1f ms, integ->host %.1f ms, free %.1f ms, total %.1f ms -> %.2f M el/s" % (1e3*a, 1e3*b, 1e3*c, 1e3*(a+b+c), n/(a+b+c)/1e6))
keep = pin(sub)
step()
for _ in range(3):
    a, b, c = step()
    print("upload %.1f ms, integ->host %.1f ms, free %.1f ms, total %.1f ms -> %.2f M el/s" % (1e3*a, 1e3*b, 1e3*c, 1e3*(a+b+c), n/(a+b+c)/1e6))

import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import gemlab_b200 as g
from gemlab_b200 import Coef, CommonArgs, Context, GeoKind, integ
ctx = Context(0)
dm = ctx.block(GeoKind.Hex8, (100, 100, 50))
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1/3, False))
n = dm.ncell * 576
pinned = torch.empty(n, dtype=torch.float64).pin_memory()
pageable = np.empty(n)
pageable[:] = 0.0  # touch
for name, out in (("pinned", pinned), ("pageable", pageable)):
    for rep in range(3):
        t0 = time.perf_counter()
        integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out)
        dt = time.perf_counter() - t0
    print(f"{name}: {dt*1e3:.1f} ms for {dm.ncell} Hex8 cells = {dm.ncell/dt/1e6:.1f} M el/s, {n*8/dt/1e9:.1f} GB/s")

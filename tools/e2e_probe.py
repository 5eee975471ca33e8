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
print("plain D2H 9.2GB pinned: %.1f ms -> %.1f GB/s" % (1e3 * t(lambda: host.copy_(dev, non_blocking=True)), n * 576 * 8 / t(lambda: host.copy_(dev, non_blocking=True)) / 1e9))
dm = None
def up():
    global dm
    dm = ctx.upload(sub)
print("upload: %.1f ms" % (1e3 * t(up)))
print("integ -> pinned host: %.1f ms" % (1e3 * t(lambda: integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=host))))
print("integ -> device: %.2f ms" % (1e3 * t(lambda: (integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=dev), ctx.sync()))))
def free_up():
    d = ctx.upload(sub); d.free()
print("upload+free: %.1f ms" % (1e3 * t(free_up)))

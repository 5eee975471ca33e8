"""Times gl_mat_10_bdb over an n^3 Hex8 Block mesh (device-resident), prints ms per pass. usage: time_hex8.py [n] [reps]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import gemlab_b200 as g
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, integ

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
ctx = Context(0)
mesh = Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex8)
dm = ctx.upload(mesh)
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, False))
out = torch.empty(mesh.ncell * 576, dtype=torch.float64, device="cuda")
for _ in range(3):
    integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out)
torch.cuda.synchronize()
best, tot = 1e9, 0.0
for _ in range(reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    best = min(best, ms)
    tot += ms
if len(sys.argv) > 3 and sys.argv[3] == "markers":  # two materials selected by cell marker (PER_MARKER) and one record per cell (PER_CELL)
    import numpy as np
    mesh.markers[mesh.ncell // 2:] = 2
    dm2 = ctx.upload(mesh)
    recs = np.stack([dd, 3.0 * dd])
    for name, coef in (("per_marker", Coef.per_marker([1, 2], recs)),
                       ("per_cell", Coef.per_cell(torch.tensor(np.tile(dd, (mesh.ncell, 1)), device="cuda")))):
        for _ in range(3):
            integ.mat_10_bdb(ctx, dm2, None, CommonArgs(), coef, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            integ.mat_10_bdb(ctx, dm2, None, CommonArgs(), coef, out=out)
        e1.record()
        torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / reps:.4f} ms per pass")
if len(sys.argv) > 3 and sys.argv[3] == "fused":  # stiffness + internal force from one pass vs two calls
    sig = torch.randn(mesh.ncell * 8, 6, dtype=torch.float64, device="cuda")
    out_d = torch.empty(mesh.ncell * 24, dtype=torch.float64, device="cuda")
    def one():
        integ.fused(ctx, dm, None, CommonArgs(), [("mat_10_bdb", Coef.uniform(dd), out), ("vec_04_bt", Coef.per_gauss(sig), out_d)])
    def two():
        integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out)
        integ.vec_04_bt(ctx, dm, None, CommonArgs(), Coef.per_gauss(sig), out=out_d)
    for name, fn in (("fused K + d", one), ("separate K, d", two)):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        print(f"{name}: {e0.elapsed_time(e1) / reps:.4f} ms per pass")
print(f"n={n} cells={mesh.ncell} avg_ms={tot / reps:.4f} best_ms={best:.4f} write_TBs={mesh.ncell * 4608 / (tot / reps * 1e-3) / 1e12:.3f}")

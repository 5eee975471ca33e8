"""Breakdown of BASELINE config 5 (mixed Tri3/Qua8 axisymmetric): per call and per kind.  usage: time_cfg5.py [n]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
import gemlab_b200 as g
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, Mesh, integ

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1200
ctx = Context(0)
q8 = Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.Qua8)
q8.coords[:, 0] += 1.0
c = q8.conn_matrix().astype(np.int64)
tri = np.stack([c[:, [0, 1, 2]], c[:, [0, 2, 3]]], axis=1).reshape(-1, 3)
meshes = {"qua8 only": Mesh.homogeneous(2, q8.coords, GeoKind.Qua8, c), "tri3 only": Mesh.homogeneous(2, q8.coords, GeoKind.Tri3, tri)}
# mixed: alternate rows
kinds, conn, lens = [], [], []
for r in range(n):
    cells = c[r * n:(r + 1) * n]
    if r % 2 == 0:
        kinds.append(np.full(n, int(GeoKind.Qua8), dtype=np.uint8)); conn.append(cells.reshape(-1)); lens.append(np.full(n, 8))
    else:
        t = np.stack([cells[:, [0, 1, 2]], cells[:, [0, 2, 3]]], axis=1).reshape(-1)
        kinds.append(np.full(2 * n, int(GeoKind.Tri3), dtype=np.uint8)); conn.append(t); lens.append(np.full(2 * n, 3))
lens = np.concatenate(lens)
meshes["mixed"] = Mesh(2, q8.coords, np.concatenate(kinds), np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64), np.concatenate(conn).astype(np.uint64))
dd = g.mandel_matrix(g.lin_elasticity(96.0, 1 / 3, True, False))
a = CommonArgs(axisymmetric=True)

def timed(fn, reps=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

for name, mesh in meshes.items():
    dm = ctx.upload(mesh)
    ok = torch.empty(integ.out_total(dm, "mat_10_bdb", None, a), dtype=torch.float64, device="cuda")
    oh = torch.empty(integ.out_total(dm, "mat_03_btb", None, a), dtype=torch.float64, device="cuda")
    ov = torch.empty(integ.out_total(dm, "vec_03_bv", None, a), dtype=torch.float64, device="cuda")
    t1 = timed(lambda: integ.mat_10_bdb(ctx, dm, None, a, Coef.uniform(dd), out=ok))
    t2 = timed(lambda: integ.fused(ctx, dm, None, a, [("mat_03_btb", Coef.uniform([2.0, 2.0, 0.0, 0.0]), oh), ("vec_03_bv", Coef.uniform([1.0, 2.0]), ov)]))
    t3 = timed(lambda: integ.fused(ctx, dm, None, a, [("mat_10_bdb", Coef.uniform(dd), ok), ("mat_03_btb", Coef.uniform([2.0, 2.0, 0.0, 0.0]), oh),
                                                       ("vec_03_bv", Coef.uniform([1.0, 2.0]), ov)]))
    print(f"{name:10s} all three from one gl_integ_fused call: {t3:.3f} ms ({mesh.ncell / t3 / 1e6:.3f} G el/s)  [{ctx.last_kernel()}]")
    print(f"{name:10s} {mesh.ncell:8d} cells: mat_10_bdb {t1:.3f} ms ({ok.numel()*8/t1/1e6:.0f} GB/s out), fused btb+bv {t2:.3f} ms ({(oh.numel()+ov.numel())*8/t2/1e6:.0f} GB/s out)")

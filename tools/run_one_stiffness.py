#!/usr/bin/env python3
"""Runs mat_10_bdb a few times on one kind (uniform modulus, device-resident output): the target of an ncu capture.
usage: python tools/run_one_stiffness.py tet10|qua4|qua8|qua9|hex8|hex20 [n] [axi]"""
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import gemlab_b200 as g  # noqa: E402
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, integ  # noqa: E402

kind = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 50
axi = len(sys.argv) > 3 and sys.argv[3] == "axi"
ctx = Context(0)
if kind == "tet10":
    mesh = g.split_hex_cells_into_tet10(n, n, n, seed=12345)
elif kind.startswith("qua"):
    mesh = Block.new_square(1.0).set_ndiv([20 * n, 20 * n]).subdivide(GeoKind.from_str(kind))
else:
    mesh = Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.from_str(kind))
dm = ctx.upload(mesh)
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, mesh.ndim == 2))
bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), mesh.ndim)
out = torch.empty(mesh.ncell * bs, dtype=torch.float64, device="cuda")
for _ in range(4):
    integ.mat_10_bdb(ctx, dm, None, CommonArgs(axisymmetric=axi), Coef.uniform(dd), out=out)
ctx.sync()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    integ.mat_10_bdb(ctx, dm, None, CommonArgs(axisymmetric=axi), Coef.uniform(dd), out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"{kind}{' axisymmetric' if axi else ''}: {mesh.ncell} cells, {ms:.4f} ms, {mesh.ncell / ms / 1e3:.1f} M el/s, {mesh.ncell * bs * 8 / ms / 1e6:.0f} GB/s out")

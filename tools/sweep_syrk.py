#!/usr/bin/env python3
"""Sweeps the tuning knobs of the tensor-core stiffness kernel (needs the -DGL_TUNING library: GEMLAB_CUDA_TUNING=1).
usage: GEMLAB_CUDA_TUNING=1 python tools/sweep_syrk.py [n]"""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
os.environ.setdefault("GEMLAB_CUDA_TUNING", "1")

import torch  # noqa: E402

import gemlab_b200 as g  # noqa: E402
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, integ  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
ctx = Context(0)
mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex20), 0.05 / n)
dm = ctx.upload(mesh)
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
out = torch.empty(mesh.ncell * 3600, dtype=torch.float64, device="cuda")


def run(label):
    for _ in range(2):
        integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out)
    ctx.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{label:40s} {ctx.last_kernel():28s} {ms:8.4f} ms  {mesh.ncell / ms / 1e3:7.1f} M el/s", flush=True)


for st in ():
    os.environ["GL_SYRK_STAGGER"] = str(st)
    run(f"stagger {st} ns/slot")
os.environ.pop("GL_SYRK_STAGGER")
for grid in (74, 148, 296):
    os.environ["GL_SYRK_GRID"] = str(grid)
    run(f"grid {grid}")
os.environ.pop("GL_SYRK_GRID")
os.environ["GL_DISABLE_SYRK"] = "1"
run("tile kernel (round 1)")

#!/usr/bin/env python3
"""mat_10_bdb on the quadratic 2D kinds (Tri6, Qua8, Qua9), plane and axisymmetric, uniform modulus, device-resident output:
which kernel runs, elements/s and the fraction of the HBM peak by output bytes.  usage: time_2d_kinds.py [n]"""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import gemlab_b200 as g  # noqa: E402
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, Mesh, integ  # noqa: E402
from tools.bench_configs import timed  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
ctx = Context(0)
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, True))
for kind in ("tri6", "qua8", "qua9"):
    base = Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.Qua9 if kind == "tri6" else GeoKind.from_str(kind))
    base.coords[:, 0] += 1.0
    if kind == "tri6":
        c = base.conn_matrix()
        tris = np.concatenate([c[:, [0, 1, 2, 4, 5, 8]], c[:, [0, 2, 3, 8, 6, 7]]])
        mesh = Mesh.homogeneous(2, base.coords, GeoKind.Tri6, tris)
    else:
        mesh = base
    dm = ctx.upload(mesh)
    bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), 2)
    out = torch.empty(mesh.ncell * bs, dtype=torch.float64, device="cuda")
    for axi in (False, True):
        ms = timed(lambda: integ.mat_10_bdb(ctx, dm, None, CommonArgs(axisymmetric=axi), Coef.uniform(dd), out=out))
        print(json.dumps({"kind": kind, "axisymmetric": axi, "kernel": ctx.last_kernel(), "cells": int(mesh.ncell), "ms": ms,
                          "elements_per_s": mesh.ncell / (ms * 1e-3), "out_gbs": mesh.ncell * bs * 8 / (ms * 1e-3) / 1e9}), flush=True)
    dm.free()

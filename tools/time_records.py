#!/usr/bin/env python3
"""Times mat_10_bdb on Qua4 / Qua8 / Hex20 / Tet10 with a uniform modulus, one modulus per marker and one per cell (device-resident
outputs) and checks a sample of cells against the oracle.  GL_DISABLE_TILE=1 GL_DISABLE_SMALL2D=1 shows the register-blocked fallback."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import gemlab_b200 as g  # noqa: E402
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, integ  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from tools.bench_configs import timed  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
    ctx = Context(0)
    rng = np.random.default_rng(5)
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ("qua4", "qua8", "hex20", "tet10")
    for kind in kinds:
        two_dim = kind in ("qua4", "qua8")
        if two_dim:
            mesh = g.jitter_interior_points(Block.new_square(1.0).set_ndiv([20 * n, 20 * n]).subdivide(GeoKind.from_str(kind)), 0.01 / n)
        elif kind == "hex20":
            mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex20), 0.05 / n)
        else:
            mesh = g.split_hex_cells_into_tet10(n, n, n, seed=12345)
        mesh.markers[:] = rng.integers(1, 4, mesh.ncell)
        dm = ctx.upload(mesh)
        recs = np.stack([g.mandel_matrix(g.lin_elasticity(100.0 * (k + 1), 0.1 * (k + 1), two_dim)) for k in range(3)])
        per_cell = np.ascontiguousarray(recs[mesh.markers - 1])
        per_cell_dev = torch.from_numpy(per_cell).cuda()
        ns2 = recs.shape[1]
        bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), mesh.ndim)
        out = torch.empty(mesh.ncell * bs, dtype=torch.float64, device="cuda")
        sel = np.sort(rng.choice(mesh.ncell, size=200, replace=False))
        sub = mesh.permuted(sel)
        cases = {
            "uniform": (Coef.uniform(recs[0]), np.broadcast_to(recs[0], (len(sel), ns2))),
            "per_marker": (Coef.per_marker([1, 2, 3], recs), per_cell[sel]),
            "per_cell_host": (Coef.per_cell(per_cell), per_cell[sel]),
            "per_cell_device": (Coef.per_cell(per_cell_dev), per_cell[sel]),
        }
        for name, (coef, ref_recs) in cases.items():
            ms = timed(lambda: integ.mat_10_bdb(ctx, dm, None, CommonArgs(), coef, out=out), steps=3, warmup=1)
            ctx.sync()
            got = out.cpu().numpy().reshape(mesh.ncell, bs)[sel]
            ref = po.mesh_integ("mat_10_bdb", mesh.ndim, sub.coords, sub.kinds, sub.conn_off, sub.conn,
                                coef=np.ascontiguousarray(ref_recs).reshape(-1), coef_cell_stride=ns2).reshape(len(sel), bs)
            err = float(np.max(np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)))
            print(json.dumps({"kind": kind, "coef": name, "kernel": ctx.last_kernel(), "cells": mesh.ncell, "ms": ms,
                              "elements_per_s": mesh.ncell / (ms * 1e-3), "max_rel_err": err}), flush=True)


if __name__ == "__main__":
    main()

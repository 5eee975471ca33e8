#!/usr/bin/env python3
"""Times mat_10_bdb with one modulus per Gauss point (device-resident records, device output) on Hex8 / Hex20 / Tet10 / Qua8 / Qua4
and checks a sample of cells against the oracle.  usage: time_per_gauss.py [scale] [kinds,comma,separated]"""
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
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    kinds = sys.argv[2].split(",") if len(sys.argv) > 2 else ["hex8", "hex20", "tet10", "qua8", "qua4"]
    ctx = Context(0)
    rng = np.random.default_rng(5)
    for kind in kinds:
        if kind == "hex8":
            n = max(2, int(round(100 * scale)))
            mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex8), 0.05 / n)
        elif kind == "hex20":
            n = max(2, int(round(50 * scale)))
            mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex20), 0.05 / n)
        elif kind == "tet10":
            n = max(2, int(round(50 * scale)))
            mesh = g.split_hex_cells_into_tet10(n, n, n, seed=12345)
        else:
            n = max(2, int(round(1000 * scale)))
            mesh = g.jitter_interior_points(Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.from_str(kind)), 0.05 / n)
        gk = GeoKind.from_str(kind)
        ng = g.Gauss.new(gk).npoint()
        ns = 2 * mesh.ndim
        dm = ctx.upload(mesh)
        bs = integ.op_out_size("mat_10_bdb", gk, mesh.ndim)
        # symmetric positive definite records built on the device: D_p = A A^T + ns I
        a = torch.randn((mesh.ncell * ng, ns, ns), dtype=torch.float64, device="cuda")
        recs = (a @ a.transpose(1, 2) + ns * torch.eye(ns, dtype=torch.float64, device="cuda")).reshape(mesh.ncell * ng, ns * ns).contiguous()
        del a
        out = torch.empty(mesh.ncell * bs, dtype=torch.float64, device="cuda")
        sel = np.sort(rng.choice(mesh.ncell, size=min(200, mesh.ncell), replace=False))
        sub = mesh.permuted(sel)
        for name in ("symmetric", "general"):
            if name == "general":
                recs[recs.shape[0] // 2, 1] += 0.5
            coef = Coef.per_gauss(recs)
            ms = timed(lambda: integ.mat_10_bdb(ctx, dm, None, CommonArgs(), coef, out=out), steps=3, warmup=1)
            ctx.sync()
            fam = ctx.last_kernel()
            got = out.cpu().numpy().reshape(mesh.ncell, bs)[sel]
            rsel = recs.reshape(mesh.ncell, ng * ns * ns)[torch.from_numpy(sel).cuda()].cpu().numpy()
            ref = po.mesh_integ("mat_10_bdb", mesh.ndim, sub.coords, sub.kinds, sub.conn_off, sub.conn, coef=rsel.reshape(-1),
                                coef_cell_stride=ng * ns * ns, coef_gp_stride=ns * ns).reshape(len(sel), bs)
            err = float(np.max(np.abs(got - ref).max(axis=1) / np.abs(ref).max(axis=1)))
            print(json.dumps({"kind": kind, "records": name, "kernel": fam, "cells": mesh.ncell, "ngauss": ng, "ms": ms,
                              "elements_per_s": mesh.ncell / (ms * 1e-3), "max_rel_err": err}), flush=True)
        del recs, out


if __name__ == "__main__":
    main()

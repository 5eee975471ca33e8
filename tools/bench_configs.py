#!/usr/bin/env python3
"""Times the other BASELINE.json configs (1, 3, 4, 5) on one GPU, device-resident, and prints one JSON line per config.

These are parity-test cases, not bench.py lines (bench.py times configs[1]); this script records where each stands
against its roofline (per-element bytes / flops from SURVEY.md 8(d)).  Sizes can be scaled with --scale (1.0 = the
BASELINE sizes; default 0.25 of the cells to keep the run short).  Parity of a sample of cells against the oracle is
checked in the same run.
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import torch  # noqa: E402

import gemlab_b200 as g  # noqa: E402
from gemlab_b200 import Block, Coef, CommonArgs, Context, GeoKind, Mesh, integ  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HBM_GBS = 6553.0
try:
    HBM_GBS = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"])
except Exception:
    pass


def timed(fn, steps=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def parity(mesh, op, got, coef, nsample=2000, **kw):
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(mesh.ncell, size=min(nsample, mesh.ncell), replace=False))
    sub = mesh.permuted(sel)
    ref = po.mesh_integ(op, sub.ndim, sub.coords, sub.kinds, sub.conn_off, sub.conn, coef=coef, **kw)
    sizes = np.array([integ.op_out_size(op, GeoKind(int(k)), mesh.ndim) for k in mesh.kinds], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(sizes)])
    worst, pos = 0.0, 0
    for e in sel:
        n = int(sizes[e])
        r = ref[pos:pos + n]
        worst = max(worst, float(np.max(np.abs(got[off[e]:off[e] + n] - r)) / np.max(np.abs(r))))
        pos += n
    return worst


def report(name, ncell, ms, flops_per, bytes_per, fp64_peak, err, extra=None):
    els = ncell / (ms * 1e-3)
    line = {"config": name, "cells": ncell, "ms_per_pass": ms, "elements_per_s": els,
            "hbm_gbs": bytes_per * els / 1e9, "hbm_frac": bytes_per * els / 1e9 / HBM_GBS,
            "fp64_tflops": flops_per * els / 1e12, "fp64_frac": flops_per * els / 1e12 / fp64_peak,
            "max_rel_err_vs_oracle": err}
    if extra:
        line.update(extra)
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    ap.add_argument("--only", type=int, nargs="*", default=[1, 3, 4, 5])
    args = ap.parse_args()
    ctx = Context(0)
    fp64_peak = ctx.fp64_peak("dfma", 8192)
    s = args.scale

    if 1 in args.only:  # 2D Qua4 1000x1000, plane-strain mat_10_bdb, 2x2 Gauss
        n = max(10, int(round(1000 * s ** 0.5)))
        mesh = g.jitter_interior_points(Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.Qua4), 0.2 / n)
        dm = ctx.upload(mesh)
        dd = g.mandel_matrix(g.lin_elasticity(10_000.0, 0.2, True, False))
        out = torch.empty(mesh.ncell * 64, dtype=torch.float64, device="cuda")
        ms = timed(lambda: integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out))
        ctx.sync()
        err = parity(mesh, "mat_10_bdb", out.cpu().numpy(), dd)
        report("1: qua4 mat_10_bdb plane strain 2x2", mesh.ncell, ms, 1824, 544, fp64_peak, err)

    if 3 in args.only:  # 3D Hex20 100^3, mat_10_bdb, 27-point Gauss
        n = max(4, int(round(100 * s ** (1 / 3))))
        mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex20), 0.05 / n)
        dm = ctx.upload(mesh)
        dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
        out = torch.empty(mesh.ncell * 3600, dtype=torch.float64, device="cuda")
        ms = timed(lambda: integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd), out=out), steps=3, warmup=1)
        ctx.sync()
        err = parity(mesh, "mat_10_bdb", out.cpu().numpy(), dd, nsample=300)
        report("3: hex20 mat_10_bdb 27-point", mesh.ncell, ms, 662310, 28978, fp64_peak, err)

    if 4 in args.only:  # Tet10 5e6 cells, mat_01_nsn + vec_01_ns
        n = max(3, int(round(94 * s ** (1 / 3))))
        mesh = g.split_hex_cells_into_tet10(n, n, n, seed=12345)
        dm = ctx.upload(mesh)
        out_m = torch.empty(mesh.ncell * 100, dtype=torch.float64, device="cuda")
        out_v = torch.empty(mesh.ncell * 10, dtype=torch.float64, device="cuda")

        def step():
            integ.fused(ctx, dm, None, CommonArgs(), [("mat_01_nsn", Coef.uniform([2.7]), out_m),
                                                      ("vec_01_ns", Coef.uniform([9.81 * 2.7]), out_v)])

        ms = timed(step)
        ctx.sync()
        err = max(parity(mesh, "mat_01_nsn", out_m.cpu().numpy(), np.array([2.7])),
                  parity(mesh, "vec_01_ns", out_v.cpu().numpy(), np.array([9.81 * 2.7])))
        report("4: tet10 mat_01_nsn + vec_01_ns 14-point (permuted cells)", mesh.ncell, ms, 12264, 953, fp64_peak, err)

    if 5 in args.only:  # mixed Tri3 / Qua8 stripes, axisymmetric mat_10_bdb + mat_03_btb + vec_03_bv
        n = max(8, int(round((10_000_000 * s / 1.5) ** 0.5)))
        q8 = Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.Qua8)
        q8.coords[:, 0] += 1.0
        c = q8.conn_matrix().astype(np.int64)
        row = (np.arange(q8.ncell) // n)
        is_q = row % 2 == 0
        nq, nt = int(is_q.sum()), int((~is_q).sum())
        # cell order follows the stripes: a Qua8 row, then the two Tri3 of every split cell of the next row
        kinds, conn_chunks, lens = [], [], []
        for r in range(n):
            cells = c[r * n:(r + 1) * n]
            if r % 2 == 0:
                kinds.append(np.full(n, int(GeoKind.Qua8), dtype=np.uint8))
                conn_chunks.append(cells.reshape(-1))
                lens.append(np.full(n, 8))
            else:
                tri = np.stack([cells[:, [0, 1, 2]], cells[:, [0, 2, 3]]], axis=1).reshape(-1)
                kinds.append(np.full(2 * n, int(GeoKind.Tri3), dtype=np.uint8))
                conn_chunks.append(tri)
                lens.append(np.full(2 * n, 3))
        kinds = np.concatenate(kinds)
        lens = np.concatenate(lens)
        conn_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        mesh = Mesh(2, q8.coords, kinds, conn_off, np.concatenate(conn_chunks).astype(np.uint64))
        dm = ctx.upload(mesh)
        dd = g.mandel_matrix(g.lin_elasticity(96.0, 1 / 3, True, False))
        a = CommonArgs(axisymmetric=True)
        tot_k = integ.out_total(dm, "mat_10_bdb", None, a)
        tot_h = integ.out_total(dm, "mat_03_btb", None, a)
        tot_v = integ.out_total(dm, "vec_03_bv", None, a)
        out_k = torch.empty(tot_k, dtype=torch.float64, device="cuda")
        out_h = torch.empty(tot_h, dtype=torch.float64, device="cuda")
        out_v = torch.empty(tot_v, dtype=torch.float64, device="cuda")

        def step():  # one gl_integ_fused call, as bench.py issues it: the Qua8 cells take all three integrals from one kernel
            integ.fused(ctx, dm, None, a, [("mat_10_bdb", Coef.uniform(dd), out_k), ("mat_03_btb", Coef.uniform([2.0, 2.0, 0.0, 0.0]), out_h),
                                           ("vec_03_bv", Coef.uniform([1.0, 2.0]), out_v)])

        ms = timed(step, steps=3, warmup=1)
        ctx.sync()
        errs = {"mat_10_bdb": parity(mesh, "mat_10_bdb", out_k.cpu().numpy(), dd, axisymmetric=True),
                "mat_03_btb": parity(mesh, "mat_03_btb", out_h.cpu().numpy(), np.array([2.0, 2.0, 0.0, 0.0]), axisymmetric=True),
                "vec_03_bv": parity(mesh, "vec_03_bv", out_v.cpu().numpy(), np.array([1.0, 2.0]), axisymmetric=True)}
        err = max(errs.values())
        ncell = mesh.ncell
        ntri = int((kinds == int(GeoKind.Tri3)).sum())
        nqua = ncell - ntri
        flops = (ntri * (1497 + 597) + nqua * (20376 + 5976)) / ncell
        byts = (ntri * (332 + 116) + nqua * (2192 + 656)) / ncell
        report("5: mixed tri3/qua8 axisymmetric mat_10_bdb + mat_03_btb + vec_03_bv", ncell, ms, flops, byts, fp64_peak, err,
               {"tri3_cells": ntri, "qua8_cells": nqua, "errs": errs})


if __name__ == "__main__":
    main()

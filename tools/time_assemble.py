"""Times the assembly hand-off (gl_triplet_indices + gl_assemble) on an n^3 Hex8 (or n^2 Qua8) block generated on the device.
The CSR pattern is built on the device with torch from the triplet indices.  usage: time_assemble.py [n] [planned-only|all] [hex8|qua8]"""
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch

import gemlab_b200 as g
from gemlab_b200 import Coef, CommonArgs, Context, GeoKind, integ

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
planned_only = len(sys.argv) > 2 and sys.argv[2] == "planned-only"  # for profiling the fused kernel alone
ctx = Context(0)
kind = sys.argv[3] if len(sys.argv) > 3 else "hex8"
ndim = 3 if kind == "hex8" else 2
dm = ctx.block(GeoKind.from_str(kind), (n,) * ndim)
ncell, npoint = dm.ncell, dm.npoint
neq = npoint * ndim
eq = torch.arange(neq, dtype=torch.int32, device="cuda")
dd = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, ndim == 2))
t0 = time.perf_counter()
rows, cols = integ.triplet_indices(ctx, dm, "mat_10_bdb", None, CommonArgs(), eq)
torch.cuda.synchronize()
t_trip = time.perf_counter() - t0
keys = torch.unique(rows.to(torch.int64) * neq + cols.to(torch.int64))
rowptr = torch.zeros(neq + 1, dtype=torch.int64, device="cuda")
rowptr[1:] = torch.cumsum(torch.bincount((keys // neq), minlength=neq), 0)
colidx = (keys % neq).to(torch.int32)
del keys, rows, cols
vals = torch.zeros(colidx.numel(), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
dt = 1.0
if not planned_only:
    integ.assemble(ctx, dm, "mat_10_bdb", None, CommonArgs(), Coef.uniform(dd), eq, vals, rowptr, colidx)
    torch.cuda.synchronize()
    vals.zero_()
    t0 = time.perf_counter()
    integ.assemble(ctx, dm, "mat_10_bdb", None, CommonArgs(), Coef.uniform(dd), eq, vals, rowptr, colidx)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
# K.1 = 0 for rigid translations: row sums of the assembled matrix vanish (relative to the largest entry)
rs = torch.zeros(neq, dtype=torch.float64, device="cuda")
rs.index_add_(0, torch.repeat_interleave(torch.arange(neq, device="cuda"), rowptr[1:] - rowptr[:-1]), vals)
print(f"n={n}: {ncell} cells, {neq} equations, nnz {colidx.numel()}; triplet indices {t_trip*1e3:.1f} ms; "
      f"assemble {dt*1e3:.1f} ms = {ncell/dt/1e6:.1f} M elements/s; max |row sum| / max |K| = {float(rs.abs().max() / vals.abs().max()):.1e}")

# planned assembly: position map once, then integrate + scatter fused into the stiffness kernel (no element-matrix slab)
t0 = time.perf_counter()
plan = integ.AssemblyPlan(ctx, dm, "mat_10_bdb", None, CommonArgs(), eq, rowptr, colidx)
torch.cuda.synchronize()
t_plan = time.perf_counter() - t0
ref_vals = vals.clone()
vals.zero_()
plan.assemble(Coef.uniform(dd), vals)
torch.cuda.synchronize()
err = float((vals - ref_vals).abs().max() / ref_vals.abs().max()) if not planned_only else float("nan")
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
vals.zero_()
e0.record()
for _ in range(5):
    plan.assemble(Coef.uniform(dd), vals)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"planned: plan {t_plan*1e3:.1f} ms (once); integrate + assemble {ms:.3f} ms = {ncell/ms/1e3:.1f} M elements/s; kernel {ctx.last_kernel()}; "
      f"max |planned - unplanned| / max |K| = {err:.1e}")

# symmetric storage: the upper triangle only (args.packed on the plan)
up = torch.repeat_interleave(torch.arange(neq, device="cuda"), rowptr[1:] - rowptr[:-1]) <= colidx
rowptr_u = torch.zeros(neq + 1, dtype=torch.int64, device="cuda")
rowptr_u[1:] = torch.cumsum(torch.bincount(torch.repeat_interleave(torch.arange(neq, device="cuda"), rowptr[1:] - rowptr[:-1])[up], minlength=neq), 0)
colidx_u = colidx[up].contiguous()
vals_u = torch.zeros(colidx_u.numel(), dtype=torch.float64, device="cuda")
plan_u = integ.AssemblyPlan(ctx, dm, "mat_10_bdb", None, CommonArgs(packed=True), eq, rowptr_u, colidx_u)
plan_u.assemble(Coef.uniform(dd), vals_u)
torch.cuda.synchronize()
err_u = float((vals_u - ref_vals[up] if not planned_only else vals_u * 0).abs().max() / vals_u.abs().max())
vals_u.zero_()
e0.record()
for _ in range(5):
    plan_u.assemble(Coef.uniform(dd), vals_u)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 5
print(f"planned, symmetric storage ({colidx_u.numel()} non-zeros): integrate + assemble {ms:.3f} ms = {ncell/ms/1e3:.1f} M elements/s; "
      f"max |upper - full| / max |K| = {err_u:.1e}")

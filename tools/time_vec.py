"""Times vec_04_bt (stress per Gauss point, device records) and vec_02_nv on an n^3 Hex8 block. usage: time_vec.py [n]"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import gemlab_b200 as g
from gemlab_b200 import Coef, CommonArgs, Context, GeoKind, integ
from tools.bench_configs import timed

n = int(sys.argv[1]) if len(sys.argv) > 1 else 100
ctx = Context(0)
dm = ctx.block(GeoKind.Hex8, (n, n, n))
sig = torch.randn((dm.ncell * 8, 6), dtype=torch.float64, device="cuda")
out = torch.empty(dm.ncell * 24, dtype=torch.float64, device="cuda")
ms = timed(lambda: integ.vec_04_bt(ctx, dm, None, CommonArgs(), Coef.per_gauss(sig), out=out), steps=5, warmup=2)
print(f"vec_04_bt per-Gauss stress: {dm.ncell} cells, {ms:.4f} ms, {dm.ncell / ms / 1e3:.1f} M el/s, kernel {ctx.last_kernel()}")
ms = timed(lambda: integ.vec_02_nv(ctx, dm, None, CommonArgs(), Coef.uniform([0.0, 0.0, -9.81]), out=out), steps=5, warmup=2)
print(f"vec_02_nv uniform: {dm.ncell} cells, {ms:.4f} ms, {dm.ncell / ms / 1e3:.1f} M el/s, kernel {ctx.last_kernel()}")

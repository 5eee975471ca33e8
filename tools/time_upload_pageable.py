import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
import gemlab_b200 as g
from gemlab_b200 import Block, Context, GeoKind
ctx = Context(0)
mesh = Block.new_cube(1.0).set_ndiv([100, 100, 100]).subdivide(GeoKind.Hex8)
for rep in range(4):
    t0 = time.perf_counter(); dm = ctx.upload(mesh); ctx.sync(); dt = time.perf_counter() - t0; dm.free()
print(f"upload of {mesh.ncell} Hex8 cells from pageable numpy arrays: {dt*1e3:.2f} ms")

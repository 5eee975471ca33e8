"""The five BASELINE.json configs as concrete synthetic inputs, shared by bench.py and the tools.

Every builder returns a dict: the host Mesh in its DISTORTED variant (interior points jittered, Hex20 mid-edge nodes bulged,
Tet10 cell order permuted -- seed 12345), the integration cases with their coefficients, and the algorithmic bytes / flops
per element from SURVEY.md 8(d).  `scale` shrinks the cell count (1.0 = the BASELINE sizes).
Nothing here touches the oracle; `parity_sample` is given the oracle module by its caller (bench.py / tests / tools).
"""
import numpy as np

import gemlab_b200 as g
from gemlab_b200 import Block, GeoKind, Mesh, integ

SEED = 12345


def bulge_mid_nodes(mesh, ncorner, amplitude, seed=SEED):
    """Moves the mid-edge nodes (local index >= ncorner) that are not on the bounding box by U(-amplitude, amplitude)."""
    rng = np.random.default_rng(seed + 1)
    conn = mesh.conn_matrix()
    mid = np.unique(conn[:, ncorner:]).astype(np.int64)
    lo, hi = mesh.coords.min(axis=0), mesh.coords.max(axis=0)
    tol = 1e-12 * max(1.0, float(np.max(hi - lo)))
    inner = np.all((mesh.coords[mid] > lo + tol) & (mesh.coords[mid] < hi - tol), axis=1)
    mid = mid[inner]
    coords = mesh.coords.copy()
    coords[mid] += rng.uniform(-amplitude, amplitude, size=(mid.size, mesh.ndim))
    return Mesh(mesh.ndim, coords, mesh.kinds, mesh.conn_off, mesh.conn, mesh.markers)


def mixed_tri3_qua8(n):
    """Config 4 shape: n x n lattice over x in [1, 2] (r > 0), alternating rows of Qua8 cells and Qua -> 2 x Tri3 splits; cell
    order follows the rows."""
    q8 = Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.Qua8)
    q8.coords[:, 0] += 1.0
    c = q8.conn_matrix().astype(np.int64)
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
    conn_off = np.concatenate([[0], np.cumsum(np.concatenate(lens))]).astype(np.uint64)
    return Mesh(2, q8.coords, kinds, conn_off, np.concatenate(conn_chunks).astype(np.uint64))


def build_config(idx, scale=1.0):
    """idx = index into BASELINE.json `configs` (0 .. 4)."""
    s = float(scale)
    if idx == 0:
        n = max(10, int(round(1000 * s ** 0.5)))
        mesh = g.jitter_interior_points(Block.new_square(1.0).set_ndiv([n, n]).subdivide(GeoKind.Qua4), 0.2 / n, SEED)
        dd = g.mandel_matrix(g.lin_elasticity(10_000.0, 0.2, True, False))
        return {"name": f"configs[0] qua4 {n}x{n} mat_10_bdb plane strain E=10000 nu=0.2, 2x2 Gauss, interior points jittered 0.2h",
                "mesh": mesh, "axisymmetric": False, "fused": False, "cases": [("mat_10_bdb", dd)],
                "bytes": 544.0, "flops": 1824.0}
    if idx == 1:
        n = max(4, int(round(200 * s ** (1 / 3))))
        mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex8), 0.2 / n, SEED)
        dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
        return {"name": f"configs[1] hex8 {n}^3 mat_10_bdb isotropic E=480 nu=1/3, 2x2x2 Gauss, interior points jittered 0.2h",
                "mesh": mesh, "axisymmetric": False, "fused": False, "cases": [("mat_10_bdb", dd)],
                "bytes": 4664.0, "flops": 37264.0}
    if idx == 2:
        n = max(4, int(round(100 * s ** (1 / 3))))
        mesh = Block.new_cube(1.0).set_ndiv([n, n, n]).subdivide(GeoKind.Hex20)
        mesh = g.jitter_interior_points(mesh, 0.05 / n, SEED)  # corners and mid nodes alike ...
        mesh = bulge_mid_nodes(mesh, 8, 0.05 / n)              # ... then the mid-edge nodes off their edges by 5 % of h
        dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
        return {"name": f"configs[2] hex20 {n}^3 mat_10_bdb isotropic, 27-point Gauss, mid-edge nodes bulged 5 % of h",
                "mesh": mesh, "axisymmetric": False, "fused": False, "cases": [("mat_10_bdb", dd)],
                "bytes": 28978.0, "flops": 662310.0}
    if idx == 3:
        n = max(3, int(round(94 * s ** (1 / 3))))
        mesh = g.split_hex_cells_into_tet10(n, n, n, seed=SEED)
        return {"name": f"configs[3] tet10 {mesh.ncell} cells (6 tets per cell of a {n}^3 lattice, cell order permuted) "
                        "mat_01_nsn rho=2.7 + vec_01_ns 9.81 rho, 14-point Gauss, one fused pass",
                "mesh": mesh, "axisymmetric": False, "fused": True,
                "cases": [("mat_01_nsn", np.array([2.7])), ("vec_01_ns", np.array([9.81 * 2.7]))],
                "bytes": 953.0, "flops": 12264.0}
    if idx == 4:
        n = max(8, int(round((10_000_000 * s / 1.5) ** 0.5)))
        mesh = g.jitter_interior_points(mixed_tri3_qua8(n), 0.05 / n, SEED)
        dd = g.mandel_matrix(g.lin_elasticity(96.0, 1 / 3, True, False))
        ntri = int((mesh.kinds == int(GeoKind.Tri3)).sum())
        nqua = mesh.ncell - ntri
        return {"name": f"configs[4] mixed tri3/qua8 {mesh.ncell} cells ({ntri} tri3 + {nqua} qua8, alternating rows), axisymmetric "
                        "mat_10_bdb E=96 nu=1/3 + mat_03_btb k=2 + vec_03_bv w=(1,2); Tri3 3-point, Qua8 9-point",
                "mesh": mesh, "axisymmetric": True, "fused": True,
                "cases": [("mat_10_bdb", dd), ("mat_03_btb", np.array([2.0, 2.0, 0.0, 0.0])), ("vec_03_bv", np.array([1.0, 2.0]))],
                "bytes": (ntri * (332 + 116) + nqua * (2192 + 656)) / mesh.ncell,
                "flops": (ntri * (1497 + 597) + nqua * (20376 + 5976)) / mesh.ncell}
    raise ValueError(idx)


def parity_sample(po, mesh, op, got, coef, nsample, nthreads=1, **kw):
    """max over a seeded sample of cells of max|got - oracle| / max|oracle| per element, and the number of cells compared.
    `got` is the batched output of `op` over the whole mesh (host array); `po` is the oracle module."""
    rng = np.random.default_rng(1)
    sel = np.sort(rng.choice(mesh.ncell, size=min(int(nsample), mesh.ncell), replace=False))
    sub = mesh.permuted(sel)
    ref = po.mesh_integ(op, sub.ndim, sub.coords, sub.kinds, sub.conn_off, sub.conn, coef=coef, nthreads=nthreads, **kw)
    sizes = np.array([integ.op_out_size(op, GeoKind(int(k)), mesh.ndim) for k in np.unique(mesh.kinds)], dtype=np.int64)
    size_of = dict(zip([int(k) for k in np.unique(mesh.kinds)], sizes))
    per_cell = np.array([size_of[int(k)] for k in mesh.kinds], dtype=np.int64) if len(size_of) > 1 else None
    if per_cell is None:
        n = int(sizes[0])
        gsel = got.reshape(mesh.ncell, n)[sel]
        r = ref.reshape(sel.size, n)
        err = np.max(np.abs(gsel - r), axis=1) / np.max(np.abs(r), axis=1)
        return float(err.max()), int(sel.size)
    off = np.concatenate([[0], np.cumsum(per_cell)])
    worst, pos = 0.0, 0
    for e in sel:
        n = int(per_cell[e])
        r = ref[pos:pos + n]
        worst = max(worst, float(np.max(np.abs(got[off[e]:off[e] + n] - r)) / np.max(np.abs(r))))
        pos += n
    return worst, int(sel.size)

"""Device-side assembly hand-off (gl_triplet_indices, gl_assemble) against a scipy assembly of the oracle's element
matrices with the reference's local-dof rule ii = i + m*ndim (src/integ/mat_10_bdb.rs:31-46)."""
import numpy as np
import pytest
import scipy.sparse as sp

import gemlab_b200 as g
from gemlab_b200 import Block, Coef, CommonArgs, GeoKind, Mesh, integ
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


def reference_triplets(mesh, op, coef, npd, eq, **kw):
    """rows, cols, vals of the CooMatrix::put loop a pmsim-style assembler runs over the oracle's element blocks."""
    blocks = po.mesh_integ(op, mesh.ndim, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, coef=coef, **kw)
    rows, cols, pos = [], [], 0
    for e in range(mesh.ncell):
        pts = mesh.conn[int(mesh.conn_off[e]):int(mesh.conn_off[e + 1])].astype(np.int64)
        l2g = (eq.reshape(-1, npd)[pts]).reshape(-1)  # local dof i + m*npd -> equation
        n = l2g.size
        if op.startswith("mat"):
            rows.append(np.tile(l2g, n))              # column-major block: entry q = ii + jj*n
            cols.append(np.repeat(l2g, n))
            pos += n * n
        else:
            rows.append(l2g)
            pos += n
    assert pos == blocks.size
    return np.concatenate(rows), (np.concatenate(cols) if cols else None), blocks


def mixed_mesh():
    q = Block.new_square(2.0).set_ndiv([6, 5]).subdivide(GeoKind.Qua4)
    q = g.jitter_interior_points(q, 0.03, 5)
    c = q.conn_matrix().astype(np.int64)
    kinds, conn, lens = [], [], []
    for e in range(q.ncell):
        if e % 3 == 0:  # split into two Tri3
            for tri in ([0, 1, 2], [0, 2, 3]):
                kinds.append(int(GeoKind.Tri3)); conn.append(c[e, tri]); lens.append(3)
        else:
            kinds.append(int(GeoKind.Qua4)); conn.append(c[e]); lens.append(4)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
    return Mesh(2, q.coords, np.array(kinds, dtype=np.uint8), off, np.concatenate(conn).astype(np.uint64))


@pytest.mark.parametrize("case", ["hex8", "tet10", "mixed2d"])
def test_triplets_and_csr_assembly_match_scipy(ctx, case):
    import torch

    if case == "hex8":
        mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([5, 4, 6]).subdivide(GeoKind.Hex8), 0.02, 3)
    elif case == "tet10":
        mesh = g.split_hex_cells_into_tet10(3, 3, 2, seed=4)
    else:
        mesh = mixed_mesh()
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, mesh.ndim == 2))
    npd = integ.dofs_per_node("mat_10_bdb", mesh.ndim)
    assert npd == mesh.ndim
    rng = np.random.default_rng(8)
    eq = rng.permutation(mesh.npoint * npd).astype(np.int32)
    eq[rng.choice(eq.size, size=eq.size // 10, replace=False)] = -1  # prescribed dofs are not assembled
    neq = mesh.npoint * npd
    rr, cc, vv = reference_triplets(mesh, "mat_10_bdb", dd, npd, eq)

    # 1. triplet indices line up with the batched output
    rows, cols = integ.triplet_indices(ctx, dmesh, "mat_10_bdb", None, CommonArgs(), eq)
    assert np.array_equal(rows.cpu().numpy(), rr) and np.array_equal(cols.cpu().numpy(), cc)

    # 2. CSR assembly == scipy's duplicate-summing conversion of the oracle triplets
    keep = (rr >= 0) & (cc >= 0)
    ref = sp.coo_matrix((vv[keep], (rr[keep], cc[keep])), shape=(neq, neq)).tocsr()
    ref.sum_duplicates()
    ref.sort_indices()
    rowptr = torch.tensor(ref.indptr.astype(np.int64), device="cuda")
    colidx = torch.tensor(ref.indices.astype(np.int32), device="cuda")
    vals = torch.zeros(ref.nnz, dtype=torch.float64, device="cuda")
    integ.assemble(ctx, dmesh, "mat_10_bdb", None, CommonArgs(), Coef.uniform(dd), eq, vals, rowptr, colidx)
    got = vals.cpu().numpy()
    assert np.max(np.abs(got - ref.data)) <= 1e-12 * np.max(np.abs(ref.data))
    # two half ranges accumulate to the same matrix (the hand-off is additive, like CooMatrix::put)
    vals2 = torch.zeros_like(vals)
    h = mesh.ncell // 2
    integ.assemble(ctx, dmesh, "mat_10_bdb", (0, h), CommonArgs(), Coef.uniform(dd), eq, vals2, rowptr, colidx)
    integ.assemble(ctx, dmesh, "mat_10_bdb", (h, mesh.ncell), CommonArgs(), Coef.uniform(dd), eq, vals2, rowptr, colidx)
    assert np.max(np.abs(vals2.cpu().numpy() - ref.data)) <= 1e-12 * np.max(np.abs(ref.data))

    # 3. a pattern that lacks an entry is reported, not silently dropped
    bad_cols = colidx.clone()
    bad_cols[0] = bad_cols[0] + 1 if ref.indices[0] + 1 != (ref.indices[1] if ref.nnz > 1 else -2) else bad_cols[0] + 2
    with pytest.raises(g.GemlabError):
        integ.assemble(ctx, dmesh, "mat_10_bdb", None, CommonArgs(), Coef.uniform(dd), eq, torch.zeros_like(vals), rowptr, bad_cols)


def test_vector_assembly_matches_numpy(ctx):
    import torch

    mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([4, 4, 3]).subdivide(GeoKind.Hex20), 0.01, 2)
    dmesh = ctx.upload(mesh)
    for op, coef in [("vec_01_ns", np.array([9.81])), ("vec_02_nv", np.array([0.0, 0.0, -9.81]))]:
        npd = integ.dofs_per_node(op, 3)
        eq = np.arange(mesh.npoint * npd, dtype=np.int32)[::-1].copy()
        eq[::7] = -1
        rr, _, vv = reference_triplets(mesh, op, coef, npd, eq)
        ref = np.zeros(mesh.npoint * npd)
        np.add.at(ref, rr[rr >= 0], vv[rr >= 0])
        rhs = torch.zeros(ref.size, dtype=torch.float64, device="cuda")
        integ.assemble(ctx, dmesh, op, None, CommonArgs(), Coef.uniform(coef), eq, rhs)
        assert np.max(np.abs(rhs.cpu().numpy() - ref)) <= 1e-12 * np.max(np.abs(ref))
        rows, cols = integ.triplet_indices(ctx, dmesh, op, None, CommonArgs(), eq)
        assert cols is None and np.array_equal(rows.cpu().numpy(), rr)


def _csr_from_triplets(rr, cc, vv, neq):
    keep = (rr >= 0) & (cc >= 0)
    ref = sp.coo_matrix((vv[keep], (rr[keep], cc[keep])), shape=(neq, neq)).tocsr()
    ref.sum_duplicates()
    ref.sort_indices()
    return ref


@pytest.mark.parametrize("case", ["hex8", "hex8_27", "tet10", "hex20", "qua8_axi", "mixed2d"])
@pytest.mark.parametrize("coef_mode", ["uniform", "general", "per_cell", "per_gauss"])
def test_planned_assembly_fuses_the_scatter_into_the_stiffness_kernels(ctx, case, coef_mode):
    """gl_assemble_plan_create + gl_assemble_planned against scipy's duplicate-summing conversion of the oracle's element matrices
    (local-dof rule ii = i + m*ndim, src/integ/mat_10_bdb.rs:31-46).  The stiffness blocks must be scattered by the integration
    kernel itself (kernel family ends with ',scatter>'): the element-matrix slab is never written."""
    import torch

    from test_gpu_parity import random_coef

    args = CommonArgs()
    if case == "hex8" or case == "hex8_27":
        mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([5, 4, 6]).subdivide(GeoKind.Hex8), 0.02, 3)
        if case == "hex8_27":
            args = CommonArgs(ngauss=27)
    elif case == "tet10":
        mesh = g.split_hex_cells_into_tet10(3, 3, 2, seed=4)
    elif case == "hex20":
        mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([3, 2, 3]).subdivide(GeoKind.Hex20), 0.01, 3)
    elif case == "qua8_axi":
        mesh = Block.new_square(2.0).set_ndiv([5, 6]).subdivide(GeoKind.Qua8)
        mesh.coords[:, 0] += 1.0
        mesh = g.jitter_interior_points(mesh, 0.02, 3)
        args = CommonArgs(axisymmetric=True)
    else:
        mesh = mixed_mesh()
    dmesh = ctx.upload(mesh)
    ndim, ncell = mesh.ndim, mesh.ncell
    ns = 2 * ndim
    rng = np.random.default_rng(21)
    ng = None
    if coef_mode == "uniform":
        rec = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, ndim == 2))
        coef, okw = Coef.uniform(rec), dict(coef=rec)
    elif coef_mode == "general":
        rec = random_coef("mat_10_bdb", ndim, 1, rng, spd=False)[0]
        coef, okw = Coef.uniform(rec), dict(coef=rec)
    elif coef_mode == "per_cell":
        recs = random_coef("mat_10_bdb", ndim, ncell, rng)
        coef, okw = Coef.per_cell(recs), dict(coef=recs, coef_cell_stride=ns * ns)
    else:
        if case == "mixed2d":
            pytest.skip("PER_GAUSS needs one rule size for all kinds in the range")
        kind = GeoKind(int(mesh.kinds[0]))
        ng = g.Gauss.new_or_sized(kind, args.ngauss).npoint()
        recs = random_coef("mat_10_bdb", ndim, ncell * ng, rng)
        coef, okw = Coef.per_gauss(recs), dict(coef=recs, coef_cell_stride=ng * ns * ns, coef_gp_stride=ns * ns)
    if args.ngauss:
        okw["ngauss"] = args.ngauss
    if args.axisymmetric:
        okw["axisymmetric"] = True
    eq = rng.permutation(mesh.npoint * ndim).astype(np.int32)
    eq[rng.choice(eq.size, size=eq.size // 10, replace=False)] = -1
    neq = mesh.npoint * ndim
    rr, cc, vv = reference_triplets(mesh, "mat_10_bdb", npd=ndim, eq=eq, **okw)
    ref = _csr_from_triplets(rr, cc, vv, neq)
    rowptr = torch.tensor(ref.indptr.astype(np.int64), device="cuda")
    colidx = torch.tensor(ref.indices.astype(np.int32), device="cuda")
    plan = integ.AssemblyPlan(ctx, dmesh, "mat_10_bdb", None, args, eq, rowptr, colidx)
    vals = torch.zeros(ref.nnz, dtype=torch.float64, device="cuda")
    for it in range(2):  # the plan is reusable: two assemblies accumulate twice the matrix
        plan.assemble(coef, vals)
        ctx.sync()
        assert ctx.last_kernel().endswith(",scatter>"), ctx.last_kernel()
        got = vals.cpu().numpy()
        assert np.max(np.abs(got - (it + 1) * ref.data)) <= 2e-12 * np.max(np.abs(ref.data)), f"{case} {coef_mode} pass {it}"
    plan.free()
    # a pattern that lacks an entry: no plan
    bad_cols = colidx.clone()
    bad_cols[0] = bad_cols[0] + 1 if ref.indices[0] + 1 != (ref.indices[1] if ref.nnz > 1 else -2) else bad_cols[0] + 2
    with pytest.raises(g.GemlabError):
        integ.AssemblyPlan(ctx, dmesh, "mat_10_bdb", None, args, eq, rowptr, bad_cols)


def test_planned_assembly_of_the_other_cases(ctx):
    """Mass matrix (slab + position-map scatter) and a load vector through a plan."""
    import torch

    mesh = g.split_hex_cells_into_tet10(3, 2, 2, seed=6)
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(2)
    eq = rng.permutation(mesh.npoint).astype(np.int32)
    eq[::9] = -1
    rho = np.array([2.7])
    rr, cc, vv = reference_triplets(mesh, "mat_01_nsn", rho, 1, eq)
    ref = _csr_from_triplets(rr, cc, vv, mesh.npoint)
    rowptr = torch.tensor(ref.indptr.astype(np.int64), device="cuda")
    colidx = torch.tensor(ref.indices.astype(np.int32), device="cuda")
    plan = integ.AssemblyPlan(ctx, dmesh, "mat_01_nsn", None, CommonArgs(), eq, rowptr, colidx)
    vals = torch.zeros(ref.nnz, dtype=torch.float64, device="cuda")
    plan.assemble(Coef.uniform(rho), vals)
    ctx.sync()
    assert np.max(np.abs(vals.cpu().numpy() - ref.data)) <= 1e-12 * np.max(np.abs(ref.data))
    # vector case
    eq3 = rng.permutation(mesh.npoint * 3).astype(np.int32)
    eq3[::5] = -1
    bf = np.array([0.0, -1.0, -9.81])
    rr, _, vv = reference_triplets(mesh, "vec_02_nv", bf, 3, eq3)
    refv = np.zeros(mesh.npoint * 3)
    np.add.at(refv, rr[rr >= 0], vv[rr >= 0])
    planv = integ.AssemblyPlan(ctx, dmesh, "vec_02_nv", None, CommonArgs(), eq3)
    rhs = torch.zeros(refv.size, dtype=torch.float64, device="cuda")
    planv.assemble(Coef.uniform(bf), rhs)
    ctx.sync()
    assert np.max(np.abs(rhs.cpu().numpy() - refv)) <= 1e-12 * np.max(np.abs(refv))


def test_planned_assembly_into_symmetric_storage(ctx):
    """args.packed on a plan: only the upper triangle (global row <= column) of the matrix is assembled."""
    import torch

    mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([4, 3, 5]).subdivide(GeoKind.Hex8), 0.02, 3)
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, False))
    rng = np.random.default_rng(5)
    eq = rng.permutation(mesh.npoint * 3).astype(np.int32)
    eq[::11] = -1
    rr, cc, vv = reference_triplets(mesh, "mat_10_bdb", dd, 3, eq)
    keep = (rr >= 0) & (cc >= 0) & (rr <= cc)
    ref = sp.coo_matrix((vv[keep], (rr[keep], cc[keep])), shape=(eq.size, eq.size)).tocsr()
    ref.sum_duplicates()
    ref.sort_indices()
    rowptr = torch.tensor(ref.indptr.astype(np.int64), device="cuda")
    colidx = torch.tensor(ref.indices.astype(np.int32), device="cuda")
    plan = integ.AssemblyPlan(ctx, dmesh, "mat_10_bdb", None, CommonArgs(packed=True), eq, rowptr, colidx)
    vals = torch.zeros(ref.nnz, dtype=torch.float64, device="cuda")
    plan.assemble(Coef.uniform(dd), vals)
    ctx.sync()
    assert ctx.last_kernel().endswith(",scatter>")
    assert np.max(np.abs(vals.cpu().numpy() - ref.data)) <= 1e-12 * np.max(np.abs(ref.data))
    # the full pattern is still required without the flag: an upper-only pattern then lacks entries
    with pytest.raises(g.GemlabError):
        integ.AssemblyPlan(ctx, dmesh, "mat_10_bdb", None, CommonArgs(), eq, rowptr, colidx)

"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the reference's known answers.

Bar (BASELINE.json north_star): per element max|K_gpu - K_ref| / max|K_ref| <= 1e-12 (f64); cell order, element-block
offsets and the index layout ii = i + m*ndim are exact.  Run on the B200 box with `pytest -m gpu`.
"""
import json
from pathlib import Path

import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Block, Coef, CommonArgs, GeoKind, Mesh, integ
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

GOLD = json.loads((Path(__file__).parent / "golden" / "known_answers.json").read_text())
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


def rel_err(got, ref, sizes=None):
    """per-element max|d| / max|ref|; elements delimited by `sizes` offsets (or one element)."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    if sizes is None:
        scale = np.max(np.abs(ref))
        return float(np.max(np.abs(got - ref)) / (scale if scale > 0 else 1.0))
    worst = 0.0
    for a, b in zip(sizes[:-1], sizes[1:]):
        a, b = int(a), int(b)
        scale = np.max(np.abs(ref[a:b]))
        worst = max(worst, float(np.max(np.abs(got[a:b] - ref[a:b])) / (scale if scale > 0 else 1.0)))
    return worst


def oracle_integ(op, mesh, coef=None, cell_stride=0, gp_stride=0, **kw):
    return po.mesh_integ(op, mesh.ndim, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, coef=coef,
                         coef_cell_stride=cell_stride, coef_gp_stride=gp_stride, **kw)


# ---------------------------------------------------------------------------------------------------------------
# 1. the reference's own known-answer tests, run through the GPU path (single-cell meshes)
# ---------------------------------------------------------------------------------------------------------------

def _golden_coef(case, mesh, ngauss):
    op = case["op"]
    if op == "mat_10_bdb":
        return Coef.uniform(np.array(case["dd"]).ravel(order="F"))
    if op in ("mat_03_btb", "vec_04_bt", "mat_08_ntn"):
        return Coef.uniform(case["tt"])
    if "s" in case:
        return Coef.uniform([case["s"]])
    if "v" in case:
        return Coef.uniform(case["v"])
    # x-dependent coefficients: evaluate at the Gauss points with the oracle's get_points_coords -> PER_GAUSS records
    pad = po.Scratchpad.new(case["ndim"], case["kind"]).set_coords(case["xx"])
    gauss = po.Gauss(case["kind"], ngauss)
    x = np.array(po.get_points_coords(pad, gauss))
    if "s_field" in case:
        return Coef.per_gauss(x[:, int(case["s_field"][1])].copy())
    if case.get("v_field") == "x":
        return Coef.per_gauss(x.copy())
    if case.get("v_field") == "x0x0":
        return Coef.per_gauss(np.repeat(x[:, [0]], case["ndim"], axis=1).copy())
    raise AssertionError(case["name"])


@pytest.mark.parametrize("case", GOLD["cases"], ids=[c["name"] for c in GOLD["cases"]])
def test_reference_known_answers_on_gpu(ctx, case):
    kind = GeoKind.from_str(case["kind"])
    xx = np.array(case["xx"])
    mesh = Mesh.homogeneous(case["ndim"], xx, kind, np.arange(kind.nnode()).reshape(1, -1))
    dmesh = ctx.upload(mesh)
    expect = np.array(case["expect"])
    tols = case["tol"] if isinstance(case["tol"], list) else [case["tol"]] * len(case["ngauss"])
    for n, tol in zip(case["ngauss"], tols):
        args = CommonArgs(ngauss=None if n == 0 else n, alpha=case.get("alpha", 1.0),
                          axisymmetric=case.get("axisymmetric", False))
        # outputs pre-filled with NOISE (src/integ/testing.rs:7): clear=True must overwrite
        out = np.full(expect.size, GOLD["noise"])
        fn = getattr(integ, case["op"])
        fn(ctx, dmesh, None, args, _golden_coef(case, mesh, n), out=out)
        got = out.reshape(expect.shape, order="F") if expect.ndim == 2 else out
        assert np.max(np.abs(got - expect)) <= max(tol, 4e-16 * np.max(np.abs(expect))), f"{case['name']} n={n} ({case['ref']})"


def test_jacobian_doctest_on_gpu(ctx):
    j = GOLD["jacobian_doctest"]
    mesh = Mesh.homogeneous(2, np.array(j["xx"]), GeoKind.Qua4, [[0, 1, 2, 3]])
    det = integ.jacobian(ctx, ctx.upload(mesh), None, j["ksi"])
    assert abs(det[0] - j["det"]) <= j["tol"]


@pytest.mark.parametrize("sf", GOLD["scalar_field"], ids=[s["name"] for s in GOLD["scalar_field"]])
def test_scalar_field_integrals_on_gpu(ctx, sf):
    """integ::scalar_field (src/integ/scalar_field.rs:153-230): sum of vec_01_ns entries = int s dOmega."""
    kind = GeoKind.from_str(sf["kind"])
    mesh = Mesh.homogeneous(sf["ndim"], np.array(sf["xx"], dtype=float), kind, np.arange(kind.nnode()).reshape(1, -1))
    dmesh = ctx.upload(mesh)
    for n in sf["ngauss"]:
        a = integ.vec_01_ns(ctx, dmesh, None, CommonArgs(ngauss=n), Coef.uniform([1.0]))
        assert abs(a.sum() - sf.get("volume", sf.get("area"))) <= 1e-13
        if "int_x2" in sf:
            pad = po.Scratchpad.new(sf["ndim"], sf["kind"]).set_coords(sf["xx"])
            x = np.array(po.get_points_coords(pad, po.Gauss(sf["kind"], n)))
            a2 = integ.vec_01_ns(ctx, dmesh, None, CommonArgs(ngauss=n), Coef.per_gauss(np.sum(x**2, axis=1)))
            a3 = integ.vec_01_ns(ctx, dmesh, None, CommonArgs(ngauss=n), Coef.per_gauss(np.sum(x**3, axis=1)))
            assert abs(a2.sum() - sf["int_x2"]) <= 1e-13
            assert abs(a3.sum() - sf["int_x3"]) <= 1e-13


# ---------------------------------------------------------------------------------------------------------------
# 2. distorted meshes of every hot-path kind: all ops, GPU vs oracle
# ---------------------------------------------------------------------------------------------------------------

def make_mesh(kind, n=5, seed=3):
    if kind in ("qua4", "qua8", "qua9"):
        m = Block.new_square(2.0).set_ndiv([n, n + 1]).subdivide(GeoKind.from_str(kind))
        m.coords[:, 0] += 1.0  # r > 0 for axisymmetric cases
        return g.jitter_interior_points(m, 0.05 * 2.0 / n, seed)
    if kind in ("hex8", "hex20"):
        m = Block.new_cube(1.5).set_ndiv([n, n - 1, n + 1]).subdivide(GeoKind.from_str(kind))
        return g.jitter_interior_points(m, 0.05 * 1.5 / n, seed)
    if kind == "tet10":
        return g.jitter_interior_points(g.split_hex_cells_into_tet10(n, n, n - 1, seed=seed), 0.02 / n, seed)
    if kind == "tet4":
        t10 = g.split_hex_cells_into_tet10(n, n, n - 1, seed=seed)
        conn = t10.conn_matrix()[:, :4]
        used, inv = np.unique(conn, return_inverse=True)
        m = Mesh.homogeneous(3, t10.coords[used.astype(np.int64)], GeoKind.Tet4, inv.reshape(-1, 4))
        return g.jitter_interior_points(m, 0.05 / n, seed)
    if kind in ("tri3", "tri6"):
        q = Block.new_square(2.0).set_ndiv([n, n]).subdivide(GeoKind.Qua9 if kind == "tri6" else GeoKind.Qua4)
        q.coords[:, 0] += 1.0
        c = q.conn_matrix()
        if kind == "tri3":
            tris = np.concatenate([c[:, [0, 1, 2]], c[:, [0, 2, 3]]])
        else:  # split a Qua9 along the 0-2 diagonal: mid nodes 4 (0-1), 5 (1-2), 8 (centre), 6 (2-3), 7 (3-0)
            tris = np.concatenate([c[:, [0, 1, 2, 4, 5, 8]], c[:, [0, 2, 3, 8, 6, 7]]])
        m = Mesh.homogeneous(2, q.coords, GeoKind.from_str(kind), tris)
        return g.jitter_interior_points(m, 0.04 * 2.0 / n, seed)
    raise AssertionError(kind)


def random_coef(op, ndim, nrec, rng, spd=True):
    ns = 2 * ndim
    if op == "mat_10_bdb":
        a = rng.standard_normal((nrec, ns, ns))
        d = a @ np.transpose(a, (0, 2, 1)) + ns * np.eye(ns) if spd else a  # column-major flatten of a symmetric matrix
        return np.ascontiguousarray(np.transpose(d, (0, 2, 1)).reshape(nrec, ns * ns))
    n = integ.op_coef_size(op, ndim)
    return rng.standard_normal((nrec, n)) + 2.0


SOLID_KINDS = ["tri3", "tri6", "qua4", "qua8", "qua9", "tet4", "tet10", "hex8", "hex20"]
OPS = ["mat_10_bdb", "mat_01_nsn", "mat_02_bvn", "mat_03_btb", "mat_08_ntn", "mat_09_nvb", "vec_01_ns", "vec_02_nv", "vec_03_bv",
       "vec_04_bt"]


@pytest.mark.parametrize("kind", SOLID_KINDS)
@pytest.mark.parametrize("op", OPS)
def test_all_ops_uniform_coefficient(ctx, kind, op):
    mesh = make_mesh(kind)
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(11)
    coef = random_coef(op, mesh.ndim, 1, rng)[0]
    for axisym in ([False, True] if mesh.ndim == 2 else [False]):
        args = CommonArgs(alpha=0.7, axisymmetric=axisym)
        got = getattr(integ, op)(ctx, dmesh, None, args, Coef.uniform(coef))
        ref = oracle_integ(op, mesh, coef, alpha=0.7, axisymmetric=axisym)
        off = integ.out_offsets(dmesh, op, None, args)
        assert off[-1] == ref.size
        assert rel_err(got, ref, off) <= TOL, f"{kind} {op} axisym={axisym}"


@pytest.mark.parametrize("kind", ["qua4", "qua8", "tri3", "hex8", "hex20", "tet10"])
def test_parity_does_not_depend_on_mesh_conditioning(ctx, kind):
    """Cells far from the origin (|x|/h ~ 1e5, as in a 10^7-cell mesh): J = Xt.L cancels 5 digits, so two correct
    evaluations that sum in a different order differ by ~1e-11.  The kernels evaluate the geometry in the oracle's
    operation order (gl_geom.cuh), so the parity bar still holds -- for every kernel family (tuned, fused, generic)."""
    mesh = make_mesh(kind, n=4)
    mesh.coords[:] = mesh.coords * 0.01 + 1000.0
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(21)
    for op in ["mat_10_bdb", "mat_01_nsn", "mat_03_btb", "vec_03_bv", "vec_04_bt"]:
        coef = random_coef(op, mesh.ndim, 1, rng)[0]
        for axisym in ([False, True] if mesh.ndim == 2 else [False]):
            args = CommonArgs(axisymmetric=axisym)
            got = getattr(integ, op)(ctx, dmesh, None, args, Coef.uniform(coef))
            ref = oracle_integ(op, mesh, coef, axisymmetric=axisym)
            off = integ.out_offsets(dmesh, op, None, args)
            assert rel_err(got, ref, off) <= TOL, f"{kind} {op} axisym={axisym}"


@pytest.mark.parametrize("kind", ["qua8", "hex8", "tet10"])
def test_non_symmetric_modulus_and_rules(ctx, kind):
    """D need not be major-symmetric (the API only requires minor symmetry); also sweeps the sized Gauss rules."""
    mesh = make_mesh(kind, n=3)
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(5)
    coef = random_coef("mat_10_bdb", mesh.ndim, 1, rng, spd=False)[0]
    rules = {"qua8": [4, 9, 16], "hex8": [6, 8, 14, 27, 64], "tet10": [4, 5, 8, 14, 15, 24]}[kind]
    for n in rules:
        got = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(ngauss=n), Coef.uniform(coef))
        ref = oracle_integ("mat_10_bdb", mesh, coef, ngauss=n)
        bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), mesh.ndim)
        assert rel_err(got, ref, np.arange(mesh.ncell + 1) * bs) <= TOL, f"{kind} rule {n}"


@pytest.mark.parametrize("kind,op", [("hex8", "mat_10_bdb"), ("qua4", "mat_10_bdb"), ("tet10", "mat_01_nsn"),
                                     ("tri3", "vec_03_bv"), ("qua8", "mat_03_btb"), ("hex20", "vec_04_bt")])
def test_coefficient_modes(ctx, kind, op):
    import torch

    mesh = make_mesh(kind, n=4)
    ncell = mesh.ncell
    rng = np.random.default_rng(17)
    mesh.markers[:] = rng.integers(0, 3, ncell) * 10 - 5  # markers -5, 5, 15
    dmesh = ctx.upload(mesh)
    ng = g.Gauss.new(GeoKind.from_str(kind)).npoint()
    clen = integ.op_coef_size(op, mesh.ndim)
    args = CommonArgs()
    bs = integ.op_out_size(op, GeoKind.from_str(kind), mesh.ndim)
    sizes = np.arange(ncell + 1) * bs
    fn = getattr(integ, op)
    # PER_MARKER
    recs = random_coef(op, mesh.ndim, 3, rng)
    got = fn(ctx, dmesh, None, args, Coef.per_marker([15, -5, 5], recs))
    per_cell = recs[np.searchsorted([-5, 5, 15], mesh.markers)][:, :]
    per_cell = recs[[{15: 0, -5: 1, 5: 2}[int(m)] for m in mesh.markers]]
    ref = oracle_integ(op, mesh, per_cell, cell_stride=clen)
    assert rel_err(got, ref, sizes) <= TOL
    # PER_CELL (host) and a sub-range
    recs = random_coef(op, mesh.ndim, ncell, rng)
    got = fn(ctx, dmesh, None, args, Coef.per_cell(recs))
    ref = oracle_integ(op, mesh, recs, cell_stride=clen)
    assert rel_err(got, ref, sizes) <= TOL
    b, e = ncell // 3, ncell - 2
    got = fn(ctx, dmesh, (b, e), args, Coef.per_cell(recs[b:e]))
    assert rel_err(got, ref[b * bs:e * bs], sizes[: e - b + 1]) <= TOL
    # PER_GAUSS from a device tensor, device output
    recs = random_coef(op, mesh.ndim, ncell * ng, rng)
    out = torch.full((ncell * bs,), 7.0, dtype=torch.float64, device="cuda")
    fn(ctx, dmesh, None, args, Coef.per_gauss(torch.from_numpy(recs).cuda()), out=out)
    ctx.sync()
    ref = oracle_integ(op, mesh, recs, cell_stride=ng * clen, gp_stride=clen)
    assert rel_err(out.cpu().numpy(), ref, sizes) <= TOL


@pytest.mark.parametrize("kind", ["tri3", "tri6", "qua4", "qua8", "qua9", "tet4", "tet10", "hex8", "hex20"])
@pytest.mark.parametrize("structure", ["isotropic", "symmetric", "general"])
def test_stiffness_with_one_modulus_per_marker_and_per_cell(ctx, kind, structure):
    """Multi-material meshes through the tuned stiffness kernels: the record structure common to all records (isotropic /
    merely symmetric / general) selects the kernel variant; device-resident records take the general one."""
    import torch

    mesh = make_mesh(kind, n=3)
    ncell = mesh.ncell
    rng = np.random.default_rng(23)
    mesh.markers[:] = rng.integers(1, 4, ncell)
    dmesh = ctx.upload(mesh)
    two_dim = mesh.ndim == 2
    if structure == "isotropic":
        recs = np.stack([g.mandel_matrix(g.lin_elasticity(100.0 * (k + 1), 0.1 * (k + 1), two_dim)) for k in range(3)])
    else:
        recs = random_coef("mat_10_bdb", mesh.ndim, 3, rng, spd=(structure == "symmetric"))
    ns2 = (2 * mesh.ndim) ** 2
    bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), mesh.ndim)
    sizes = np.arange(ncell + 1) * bs
    per_cell = recs[mesh.markers - 1]
    for axisym in ([False, True] if two_dim else [False]):
        args = CommonArgs(axisymmetric=axisym)
        ref = oracle_integ("mat_10_bdb", mesh, per_cell, cell_stride=ns2, axisymmetric=axisym)
        got = integ.mat_10_bdb(ctx, dmesh, None, args, Coef.per_marker([1, 2, 3], recs))
        assert rel_err(got, ref, sizes) <= TOL, f"per_marker axisym={axisym}"
        got = integ.mat_10_bdb(ctx, dmesh, None, args, Coef.per_cell(per_cell))
        assert rel_err(got, ref, sizes) <= TOL, f"per_cell host axisym={axisym}"
        out = torch.empty(ncell * bs, dtype=torch.float64, device="cuda")
        integ.mat_10_bdb(ctx, dmesh, None, args, Coef.per_cell(torch.from_numpy(np.ascontiguousarray(per_cell)).cuda()), out=out)
        ctx.sync()
        assert rel_err(out.cpu().numpy(), ref, sizes) <= TOL, f"per_cell device axisym={axisym}"


# ---------------------------------------------------------------------------------------------------------------
# 3. layout: mixed kinds, cell ranges, offsets, accumulation, ii0/jj0
# ---------------------------------------------------------------------------------------------------------------

def mixed_mesh_2d(n=6, seed=2):
    """Config 5 shape: alternating stripes of Qua8 cells and Qua -> 2 x Tri3 splits over x in [1, 3]."""
    q8 = Block.new_square(2.0).set_ndiv([n, n]).subdivide(GeoKind.Qua8)
    q8.coords[:, 0] += 1.0
    c = q8.conn_matrix()
    cells = []
    for e in range(q8.ncell):
        row = e // n
        if row % 2 == 0:
            cells.append((GeoKind.Qua8, c[e]))
        else:
            cells.append((GeoKind.Tri3, c[e, [0, 1, 2]]))
            cells.append((GeoKind.Tri3, c[e, [0, 2, 3]]))
    m = Mesh.from_cells(2, q8.coords, cells)
    return g.jitter_interior_points(m, 0.01, seed)


@pytest.mark.parametrize("op", ["mat_10_bdb", "mat_03_btb", "vec_03_bv", "vec_01_ns"])
def test_mixed_kind_mesh_layout_and_values(ctx, op):
    mesh = mixed_mesh_2d()
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(23)
    coef = random_coef(op, 2, 1, rng)[0]
    args = CommonArgs(axisymmetric=True)
    off = integ.out_offsets(dmesh, op, None, args)
    sizes = np.array([integ.op_out_size(op, GeoKind(int(k)), 2) for k in mesh.kinds])
    np.testing.assert_array_equal(off, np.concatenate([[0], np.cumsum(sizes)]))  # exact layout
    got = getattr(integ, op)(ctx, dmesh, None, args, Coef.uniform(coef))
    ref = oracle_integ(op, mesh, coef, axisymmetric=True)
    assert rel_err(got, ref, off) <= TOL
    # a range that starts and ends inside different stripes
    b, e = 4, mesh.ncell - 5
    sub = getattr(integ, op)(ctx, dmesh, (b, e), args, Coef.uniform(coef))
    np.testing.assert_array_equal(sub, got[int(off[b]):int(off[e])])  # same arithmetic => bit-identical


def test_partition_concatenation_is_bit_identical(ctx):
    mesh = make_mesh("hex8", n=6)
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
    full = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd))
    for world in (2, 3, 8):
        parts = [integ.mat_10_bdb(ctx, dmesh, g.partition(mesh.ncell, world, r), CommonArgs(), Coef.uniform(dd))
                 for r in range(world)]
        np.testing.assert_array_equal(np.concatenate(parts), full)


def test_clear_false_accumulates_and_offsets(ctx):
    mesh = make_mesh("qua4", n=3)
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(10_000.0, 0.2, True))
    base = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd))
    out = np.full(base.size, 1234.56)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(clear=False), Coef.uniform(dd), out=out)
    # (the accumulate path runs on the generic kernel, the overwrite path on the register-blocked one)
    np.testing.assert_allclose(out, base + 1234.56, rtol=1e-13, atol=0)
    # ii0 / jj0 inside a larger (nrow x ncol) block, as for coupled elements (CommonArgs.ii0/jj0)
    args = CommonArgs(ii0=2, jj0=3, nrow=12, ncol=13)
    blk = integ.mat_10_bdb(ctx, dmesh, None, args, Coef.uniform(dd)).reshape(mesh.ncell, 13, 12)  # [cell][col][row]
    kk = base.reshape(mesh.ncell, 8, 8)
    np.testing.assert_allclose(blk[:, 3:11, 2:10], kk, rtol=0, atol=1e-13 * np.max(np.abs(kk)))
    outside = blk.copy()
    outside[:, 3:11, 2:10] = 0.0
    assert np.count_nonzero(outside) == 0
    # empty range
    assert integ.mat_10_bdb(ctx, dmesh, (2, 2), CommonArgs(), Coef.uniform(dd)).size == 0


def test_host_path_chunking_matches_device_path(ctx):
    """The host-output path pipelines 128 MiB chunks; results must equal the device path bit for bit."""
    import torch

    mesh = Block.new_cube(1.0).set_ndiv([40, 40, 30]).subdivide(GeoKind.Hex8)  # 48k cells * 4608 B = 221 MB -> 2 chunks
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False))
    host = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd))
    dev = torch.empty(host.size, dtype=torch.float64, device="cuda")
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd), out=dev)
    ctx.sync()
    np.testing.assert_array_equal(host, dev.cpu().numpy())
    # every cell of an undistorted block has the same matrix; rigid translations are in the null space
    k0 = host[:576].reshape(24, 24, order="F")
    np.testing.assert_allclose(host.reshape(-1, 576), np.broadcast_to(host[:576], (mesh.ncell, 576)), rtol=0, atol=1e-11)
    for i in range(3):
        t = np.zeros(24)
        t[i::3] = 1.0
        assert np.max(np.abs(k0 @ t)) <= 1e-12 * np.max(np.abs(k0))
    np.testing.assert_allclose(k0, k0.T, rtol=0, atol=1e-12 * np.max(np.abs(k0)))


def test_pageable_and_pinned_host_outputs_agree_with_the_device_path(ctx):
    """Host outputs in pageable memory (a numpy array, a Rust Vec) are drained through pinned bounce buffers in 16 MB pieces, one
    128 MiB chunk behind the kernels; pinned outputs take the direct copy.  On a jittered mesh (no two element matrices alike) both
    must equal the device-resident result bit for bit, also when accumulating (clear = false reads the destination first)."""
    import torch

    mesh = g.jitter_interior_points(Block.new_cube(1.0).set_ndiv([40, 40, 35]).subdivide(GeoKind.Hex8), 0.2 / 40, 17)  # 258 MB: 3 chunks
    dmesh = ctx.upload(mesh)
    dd = Coef.uniform(g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False)))
    dev = torch.empty(mesh.ncell * 576, dtype=torch.float64, device="cuda")
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), dd, out=dev)
    ctx.sync()
    ref = dev.cpu().numpy()
    pageable = np.full(ref.size, -3.0)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), dd, out=pageable)
    np.testing.assert_array_equal(pageable, ref)
    pinned = torch.full((ref.size,), -3.0, dtype=torch.float64).pin_memory()
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), dd, out=pinned)
    np.testing.assert_array_equal(pinned.numpy(), ref)
    acc = np.full(ref.size, 0.5)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(clear=False), dd, out=acc)
    np.testing.assert_allclose(acc, ref + 0.5, rtol=1e-13, atol=1e-13 * float(np.abs(ref).max()))
    # a ragged range through the packed path (upper triangles) into pageable memory
    b, e = 7, mesh.ncell - 5
    packed = integ.mat_10_bdb(ctx, dmesh, (b, e), CommonArgs(packed=True), dd)
    jj = np.concatenate([np.full(j + 1, j) for j in range(24)])  # packed order: column by column, rows 0 .. j
    ii = np.concatenate([np.arange(j + 1) for j in range(24)])
    full = ref.reshape(mesh.ncell, 24, 24)[b:e]  # column-major blocks: [cell][col][row]
    np.testing.assert_array_equal(packed.reshape(e - b, 300), full[:, jj, ii])


# ---------------------------------------------------------------------------------------------------------------
# 4. error behaviour (the reference's strings)
# ---------------------------------------------------------------------------------------------------------------

def test_error_strings(ctx):
    mesh = make_mesh("qua4", n=2)
    dmesh = ctx.upload(mesh)
    dd = Coef.uniform(g.mandel_matrix(g.lin_elasticity(1.0, 0.25, True)))
    with pytest.raises(g.GemlabError, match=r"nrow\(K\) must be ≥ ii0 \+ nnode ⋅ space_ndim"):
        integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(ii0=1, nrow=8, ncol=8), dd)
    with pytest.raises(g.GemlabError, match=r"ncol\(K\) must be ≥ jj0 \+ nnode ⋅ space_ndim"):
        integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(jj0=1, nrow=8, ncol=8), dd)
    with pytest.raises(g.GemlabError, match=r"nrow\(K\) must be ≥ ii0 \+ nnode$"):
        integ.mat_01_nsn(ctx, dmesh, None, CommonArgs(ii0=1, nrow=4, ncol=4), Coef.uniform([1.0]))
    with pytest.raises(g.GemlabError, match=r"a.len\(\) must be ≥ ii0 \+ nnode"):
        integ.vec_01_ns(ctx, dmesh, None, CommonArgs(ii0=1, nrow=4), Coef.uniform([1.0]))
    with pytest.raises(g.GemlabError, match="not available for Qua class"):
        integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(ngauss=5), dd)
    with pytest.raises(g.GemlabError, match="cell range is out of bounds"):
        integ.mat_10_bdb(ctx, dmesh, (0, mesh.ncell + 1), CommonArgs(), dd)
    # axisymmetric needs 2D
    m3 = make_mesh("tet4", n=2)
    d3 = ctx.upload(m3)
    with pytest.raises(g.GemlabError, match="axisymmetric requires space_ndim = 2"):
        integ.mat_10_bdb(ctx, d3, None, CommonArgs(axisymmetric=True), Coef.uniform(np.eye(6).ravel()))
    # cable cells: gradient-based cases fail, N-based cases work (src/integ/vec_01_ns.rs works on Lin2 in 2D)
    lin = Mesh.homogeneous(2, np.array([[3.0, 4.0], [3.0 + 6 / np.sqrt(2), 4.0 + 6 / np.sqrt(2)]]), GeoKind.Lin2, [[0, 1]])
    dl = ctx.upload(lin)
    with pytest.raises(g.GemlabError, match="calc_gradient requires that geo_ndim = space_ndim"):
        integ.mat_01_nsn(ctx, dl, None, CommonArgs(), Coef.uniform([1.0]))
    a = integ.vec_01_ns(ctx, dl, None, CommonArgs(), Coef.uniform([2.0]))
    np.testing.assert_allclose(a, [6.0, 6.0], rtol=1e-15)
    # zero determinant: reports the smallest failing cell id
    bad = make_mesh("qua4", n=3)
    c = bad.conn_matrix()
    bad.coords[c[4]] = bad.coords[c[4, 0]]  # collapse cell 4 (also damages its neighbours' geometry, not their det)
    db = ctx.upload(bad)
    with pytest.raises(g.GemlabError, match="cannot compute inverse due to zero determinant") as ei:
        integ.mat_10_bdb(ctx, db, None, CommonArgs(), dd)
    assert ei.value.cell == 4
    # the context stays usable afterwards
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), dd)


def test_check_jacobian_loop(ctx):
    """Mesh::check_jacobian (src/mesh/check.rs:43-60): det J at ksi = [1/3, 1/3, 1/3] for every cell."""
    mesh = make_mesh("tet10", n=3)
    dmesh = ctx.upload(mesh)
    ksi = [1 / 3, 1 / 3, 1 / 3]
    got = integ.jacobian(ctx, dmesh, None, ksi)
    ref = oracle_integ("jacobian", mesh, ksi=ksi)
    np.testing.assert_allclose(got, ref, rtol=1e-13, atol=0)
    assert np.all(got > 0)


def test_cpp_host_mirror(tmp_path):
    """cpp/gemlab.hpp (the compiled-language host mirror) drives the same C ABI: the reference's mat_10_bdb doc-test."""
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "test_batch"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-o", str(exe), str(root / "tests" / "cpp" / "test_batch.cpp"),
                    "-L", str(root / "gemlab_b200"), "-lgemlab_cuda", f"-Wl,-rpath,{root / 'gemlab_b200'}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


def test_cpp_host_mirror_device_paths(tmp_path):
    """DeviceSlab, MultiContext, the device Block, the MSH reader and gl_integ_fused through cpp/gemlab.hpp."""
    import subprocess
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    exe = tmp_path / "test_device_paths"
    subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-o", str(exe), str(root / "tests" / "cpp" / "test_device_paths.cpp"),
                    "-L", str(root / "gemlab_b200"), "-lgemlab_cuda", f"-Wl,-rpath,{root / 'gemlab_b200'}"], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("OK"), r.stdout + r.stderr


@pytest.mark.parametrize("kind", ["tri3", "qua8", "tet10", "hex8"])
def test_fused_scalar_cases_share_one_geometry_pass(ctx, kind):
    """gl_integ_fused: mat_01_nsn + mat_03_btb + vec_01_ns + vec_03_bv from one kernel == the four separate calls
    (BASELINE.json config 4: fused Tet10 mass matrix + body-force vector)."""
    import torch

    mesh = make_mesh(kind, n=4)
    dmesh = ctx.upload(mesh)
    nd = mesh.ndim
    k = GeoKind.from_str(kind)
    rng = np.random.default_rng(29)
    coefs = {"mat_01_nsn": np.array([2.7]), "mat_03_btb": random_coef("mat_03_btb", nd, 1, rng)[0],
             "vec_01_ns": np.array([9.81 * 2.7]), "vec_03_bv": random_coef("vec_03_bv", nd, 1, rng)[0]}
    for axisym in ([False, True] if nd == 2 else [False]):
        args = CommonArgs(axisymmetric=axisym, alpha=1.3)
        outs = {op: torch.full((mesh.ncell * integ.op_out_size(op, k, nd),), 5.0, dtype=torch.float64, device="cuda") for op in coefs}
        launches = ctx.launch_count()
        integ.fused(ctx, dmesh, None, args, [(op, Coef.uniform(c), outs[op]) for op, c in coefs.items()])
        ctx.sync()
        assert ctx.launch_count() - launches == 1  # one kernel for the four cases
        for op, c in coefs.items():
            ref = oracle_integ(op, mesh, c, alpha=1.3, axisymmetric=axisym)
            sizes = np.arange(mesh.ncell + 1) * integ.op_out_size(op, k, nd)
            assert rel_err(outs[op].cpu().numpy(), ref, sizes) <= TOL, (kind, op, axisym)
    # sub-range + a non-fusable item (host output) falls back to sequential calls with identical results
    b, e = 3, mesh.ncell - 2
    host = integ.fused(ctx, dmesh, (b, e), CommonArgs(), [("mat_01_nsn", Coef.uniform(coefs["mat_01_nsn"]), None),
                                                          ("vec_01_ns", Coef.uniform(coefs["vec_01_ns"]), None)])
    np.testing.assert_array_equal(host[0], integ.mat_01_nsn(ctx, dmesh, (b, e), CommonArgs(), Coef.uniform(coefs["mat_01_nsn"])))
    np.testing.assert_array_equal(host[1], integ.vec_01_ns(ctx, dmesh, (b, e), CommonArgs(), Coef.uniform(coefs["vec_01_ns"])))


def test_fused_on_mixed_mesh(ctx):
    import torch

    mesh = mixed_mesh_2d()
    dmesh = ctx.upload(mesh)
    args = CommonArgs(axisymmetric=True)
    tt, w = np.array([2.0, 3.0, 0.0, 0.5]), np.array([1.0, 2.0])
    om = torch.empty(integ.out_total(dmesh, "mat_03_btb", None, args), dtype=torch.float64, device="cuda")
    ov = torch.empty(integ.out_total(dmesh, "vec_03_bv", None, args), dtype=torch.float64, device="cuda")
    integ.fused(ctx, dmesh, None, args, [("mat_03_btb", Coef.uniform(tt), om), ("vec_03_bv", Coef.uniform(w), ov)])
    ctx.sync()
    assert rel_err(om.cpu().numpy(), oracle_integ("mat_03_btb", mesh, tt, axisymmetric=True),
                   integ.out_offsets(dmesh, "mat_03_btb", None, args)) <= TOL
    assert rel_err(ov.cpu().numpy(), oracle_integ("vec_03_bv", mesh, w, axisymmetric=True),
                   integ.out_offsets(dmesh, "vec_03_bv", None, args)) <= TOL


def test_upload_paths_agree_and_validate(ctx):
    """gl_mesh_upload (u64 ids, device-side transposition for single-kind meshes) == gl_mesh_upload_homogeneous (u32)."""
    import ctypes as C

    from gemlab_b200._lib import lib

    mesh = make_mesh("hex8", n=4)
    dd = Coef.uniform(g.mandel_matrix(g.lin_elasticity(480.0, 1 / 3, False)))
    a = integ.mat_10_bdb(ctx, ctx.upload(mesh), None, CommonArgs(), dd)
    h = C.c_void_p()
    conn32 = np.ascontiguousarray(mesh.conn, dtype=np.uint32)
    st = lib().gl_mesh_upload_homogeneous(ctx._h, 3, mesh.npoint, mesh.coords.ctypes.data, mesh.ncell, int(GeoKind.Hex8),
                                          mesh.markers.ctypes.data, conn32.ctypes.data, C.byref(h))
    assert st == 0
    dm = g.DeviceMesh(ctx, h, 3, mesh.ncell, mesh.npoint, mesh.kinds.copy())
    b = integ.mat_10_bdb(ctx, dm, None, CommonArgs(), dd)
    np.testing.assert_array_equal(a, b)
    bad = Mesh(3, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn.copy(), mesh.markers)
    bad.conn[5] = mesh.npoint  # out of range
    with pytest.raises(g.GemlabError, match="point id out of range"):
        ctx.upload(bad)


# ---------------------------------------------------------------------------------------------------------------
# the cubic / quartic kinds (Lin4/5, Tri10/15, Qua12/16/17, Tet20, Hex32): table-driven on both sides
# ---------------------------------------------------------------------------------------------------------------

REF_COORDS = json.loads((Path(__file__).parent / "golden" / "kind_reference_coords.json").read_text())


def mapped_reference_cells(kind, space_ndim, ncell=3, seed=9):
    """A few cells of `kind`: the reference element pushed through a smooth non-affine map, then translated."""
    xi = np.array(REF_COORDS[kind])                      # (nnode, geo_ndim)
    nn, gd = xi.shape
    rng = np.random.default_rng(seed)
    a = np.eye(space_ndim, gd) + 0.15 * rng.standard_normal((space_ndim, gd))
    coords, conn = [], []
    for c in range(ncell):
        bend = 0.04 * rng.standard_normal((space_ndim, gd))
        x = xi @ a.T + (xi ** 2) @ bend.T + np.array([3.0 * c + 2.0] + [0.5] * (space_ndim - 1))
        conn.append(np.arange(nn) + c * nn)
        coords.append(x)
    return Mesh.homogeneous(space_ndim, np.concatenate(coords), GeoKind.from_str(kind), np.array(conn))


@pytest.mark.parametrize("kind,space_ndim", [("tri10", 2), ("tri15", 2), ("qua12", 2), ("qua16", 2), ("qua17", 2), ("tet20", 3),
                                              ("hex32", 3), ("lin4", 2), ("lin5", 3)])
def test_higher_order_kinds(ctx, kind, space_ndim):
    mesh = mapped_reference_cells(kind, space_ndim)
    dmesh = ctx.upload(mesh)
    solid = GeoKind.from_str(kind).ndim() == space_ndim
    ops = OPS if solid else ["vec_01_ns", "vec_02_nv"]   # the other cases call calc_gradient: geo_ndim == space_ndim only
    rng = np.random.default_rng(13)
    for op in ops:
        coef = random_coef(op, space_ndim, 1, rng)[0]
        for axisym in ([False, True] if (space_ndim == 2 and solid) else [False]):
            args = CommonArgs(alpha=1.3, axisymmetric=axisym)
            got = getattr(integ, op)(ctx, dmesh, None, args, Coef.uniform(coef))
            ref = oracle_integ(op, mesh, coef, alpha=1.3, axisymmetric=axisym)
            off = integ.out_offsets(dmesh, op, None, args)
            assert off[-1] == ref.size
            assert rel_err(got, ref, off) <= TOL, f"{kind} {op} axisym={axisym}"
    # the default rules are the reference's (Gauss::new, src/integ/gauss.rs:87-121); a sized rule works as well
    if solid:
        n = {"tri10": 7, "tri15": 12, "qua12": 9, "qua16": 9, "qua17": 9, "tet20": 14, "hex32": 27}[kind]
        coef = random_coef("mat_10_bdb", space_ndim, 1, rng)[0]
        got = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(ngauss=n), Coef.uniform(coef))
        ref = oracle_integ("mat_10_bdb", mesh, coef, ngauss=n)
        assert rel_err(got, ref, integ.out_offsets(dmesh, "mat_10_bdb", None, CommonArgs(ngauss=n))) <= TOL
    # Mesh::check_jacobian's loop (src/mesh/check.rs:43-60) on these kinds too: det J at an arbitrary point comes from the nodal
    # basis of gl_shapes.cpp; the oracle knows these kinds at the reference's Gauss points, so one of those is the probe
    if solid:
        gs = po.Gauss(kind, 0)
        ksi = list(gs.coords(gs.npoint() // 2))[:space_ndim]
        det = integ.jacobian(ctx, dmesh, None, ksi)
        ref = oracle_integ("jacobian", mesh, ksi=np.array(ksi))
        assert np.max(np.abs(det - ref)) <= 1e-12 * np.max(np.abs(ref))
    # volume check independent of the oracle: sum of vec_01_ns with s = 1 is the cell measure; refine the rule
    if solid:
        v1 = integ.vec_01_ns(ctx, dmesh, None, CommonArgs(), Coef.uniform([1.0])).sum()
        hi = {"tri10": 16, "tri15": 16, "qua12": 16, "qua16": 16, "qua17": 16, "tet20": 24, "hex32": 64}[kind]
        v2 = integ.vec_01_ns(ctx, dmesh, None, CommonArgs(ngauss=hi), Coef.uniform([1.0])).sum()
        assert abs(v1 - v2) <= 1e-9 * abs(v2)


# ---------------------------------------------------------------------------------------------------------------
# out-of-bounds discipline: guard bands around device outputs must survive every kernel family
# ---------------------------------------------------------------------------------------------------------------

@pytest.mark.parametrize("guard", [2048, 2049])  # 2049: the output is only 8-byte aligned -> the tuned kernels must decline
@pytest.mark.parametrize("kind", ["tri3", "qua4", "qua8", "tet10", "hex8", "hex20"])
def test_kernels_write_only_their_own_blocks(ctx, kind, guard):
    """compute-sanitizer is not available on the GPU pool, so stray writes are caught the plain way: every output sits between
    two guard bands of a sentinel value, ranges are ragged (not multiples of any tile size), and the bands must be intact."""
    import torch

    mesh = make_mesh(kind, n=5)
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(31)
    sentinel = -7.25
    b, e = 3, mesh.ncell - 5  # ragged sub-range
    for op in ["mat_10_bdb", "mat_01_nsn", "mat_03_btb", "vec_01_ns", "vec_03_bv", "vec_04_bt"]:
        coef = random_coef(op, mesh.ndim, 1, rng)[0]
        total = integ.out_total(dmesh, op, (b, e), CommonArgs())
        buf = torch.full((total + 2 * guard,), sentinel, dtype=torch.float64, device="cuda")
        getattr(integ, op)(ctx, dmesh, (b, e), CommonArgs(), Coef.uniform(coef), out=buf[guard:guard + total])
        ctx.sync()
        assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + total:] == sentinel).all()), f"{kind} {op}"
        ref = oracle_integ(op, mesh, coef)
        bs = integ.op_out_size(op, GeoKind.from_str(kind), mesh.ndim)
        assert rel_err(buf[guard:guard + total].cpu().numpy(), ref[b * bs:e * bs], np.arange(e - b + 1) * bs) <= TOL, f"{kind} {op}"
    # the fused entry point (mass + source, conductivity + flux) with guard bands as well
    for pair in ([("mat_01_nsn", [2.7]), ("vec_01_ns", [9.81])], [("mat_03_btb", [2.0] * mesh.ndim + [0.0] * mesh.ndim), ("vec_03_bv", [1.0] * mesh.ndim)]):
        bufs = []
        for op, _ in pair:
            total = integ.out_total(dmesh, op, (b, e), CommonArgs())
            bufs.append((torch.full((total + 2 * guard,), sentinel, dtype=torch.float64, device="cuda"), total))
        integ.fused(ctx, dmesh, (b, e), CommonArgs(), [(op, Coef.uniform(c), buf[guard:guard + total]) for (op, c), (buf, total) in zip(pair, bufs)])
        ctx.sync()
        for (op, c), (buf, total) in zip(pair, bufs):
            assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + total:] == sentinel).all()), f"{kind} fused {op}"
            ref = oracle_integ(op, mesh, np.array(c))
            bs = integ.op_out_size(op, GeoKind.from_str(kind), mesh.ndim)
            assert rel_err(buf[guard:guard + total].cpu().numpy(), ref[b * bs:e * bs], np.arange(e - b + 1) * bs) <= TOL, f"{kind} fused {op}"


def test_fused_stiffness_and_internal_force_hex8(ctx):
    """gl_integ_fused({mat_10_bdb, vec_04_bt}) on a Hex8 mesh: K and d = int B.sigma from one geometry pass (sigma per Gauss
    point), against the two separate oracle loops; also with one modulus per marker, on a sub-range, and with host sigma."""
    import torch

    mesh = make_mesh("hex8", n=5)
    ncell = mesh.ncell
    rng = np.random.default_rng(41)
    mesh.markers[:] = rng.integers(1, 3, ncell)
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, False))
    sig = rng.standard_normal((ncell * 8, 6))
    ref_k = oracle_integ("mat_10_bdb", mesh, dd)
    ref_d = oracle_integ("vec_04_bt", mesh, sig, cell_stride=48, gp_stride=6)
    launches0 = ctx.launch_count()
    out_k = torch.empty(ncell * 576, dtype=torch.float64, device="cuda")
    out_d = torch.empty(ncell * 24, dtype=torch.float64, device="cuda")
    integ.fused(ctx, dmesh, None, CommonArgs(), [("mat_10_bdb", Coef.uniform(dd), out_k),
                                                 ("vec_04_bt", Coef.per_gauss(torch.from_numpy(sig).cuda()), out_d)])
    ctx.sync()
    assert ctx.launch_count() - launches0 == 1  # one kernel produced both
    assert rel_err(out_k.cpu().numpy(), ref_k, np.arange(ncell + 1) * 576) <= TOL
    assert rel_err(out_d.cpu().numpy(), ref_d, np.arange(ncell + 1) * 24) <= TOL
    # sub-range, host-resident sigma records, one modulus per marker
    b, e = 7, ncell - 3
    recs = np.stack([dd, 2.5 * dd])
    per_cell = recs[mesh.markers - 1]
    ref_k2 = oracle_integ("mat_10_bdb", mesh, per_cell, cell_stride=36)
    out_k.fill_(0.0)
    out_d.fill_(0.0)
    integ.fused(ctx, dmesh, (b, e), CommonArgs(alpha=0.5), [("mat_10_bdb", Coef.per_marker([1, 2], recs), out_k[: (e - b) * 576]),
                                                            ("vec_04_bt", Coef.per_gauss(sig[b * 8:e * 8]), out_d[: (e - b) * 24])])
    ctx.sync()
    assert rel_err(out_k[: (e - b) * 576].cpu().numpy(), 0.5 * ref_k2[b * 576:e * 576], np.arange(e - b + 1) * 576) <= TOL
    assert rel_err(out_d[: (e - b) * 24].cpu().numpy(), 0.5 * ref_d[b * 24:e * 24], np.arange(e - b + 1) * 24) <= TOL
    # a rule the tuned kernel does not cover: the pair still works (run one after the other)
    integ.fused(ctx, dmesh, None, CommonArgs(ngauss=27), [("mat_10_bdb", Coef.uniform(dd), out_k),
                                                           ("vec_04_bt", Coef.per_gauss(np.tile(sig[:1], (ncell * 27, 1))), out_d)])
    ctx.sync()
    ref_k27 = oracle_integ("mat_10_bdb", mesh, dd, ngauss=27)
    assert rel_err(out_k.cpu().numpy(), ref_k27, np.arange(ncell + 1) * 576) <= TOL


@pytest.mark.parametrize("guard", [1024, 1025])
@pytest.mark.parametrize("kind", ["tri3", "qua4", "qua8", "qua9", "tet4", "tet10", "hex8", "hex20"])
def test_round2_kernels_write_only_their_own_blocks(ctx, kind, guard):
    """Guard bands around the outputs of the round-2 paths: one modulus per Gauss point (hex8_gauss / bdb<..,gauss>), the general
    store epilogue (nrow / ncol / ii0 / jj0, clear = false) and the planned assembly (vals and nothing else is touched)."""
    import scipy.sparse as sp
    import torch

    mesh = make_mesh(kind, n=4)
    dmesh = ctx.upload(mesh)
    gk = GeoKind.from_str(kind)
    ng = g.Gauss.new(gk).npoint()
    nd = gk.nnode() * mesh.ndim
    rng = np.random.default_rng(8)
    sentinel = -3.5
    b, e = 2, mesh.ncell - 3
    recs = random_coef("mat_10_bdb", mesh.ndim, (e - b) * ng, rng)
    ns2 = (2 * mesh.ndim) ** 2
    sub = mesh.permuted(np.arange(b, e))
    ref = oracle_integ("mat_10_bdb", sub, recs, cell_stride=ng * ns2, gp_stride=ns2).reshape(e - b, nd, nd)
    # 1. per-Gauss moduli, tight blocks
    total = (e - b) * nd * nd
    buf = torch.full((total + 2 * guard,), sentinel, dtype=torch.float64, device="cuda")
    integ.mat_10_bdb(ctx, dmesh, (b, e), CommonArgs(), Coef.per_gauss(torch.from_numpy(recs).cuda()), out=buf[guard:guard + total])
    ctx.sync()
    assert "gauss" in ctx.last_kernel()
    assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + total:] == sentinel).all()), kind
    got = buf[guard:guard + total].cpu().numpy().reshape(e - b, nd, nd)
    assert rel_err(got.ravel(), ref.ravel(), np.arange(e - b + 1) * nd * nd) <= TOL
    # 2. per-Gauss moduli accumulated into an offset window of larger blocks
    nrow, ncol = nd + 3, nd + 4
    total = (e - b) * nrow * ncol
    buf = torch.full((total + 2 * guard,), sentinel, dtype=torch.float64, device="cuda")
    buf[guard:guard + total] = 1.0
    integ.mat_10_bdb(ctx, dmesh, (b, e), CommonArgs(ii0=2, jj0=1, nrow=nrow, ncol=ncol, clear=False),
                     Coef.per_gauss(torch.from_numpy(recs).cuda()), out=buf[guard:guard + total])
    ctx.sync()
    assert ctx.last_kernel().endswith(",gen>"), ctx.last_kernel()
    assert bool((buf[:guard] == sentinel).all()) and bool((buf[guard + total:] == sentinel).all()), kind
    blk = buf[guard:guard + total].cpu().numpy().reshape(e - b, ncol, nrow)
    inner = blk[:, 1:1 + nd, 2:2 + nd]
    np.testing.assert_allclose(inner, ref + 1.0, rtol=1e-12, atol=1e-12 * np.abs(ref).max())  # both are [cell][column][row]
    rest = blk.copy()
    rest[:, 1:1 + nd, 2:2 + nd] = 1.0
    assert np.all(rest == 1.0)
    # 3. planned assembly: only `vals` changes
    eq = np.arange(mesh.npoint * mesh.ndim, dtype=np.int32)
    rows, cols = integ.triplet_indices(ctx, dmesh, "mat_10_bdb", (b, e), CommonArgs(), eq)
    rr, cc = rows.cpu().numpy(), cols.cpu().numpy()
    pat = sp.coo_matrix((np.ones(rr.size), (rr, cc)), shape=(eq.size, eq.size)).tocsr()
    pat.sum_duplicates()
    pat.sort_indices()
    rowptr = torch.tensor(pat.indptr.astype(np.int64), device="cuda")
    colidx = torch.tensor(pat.indices.astype(np.int32), device="cuda")
    vbuf = torch.full((pat.nnz + 2 * guard,), sentinel, dtype=torch.float64, device="cuda")
    vbuf[guard:guard + pat.nnz] = 0.0
    plan = integ.AssemblyPlan(ctx, dmesh, "mat_10_bdb", (b, e), CommonArgs(), eq, rowptr, colidx)
    plan.assemble(Coef.per_gauss(torch.from_numpy(recs).cuda()), vbuf[guard:guard + pat.nnz])
    ctx.sync()
    assert bool((vbuf[:guard] == sentinel).all()) and bool((vbuf[guard + pat.nnz:] == sentinel).all()), kind
    # column-major blocks: entry q = ii + jj * nd of element e is ref[e, jj, ii]
    want = sp.coo_matrix((ref.ravel(), (rr, cc)), shape=(eq.size, eq.size)).tocsr()
    want.sum_duplicates()
    want.sort_indices()
    assert np.max(np.abs(vbuf[guard:guard + pat.nnz].cpu().numpy() - want.data)) <= 2e-12 * np.abs(want.data).max()
    plan.free()

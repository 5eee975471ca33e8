"""GPU parity of the two-pad coupling matrices (integ::mat_04_nsb .. mat_07_bsn), the boundary integrals (*_bry) and the batched
calc_normal_vector: the reference's known answers replayed through the C ABI, then distorted meshes against the oracle."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "golden"))
import coupled_and_boundary as gold  # noqa: E402

import gemlab_b200 as g  # noqa: E402
from gemlab_b200 import Coef, CommonArgs, GeoKind, Mesh, integ  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

from test_gpu_parity import make_mesh, rel_err  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


COUPLED = gold.coupled_cases()
BOUNDARY = gold.boundary_cases()


@pytest.mark.parametrize("case", COUPLED, ids=[c[0] for c in COUPLED])
def test_coupled_known_answers_on_gpu(ctx, case):
    name, ref, op, kind, kind_b, ndim, coords, rec, expect, rules, tol = case
    k = GeoKind.from_str(kind)
    dm = ctx.upload(Mesh.homogeneous(ndim, np.asarray(coords, dtype=float), k, np.arange(k.nnode()).reshape(1, -1)))
    for n in rules:
        out = np.full(expect.size, 1234.56)
        getattr(integ, op)(ctx, dm, None, CommonArgs(ngauss=n, kind_b=GeoKind.from_str(kind_b)), Coef.uniform(rec), out=out)
        assert np.max(np.abs(out.reshape(expect.shape, order="F") - expect)) <= tol, f"{name} rule {n} ({ref})"


@pytest.mark.parametrize("case", BOUNDARY, ids=[c[0] for c in BOUNDARY])
def test_boundary_known_answers_on_gpu(ctx, case):
    name, ref, op, kind, ndim, coords, rec, expect, rule, tol = case
    k = GeoKind.from_str(kind)
    dm = ctx.upload(Mesh.homogeneous(ndim, np.asarray(coords, dtype=float), k, np.arange(k.nnode()).reshape(1, -1)))
    out = np.full(expect.size, 1234.56)
    getattr(integ, op)(ctx, dm, None, CommonArgs(ngauss=None if rule == 0 else rule), Coef.uniform(rec), out=out)
    got = out.reshape(expect.shape, order="F") if expect.ndim == 2 else out
    assert np.max(np.abs(got - expect)) <= tol, f"{name} ({ref})"


def oracle_ext(op, mesh, kind_b=None, coef=None, **kw):
    return po.mesh_integ_ext(op, mesh.ndim, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn,
                             kind_b=None if kind_b is None else int(kind_b), coef=coef, **kw)


@pytest.mark.parametrize("kind,kind_b", [("qua8", "qua4"), ("tri6", "tri3"), ("qua9", "qua4"), ("hex20", "hex8"), ("tet10", "tet4"),
                                         ("qua8", "qua8"), ("hex8", "hex8")])
@pytest.mark.parametrize("op", ["mat_04_nsb", "mat_05_btn", "mat_06_nvn", "mat_07_bsn"])
def test_coupled_on_distorted_meshes(ctx, kind, kind_b, op):
    mesh = make_mesh(kind, n=3)
    dm = ctx.upload(mesh)
    kb = GeoKind.from_str(kind_b)
    rng = np.random.default_rng(13)
    nrec = {"mat_04_nsb": 1, "mat_05_btn": 2 * mesh.ndim, "mat_06_nvn": mesh.ndim, "mat_07_bsn": 1}[op]
    rec = rng.standard_normal(nrec) + 2.0
    for axisym in ([False, True] if mesh.ndim == 2 else [False]):
        args = CommonArgs(alpha=0.7, axisymmetric=axisym, kind_b=kb)
        got = getattr(integ, op)(ctx, dm, None, args, Coef.uniform(rec))
        assert ctx.last_kernel().startswith("generic<")
        ref = oracle_ext(op, mesh, kb, rec, alpha=0.7, axisymmetric=axisym)
        off = integ.out_offsets(dm, op, None, args)
        assert off[-1] == ref.size
        assert rel_err(got, ref, off) <= TOL, f"{kind}+{kind_b} {op} axisym={axisym}"
    # one record per Gauss point, sub-range
    ng = g.Gauss.new(GeoKind.from_str(kind)).npoint()
    recs = rng.standard_normal((mesh.ncell * ng, nrec)) + 2.0
    b, e = 1, mesh.ncell - 2
    got = getattr(integ, op)(ctx, dm, (b, e), CommonArgs(kind_b=kb), Coef.per_gauss(recs[b * ng:e * ng]))
    ref = oracle_ext(op, mesh, kb, recs, coef_cell_stride=ng * nrec, coef_gp_stride=nrec)
    bs = ref.size // mesh.ncell
    assert rel_err(got, ref[b * bs:e * bs], np.arange(e - b + 1) * bs) <= TOL


def boundary_mesh(kind, seed=3):
    """boundary cells of a distorted solid mesh: the edges (2D) / faces (3D) of its cells lying on x = xmin"""
    rng = np.random.default_rng(seed)
    if kind in ("lin2", "lin3"):
        n = 7
        t = np.linspace(0.0, 1.0, n + 1)
        pts = np.stack([1.0 + 0.3 * np.sin(2 * t) + 0.02 * rng.standard_normal(n + 1), 2.0 * t], axis=1)
        if kind == "lin2":
            return Mesh.homogeneous(2, pts, GeoKind.Lin2, np.stack([np.arange(n), np.arange(1, n + 1)], axis=1))
        mid = 0.5 * (pts[:-1] + pts[1:]) + 0.01 * rng.standard_normal((n, 2))
        allp = np.concatenate([pts, mid])
        return Mesh.homogeneous(2, allp, GeoKind.Lin3, np.stack([np.arange(n), np.arange(1, n + 1), n + 1 + np.arange(n)], axis=1))
    solid = make_mesh({"qua4": "hex8", "qua8": "hex20", "tri3": "tet4", "tri6": "tet10"}[kind], n=3)
    c = solid.conn_matrix().astype(np.int64)
    local = {"qua4": [0, 3, 7, 4], "qua8": [0, 3, 7, 4, 11, 19, 15, 16], "tri3": [0, 1, 2], "tri6": [0, 1, 2, 4, 5, 6]}[kind]
    faces = c[:, local]
    return Mesh.homogeneous(3, solid.coords, GeoKind.from_str(kind), faces)


@pytest.mark.parametrize("kind", ["lin2", "lin3", "qua4", "qua8", "tri3", "tri6"])
def test_boundary_integrals_and_normals(ctx, kind):
    mesh = boundary_mesh(kind)
    dm = ctx.upload(mesh)
    nd = mesh.ndim
    rng = np.random.default_rng(19)
    args = CommonArgs(alpha=1.3)
    nrm = integ.normal_vectors(ctx, dm, None, args)
    ref = oracle_ext("normal", mesh)
    assert np.max(np.abs(nrm - ref)) <= 1e-13 * max(1.0, np.max(np.abs(ref)))
    for op in ("mat_01_nsn_bry", "vec_01_ns_bry", "vec_02_nv_bry"):
        rec = rng.standard_normal(nd + 1) + 1.5
        for axisym in ([False, True] if nd == 2 else [False]):
            a = CommonArgs(alpha=1.3, axisymmetric=axisym)
            got = getattr(integ, op)(ctx, dm, None, a, Coef.uniform(rec))
            ref = oracle_ext(op, mesh, None, rec, alpha=1.3, axisymmetric=axisym)
            off = integ.out_offsets(dm, op, None, a)
            assert rel_err(got, ref, off) <= TOL, f"{kind} {op} axisym={axisym}"


def test_error_strings_of_the_new_cases(ctx):
    tri = make_mesh("tri3", n=2)
    dt = ctx.upload(tri)
    with pytest.raises(g.GemlabError, match=r"in 2D, geometry ndim must be equal to 1 \(a line\)"):
        integ.mat_01_nsn_bry(ctx, dt, None, CommonArgs(), Coef.uniform([1.0, 0.0, 0.0]))
    with pytest.raises(g.GemlabError, match=r"calc_normal_vector requires geo_ndim = 1 in 2D"):
        integ.normal_vectors(ctx, dt, None, CommonArgs())
    tet = make_mesh("tet4", n=2)
    d3 = ctx.upload(tet)
    with pytest.raises(g.GemlabError, match=r"in 3D, geometry ndim must be equal to 2 \(a surface\)"):
        integ.vec_02_nv_bry(ctx, d3, None, CommonArgs(), Coef.uniform([0.0, 0.0, 0.0, 1.0]))
    q8 = make_mesh("qua8", n=2)
    dq = ctx.upload(q8)
    with pytest.raises(g.GemlabError, match=r"nrow\(K\) must be ≥ ii0 \+ pad_b.nnode"):
        integ.mat_04_nsb(ctx, dq, None, CommonArgs(kind_b=GeoKind.Qua4, ii0=1, nrow=4, ncol=16), Coef.uniform([1.0]))
    with pytest.raises(g.GemlabError, match=r"ncol\(K\) must be ≥ jj0 \+ pad.nnode ⋅ space_ndim"):
        integ.mat_04_nsb(ctx, dq, None, CommonArgs(kind_b=GeoKind.Qua4, jj0=1, nrow=4, ncol=16), Coef.uniform([1.0]))
    with pytest.raises(g.GemlabError, match=r"nrow\(K\) must be ≥ ii0 \+ pad.nnode ⋅ space_ndim"):
        integ.mat_06_nvn(ctx, dq, None, CommonArgs(kind_b=GeoKind.Qua4, ii0=1, nrow=16, ncol=4), Coef.uniform([1.0, 2.0]))
    with pytest.raises(g.GemlabError, match=r"ncol\(K\) must be ≥ jj0 \+ pad_b.nnode"):
        integ.mat_07_bsn(ctx, dq, None, CommonArgs(kind_b=GeoKind.Qua4, jj0=1, nrow=16, ncol=4), Coef.uniform([1.0]))

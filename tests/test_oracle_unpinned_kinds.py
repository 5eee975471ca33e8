"""Cross-checks of the oracle for the kinds whose element matrices NO reference test pins (Hex8, Hex20, Tet10 and
Qua8 stiffness; only their N, L, J and B are pinned by the reference's own tests).  Four independent lines of evidence:

 1. integ::mat_10_bdb against integ::mat_10_bdb_alt (the B-matrix form, src/integ/mat_10_bdb.rs:304-354) -- the reference's own
    cross-check pattern (mat_10_bdb.rs:493-499, 544, 626) -- on distorted cells;
 2. an axis-aligned Hex8 brick against the CLOSED FORM of its stiffness (products of linear shape functions integrate exactly:
    no quadrature involved);
 3. invariants: K symmetric for a symmetric D, K.(rigid translations and rotations) = 0, sum(mass) = rho.volume,
    sum(a) = s.volume;
 4. rule refinement: on affine cells the integrand is a polynomial the default rule already integrates exactly, so every finer
    rule must give the same K (8 -> 27 -> 64 points on Hex8, ...).
"""
import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Block, GeoKind
from oracle import pyoracle as po


def integ(op, mesh, coef, **kw):
    return po.mesh_integ(op, mesh.ndim, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, coef=coef, **kw)


def distorted(kind, seed=4):
    if kind == "hex8":
        return g.jitter_interior_points(Block.new_cube(1.5).set_ndiv([2, 3, 2]).subdivide(GeoKind.Hex8), 0.08, seed)
    if kind == "hex20":
        return g.jitter_interior_points(Block.new_cube(1.5).set_ndiv([2, 2, 2]).subdivide(GeoKind.Hex20), 0.04, seed)
    if kind == "tet10":
        return g.jitter_interior_points(g.split_hex_cells_into_tet10(2, 2, 1, seed=seed), 0.02, seed)
    if kind == "qua8":
        m = Block.new_square(2.0).set_ndiv([3, 2]).subdivide(GeoKind.Qua8)
        m.coords[:, 0] += 1.0
        return g.jitter_interior_points(m, 0.05, seed)
    raise AssertionError(kind)


def affine(kind):
    """cells that are affine images of the reference cell: constant Jacobian"""
    a = np.array([[1.0, 0.3, 0.1], [0.2, 0.8, -0.2], [-0.1, 0.25, 1.2]])
    if kind == "qua8":
        m = Block.new_square(1.0).set_ndiv([2, 2]).subdivide(GeoKind.Qua8)
        m.coords[:] = m.coords @ a[:2, :2].T + np.array([3.0, 1.0])
        return m
    if kind == "tet10":
        m = g.split_hex_cells_into_tet10(1, 1, 1)
    else:
        m = Block.new_cube(1.0).set_ndiv([2, 1, 2]).subdivide(GeoKind.from_str(kind))
    m.coords[:] = m.coords @ a.T + np.array([3.0, 1.0, -2.0])
    return m


def modulus(ndim, rng=None):
    if rng is None:
        return g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, ndim == 2))
    ns = 2 * ndim
    a = rng.standard_normal((ns, ns))
    return (a @ a.T + ns * np.eye(ns)).ravel(order="F")


KINDS = ["hex8", "hex20", "tet10", "qua8"]


@pytest.mark.parametrize("kind", KINDS)
def test_bdb_agrees_with_the_b_matrix_form(kind):
    mesh = distorted(kind)
    rng = np.random.default_rng(8)
    for dd in (modulus(mesh.ndim), modulus(mesh.ndim, rng)):
        for axisym in ([False, True] if mesh.ndim == 2 else [False]):
            k1 = integ("mat_10_bdb", mesh, dd, axisymmetric=axisym)
            k2 = integ("mat_10_bdb_alt", mesh, dd, axisymmetric=axisym)
            assert np.max(np.abs(k1 - k2)) <= 1e-12 * np.max(np.abs(k1)), kind


def brick_stiffness_closed_form(a, b, c, dd):
    """Stiffness of the Hex8 brick [0,a]x[0,b]x[0,c] with Mandel modulus dd (6 x 6): every entry is a sum of products of three 1D
    integrals of linear functions over [-1, 1] -- int (1+pz)(1+qz) = 2 + 2pq/3, int p (1+qz) = 2p, int p q = 2pq."""
    ref = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
    h = np.array([a, b, c])
    vol = a * b * c / 8.0  # det J
    # tensor D_ikjl from the Mandel matrix
    idx = {(0, 0): 0, (1, 1): 1, (2, 2): 2, (0, 1): 3, (1, 0): 3, (1, 2): 4, (2, 1): 4, (0, 2): 5, (2, 0): 5}
    d4 = np.zeros((3, 3, 3, 3))
    for i in range(3):
        for k in range(3):
            for j in range(3):
                for l in range(3):
                    f = (1.0 if i == k else np.sqrt(2.0)) * (1.0 if j == l else np.sqrt(2.0))
                    d4[i, k, j, l] = dd[idx[(i, k)], idx[(j, l)]] / f
    kk = np.zeros((24, 24))
    for m in range(8):
        for n in range(8):
            # G[k][l] = int dN_m/dx_k dN_n/dx_l dV
            gg = np.zeros((3, 3))
            for k in range(3):
                for l in range(3):
                    val = 1.0
                    for ax in range(3):
                        p, q = ref[m, ax], ref[n, ax]
                        dm, dn = (ax == k), (ax == l)
                        if dm and dn:
                            one = 2.0 * p * q
                        elif dm:
                            one = 2.0 * p
                        elif dn:
                            one = 2.0 * q
                        else:
                            one = 2.0 + 2.0 * p * q / 3.0
                        val *= one
                    gg[k, l] = val / 64.0 * (2.0 / h[k]) * (2.0 / h[l]) * vol
            for i in range(3):
                for j in range(3):
                    kk[3 * m + i, 3 * n + j] = np.sum(d4[i, :, j, :] * gg)
    return kk


def test_hex8_brick_against_the_closed_form():
    a, b, c = 0.7, 1.3, 0.45
    coords = np.array([[0, 0, 0], [a, 0, 0], [a, b, 0], [0, b, 0], [0, 0, c], [a, 0, c], [a, b, c], [0, b, c]], dtype=float)
    mesh = g.Mesh.homogeneous(3, coords, GeoKind.Hex8, [[0, 1, 2, 3, 4, 5, 6, 7]])
    rng = np.random.default_rng(2)
    for dd in (modulus(3), modulus(3, rng)):
        k = integ("mat_10_bdb", mesh, dd).reshape(24, 24, order="F")
        ref = brick_stiffness_closed_form(a, b, c, dd.reshape(6, 6, order="F"))
        assert np.max(np.abs(k - ref)) <= 1e-13 * np.max(np.abs(ref))


def rigid_modes(coords, ndim):
    n = coords.shape[0]
    modes = []
    for i in range(ndim):
        t = np.zeros((n, ndim))
        t[:, i] = 1.0
        modes.append(t.ravel())
    rots = [(0, 1)] if ndim == 2 else [(0, 1), (1, 2), (0, 2)]
    for i, j in rots:  # linearised rotation in the (i, j) plane
        r = np.zeros((n, ndim))
        r[:, i] = -coords[:, j]
        r[:, j] = coords[:, i]
        modes.append(r.ravel())
    return modes


@pytest.mark.parametrize("kind", KINDS)
def test_invariants(kind):
    mesh = distorted(kind)
    nd = mesh.ndim
    kindv = GeoKind.from_str(kind)
    nn = kindv.nnode()
    ndof = nn * nd
    rng = np.random.default_rng(5)
    dd = modulus(nd, rng)
    kk = integ("mat_10_bdb", mesh, dd).reshape(mesh.ncell, ndof, ndof)
    scale = np.max(np.abs(kk))
    assert np.max(np.abs(kk - np.transpose(kk, (0, 2, 1)))) <= 1e-13 * scale  # symmetric D -> symmetric K
    conn = mesh.conn_matrix().astype(np.int64)
    for e in range(mesh.ncell):
        ke = kk[e].T  # stored column-major: [col][row] -> transpose gives K
        xe = mesh.coords[conn[e]]
        for mode in rigid_modes(xe, nd):
            assert np.max(np.abs(ke @ mode)) <= 1e-11 * scale * max(1.0, np.max(np.abs(mode))), (kind, e)
    # mass and source sums: int N_m s N_n summed over m, n = s.volume (partition of unity); same for int N_m s
    rho, s = 2.7, 9.81
    mm = integ("mat_01_nsn", mesh, np.array([rho])).reshape(mesh.ncell, nn * nn).sum(axis=1)
    aa = integ("vec_01_ns", mesh, np.array([s])).reshape(mesh.ncell, nn).sum(axis=1)
    np.testing.assert_allclose(mm / rho, aa / s, rtol=1e-12)
    total = (aa / s).sum()
    lo, hi = mesh.coords.min(axis=0), mesh.coords.max(axis=0)
    np.testing.assert_allclose(total, np.prod(hi - lo), rtol=1e-12)  # the jitter keeps the boundary: the cells tile the box


RULES = {"hex8": [8, 14, 27, 64], "hex20": [27, 64], "tet10": [4, 5, 8, 14, 15, 24], "qua8": [9, 16]}


@pytest.mark.parametrize("kind", KINDS)
def test_rule_refinement_on_affine_cells(kind):
    """Constant Jacobian: the stiffness integrand is a polynomial of the degree the coarsest listed rule already integrates
    exactly, so every rule gives the same matrix."""
    mesh = affine(kind)
    dd = modulus(mesh.ndim, np.random.default_rng(12))
    ref = integ("mat_10_bdb", mesh, dd, ngauss=RULES[kind][0])
    for n in RULES[kind][1:]:
        k = integ("mat_10_bdb", mesh, dd, ngauss=n)
        assert np.max(np.abs(k - ref)) <= 2e-13 * np.max(np.abs(ref)), f"{kind}: rule {n} vs {RULES[kind][0]}"

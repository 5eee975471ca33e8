"""recovery::get_extrap_matrix (src/recovery/get_extrap_matrix.rs:129-200) and GeoKind::reference_coords through the C ABI
(host-only entry points: no device needed).  The expected matrices restate the reference's algorithm with numpy's SVD-based
pseudo-inverse (the reference uses russell_lab::mat_pseudo_inverse); the reference's own tests are replayed as well."""
import numpy as np
import pytest

from gemlab_b200 import GeoKind, lib
from gemlab_b200.geo import Gauss

RULES = {"lin": [1, 2, 3, 4, 5], "tri": [1, 3, 4, 6, 7, 12, 16], "qua": [1, 4, 9, 16], "tet": [1, 4, 5, 8, 14, 15, 24],
         "hex": [6, 8, 14, 27, 64]}


def ref_coords(kind):
    out = np.zeros((kind.nnode(), kind.ndim()))
    for m in range(kind.nnode()):
        k = np.zeros(3)
        assert lib().gl_kind_reference_coords(int(kind), m, k.ctypes.data) == 0
        out[m] = k[: kind.ndim()]
    return out


def interp_matrix(kind, gauss):  # get_interp_matrix.rs: P[p][m] = N_m(ksi_p)
    return np.array([kind.interp_deriv(gauss.coords(p)[: kind.ndim()])[0] for p in range(gauss.npoint())])


def extrap_expected(kind, gauss):
    pp = interp_matrix(kind, gauss)
    npt, nv = pp.shape
    gd = kind.ndim()
    if npt == nv:
        return np.linalg.inv(pp)
    if npt > nv or npt == 1:
        return np.linalg.pinv(pp)
    hx = np.array([list(gauss.coords(p)[:gd]) + [1.0] for p in range(npt)])
    bb = np.linalg.pinv(hx)
    aa = np.eye(npt) - hx @ bb
    xi = np.concatenate([ref_coords(kind), np.ones((nv, 1))], axis=1)
    return xi @ bb + np.linalg.pinv(pp) @ aa


def extrap(kind, ng):
    g = Gauss.new_sized(kind.geo_class(), ng)
    ee = np.zeros(kind.nnode() * g.npoint())
    st = lib().gl_extrap_matrix(int(kind), ng, ee.ctypes.data)
    assert st == 0, st
    return ee.reshape(kind.nnode(), g.npoint(), order="F"), g


@pytest.mark.parametrize("kind", list(GeoKind), ids=[k.to_string() for k in GeoKind])
def test_extrap_matrix_all_kinds_and_rules(kind):
    cls = kind.geo_class().name.lower()
    for ng in RULES[cls]:
        ee, g = extrap(kind, ng)
        ref = extrap_expected(kind, g)
        assert np.max(np.abs(ee - ref)) <= 1e-10 * max(1.0, np.max(np.abs(ref))), f"{kind.to_string()} rule {ng}"


def test_reference_tests_of_get_extrap_matrix():
    """src/recovery/get_extrap_matrix.rs:265-345: Qua4 with 4 and 1 points, Qua8 with 9 points -- E does not depend on the cell"""
    u4 = np.array([1.0, 2.0, 3.0, 4.0])
    ee, g = extrap(GeoKind.Qua4, 4)
    np.testing.assert_allclose(ee @ (interp_matrix(GeoKind.Qua4, g) @ u4), u4, atol=1e-14)
    ee, g = extrap(GeoKind.Qua4, 1)
    np.testing.assert_allclose(ee, np.ones((4, 1)), atol=1e-14)  # the pseudo-inverse of P = [1/4, 1/4, 1/4, 1/4]
    u8 = np.arange(1.0, 9.0)
    ee, g = extrap(GeoKind.Qua8, 9)
    np.testing.assert_allclose(ee @ (interp_matrix(GeoKind.Qua8, g) @ u8), u8, atol=1e-13)
    # fewer points than nodes: a linear field is still recovered exactly (the hat_xi part of the formula)
    ee, g = extrap(GeoKind.Qua8, 4)
    xi = ref_coords(GeoKind.Qua8)
    lin = 2.0 + 3.0 * xi[:, 0] - 1.5 * xi[:, 1]
    np.testing.assert_allclose(ee @ (interp_matrix(GeoKind.Qua8, g) @ lin), lin, atol=1e-13)


def test_reference_coords_are_the_nodes_of_the_basis():
    """N_m(reference_coords(n)) = delta_mn for every kind (src/shapes tests `interp_works` of the reference)"""
    for kind in GeoKind:
        xi = ref_coords(kind)
        for n in range(kind.nnode()):
            nn, _ = kind.interp_deriv(xi[n])
            expect = np.zeros(kind.nnode())
            expect[n] = 1.0
            assert np.max(np.abs(nn - expect)) <= 1e-14, (kind.to_string(), n)

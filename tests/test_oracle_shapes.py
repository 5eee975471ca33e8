"""Shape-level checks of the oracle, following src/shapes/scratchpad_testing.rs:405-515 and
scratchpad_calc_jacobian.rs:179-245 (Kronecker delta at nodes, dN/dksi and dx/dksi by central differences)."""
import numpy as np
import pytest

from oracle import pyoracle as po

SUPPORTED = ["lin2", "lin3", "tri3", "tri6", "qua4", "qua8", "qua9", "tet4", "tet10", "hex8", "hex20"]


@pytest.mark.parametrize("kind", SUPPORTED)
def test_interp_is_kronecker_delta_at_nodes(kind):
    n = po.kind_nnode(kind)
    for m in range(n):
        nn = po.interp(kind, po.reference_coords(kind, m))
        expect = np.zeros(n)
        expect[m] = 1.0
        np.testing.assert_allclose(nn, expect, atol=1e-15)


@pytest.mark.parametrize("kind", SUPPORTED)
def test_deriv_matches_central_difference(kind):
    gd = po.kind_geo_ndim(kind)
    at = np.full(gd, 0.25)
    ll = po.deriv(kind, at)
    h = 1e-6
    for j in range(gd):
        e = np.zeros(gd)
        e[j] = h
        num = (po.interp(kind, at + e) - po.interp(kind, at - e)) / (2 * h)
        np.testing.assert_allclose(ll[:, j], num, atol=1e-9)
    # partition of unity => derivative columns sum to zero
    np.testing.assert_allclose(ll.sum(axis=0), 0.0, atol=1e-14)


def ring_sector_coords(kind):
    """gen_scratchpad_with_coords (src/shapes/scratchpad_testing.rs:68-103): r in [5,10], alpha in [30,60] deg, z in [0,4]."""
    gd = po.kind_geo_ndim(kind)
    n = po.kind_nnode(kind)
    rmin, rmax, amin, amax, zmin, zmax = 5.0, 10.0, np.pi / 6, np.pi / 3, 0.0, 4.0
    tri_tet = po.kind_class(kind) in (po.CLASS_TRI, po.CLASS_TET)
    xx = np.zeros((n, max(gd, 2)))
    for m in range(n):
        ksi = po.reference_coords(kind, m)
        lo, span = (0.0, 1.0) if tri_tet else (-1.0, 2.0)
        r = rmin + (ksi[0] - lo) * (rmax - rmin) / span
        a = amin + ((ksi[1] if gd > 1 else 0.0) - lo) * (amax - amin) / span
        xx[m, 0], xx[m, 1] = r * np.cos(a), r * np.sin(a)
        if gd == 3:
            xx[m, 2] = zmin + (ksi[2] - lo) * (zmax - zmin) / span
    return xx


@pytest.mark.parametrize("kind", [k for k in SUPPORTED if not k.startswith("lin")])
def test_jacobian_and_gradient_by_finite_differences(kind):
    gd = po.kind_geo_ndim(kind)
    xx = ring_sector_coords(kind)
    pad = po.Scratchpad.new(gd, kind).set_coords(xx)
    at = np.full(gd, 0.25)
    det = pad.calc_gradient(at)
    jac, grad, inv = pad.jacobian, pad.gradient, pad.inv_jacobian
    h = 1e-6
    for j in range(gd):
        e = np.zeros(gd)
        e[j] = h
        num = (pad.calc_coords(at + e) - pad.calc_coords(at - e)) / (2 * h)
        np.testing.assert_allclose(jac[:, j], num, atol=1e-8)
    assert abs(det - np.linalg.det(jac)) < 1e-12 * abs(det)
    np.testing.assert_allclose(inv @ jac, np.eye(gd), atol=1e-13)
    np.testing.assert_allclose(grad, po.deriv(kind, at) @ inv, atol=1e-14)
    assert det > 0


def test_gauss_rules_integrate_reference_volume():
    vol = {po.CLASS_LIN: 2.0, po.CLASS_TRI: 0.5, po.CLASS_QUA: 4.0, po.CLASS_TET: 1.0 / 6.0, po.CLASS_HEX: 8.0}
    rules = {po.CLASS_LIN: [1, 2, 3, 4, 5], po.CLASS_TRI: [1, 3, 4, 6, 7, 12, 16], po.CLASS_QUA: [1, 4, 9, 16],
             po.CLASS_TET: [1, 4, 5, 8, 14, 15, 24], po.CLASS_HEX: [6, 8, 14, 27, 64]}
    for cls, ns in rules.items():
        for n in ns:
            g = po.Gauss.new_sized(cls, n)
            assert g.npoint() == n
            assert abs(g.data[:, 3].sum() - vol[cls]) < 1e-13
    # defaults (src/integ/gauss.rs:87-121)
    for kind, n in [("lin2", 2), ("tri3", 3), ("tri6", 7), ("qua4", 4), ("qua8", 9), ("qua9", 9), ("tet4", 4),
                    ("tet10", 14), ("hex8", 8), ("hex20", 27), ("hex32", 64), ("tet20", 24), ("qua16", 16)]:
        assert po.Gauss.new(kind).npoint() == n
    # verbatim digits: the 27-point Hex rule uses a 15-digit abscissa, the 9-point Qua rule 16 digits
    assert po.Gauss.new("hex20").data[0, 0] == -0.774596669241483
    assert po.Gauss.new("qua8").data[0, 0] == -0.7745966692414834
    assert po.Gauss.new("qua8").data[0, 3] == 25.0 / 81.0

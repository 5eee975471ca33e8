"""Pins the CPU oracle against the reference's own known-answer tests (tests/golden/known_answers.json).

Each case cites the reference file:line it was transcribed from (see tests/golden/make_goldens.py).
As in the reference's tests, outputs are pre-filled with NOISE = 1234.56 (src/integ/testing.rs:7) to
check clear-before-sum, comparisons are ABSOLUTE differences (russell_lab::mat_approx_eq), and
mat_10_bdb is cross-checked against the independent mat_10_bdb_alt (B-matrix form).
"""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import pyoracle as po

GOLD = json.loads((Path(__file__).parent / "golden" / "known_answers.json").read_text())
NOISE = GOLD["noise"]
CASES = GOLD["cases"]


def make_pad(case):
    pad = po.Scratchpad.new(case["ndim"], case["kind"])
    pad.set_coords(case["xx"])
    return pad


def make_gauss(case, n):
    return po.Gauss.new(case["kind"]) if n == 0 else po.Gauss(case["kind"], n)


def tolerances(case):
    tol = case["tol"]
    return tol if isinstance(tol, list) else [tol] * len(case["ngauss"])


def run_case(case, n, alt=False):
    pad = make_pad(case)
    gauss = make_gauss(case, n)
    args = po.CommonArgs(pad, gauss)
    args.alpha = case.get("alpha", 1.0)
    args.axisymmetric = case.get("axisymmetric", False)
    op = case["op"]
    nnode, ndim = pad.nnode, pad.space_ndim
    x_ips = po.get_points_coords(pad, gauss)
    if op == "mat_10_bdb":
        dd = np.array(case["dd"])
        kk = np.full((nnode * ndim, nnode * ndim), NOISE)

        def fn(out, p, nn, bb):
            out[:] = dd.ravel(order="F")

        return (po.mat_10_bdb_alt if alt else po.mat_10_bdb)(kk, args, fn)
    if op == "mat_01_nsn":
        kk = np.full((nnode, nnode), NOISE)

        def fn(out, p, nn, bb):
            out[0] = case["s"]

        return po.mat_01_nsn(kk, args, fn)
    if op == "mat_03_btb":
        kk = np.full((nnode, nnode), NOISE)

        def fn(out, p, nn, bb):
            out[:] = case["tt"]

        return po.mat_03_btb(kk, args, fn)
    if op in ("mat_02_bvn", "mat_08_ntn", "mat_09_nvb"):
        n = nnode if op == "mat_02_bvn" else nnode * ndim
        kk = np.full((n, n), NOISE)

        def fn(out, p, nn, bb):
            out[:] = case["tt"] if op == "mat_08_ntn" else case["v"]

        return getattr(po, op)(kk, args, fn)
    if op == "vec_01_ns":
        a = np.full(nnode, NOISE)

        def fn(out, p, nn, bb):
            if "s_field" in case:
                out[0] = x_ips[p][int(case["s_field"][1])]
            else:
                out[0] = case["s"]

        return po.vec_01_ns(a, args, fn)
    if op in ("vec_02_nv", "vec_03_bv"):
        v = np.full(nnode * ndim if op == "vec_02_nv" else nnode, NOISE)

        def fn(out, p, nn, bb):
            if case.get("v_field") == "x":
                out[:] = x_ips[p]
            elif case.get("v_field") == "x0x0":
                out[:] = x_ips[p][0]
            else:
                out[:] = case["v"]

        return (po.vec_02_nv if op == "vec_02_nv" else po.vec_03_bv)(v, args, fn)
    if op == "vec_04_bt":
        d = np.full(nnode * ndim, NOISE)

        def fn(out, p, nn, bb):
            out[:] = case["tt"]

        return po.vec_04_bt(d, args, fn)
    raise AssertionError(op)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_known_answer(case):
    expect = np.array(case["expect"])
    for n, tol in zip(case["ngauss"], tolerances(case)):
        got = run_case(case, n)
        assert got.shape == expect.shape
        assert np.max(np.abs(got - expect)) <= tol, f"{case['name']} ngauss={n} ({case['ref']})"
        if "alt_tol" in case:
            got_alt = run_case(case, n, alt=True)
            assert np.max(np.abs(got - got_alt)) <= case["alt_tol"]


def test_lin_elasticity_matches_transcribed_modulus():
    for case in CASES:
        if "young" in case:
            dd = po.lin_elasticity(case["young"], case["poisson"], case["two_dim"], case["plane_stress"])
            np.testing.assert_array_equal(dd, np.array(case["dd"]))


@pytest.mark.parametrize("sf", GOLD["scalar_field"], ids=[s["name"] for s in GOLD["scalar_field"]])
def test_scalar_field(sf):
    case = dict(kind=sf["kind"], ndim=sf["ndim"], xx=sf["xx"])
    for n in sf["ngauss"]:
        pad = make_pad(case)
        gauss = po.Gauss(sf["kind"], n)
        vol = po.scalar_field(pad, gauss, lambda p: 1.0)
        assert abs(vol - sf.get("volume", sf.get("area"))) <= sf["tol"]
        if "int_x2" in sf:
            x = po.get_points_coords(pad, gauss)
            i2 = po.scalar_field(pad, gauss, lambda p: float(np.sum(x[p] ** 2)))
            i3 = po.scalar_field(pad, gauss, lambda p: float(np.sum(x[p] ** 3)))
            assert abs(i2 - sf["int_x2"]) <= sf["tol"]
            assert abs(i3 - sf["int_x3"]) <= sf["tol"]


def test_jacobian_and_gradient_doctests():
    j = GOLD["jacobian_doctest"]
    pad = make_pad(j)
    det = pad.calc_jacobian(j["ksi"])
    assert abs(det - j["det"]) <= j["tol"]
    np.testing.assert_allclose(pad.jacobian, np.array(j["jacobian"]), atol=j["tol"], rtol=0)
    det2 = pad.calc_gradient(j["ksi"])
    assert det2 == det
    np.testing.assert_allclose(pad.gradient, np.array(j["gradient"]), atol=j["tol"], rtol=0)


def test_error_strings():
    """src/integ/mat_10_bdb.rs:367-410, mat_01_nsn.rs:111-146, vec_01_ns.rs:143-155,
    scratchpad_calc_gradient.rs:105-117, scratchpad_calc_jacobian.rs (ok_xxt)."""
    pad = po.Scratchpad.new(2, "lin2")
    pad.set_coords([[3.0, 4.0], [3.0 + 1 / np.sqrt(2), 4.0 + 1 / np.sqrt(2)]])
    gauss = po.Gauss.new("lin2")
    args = po.CommonArgs(pad, gauss)
    f = lambda out, p, nn, bb: None  # noqa: E731
    args.ii0 = 1
    with pytest.raises(po.OracleError, match="nrow\\(K\\) must be ≥ ii0 \\+ nnode ⋅ space_ndim"):
        po.mat_10_bdb(np.zeros((4, 4)), args, f)
    args.ii0, args.jj0 = 0, 1
    with pytest.raises(po.OracleError, match="ncol\\(K\\) must be ≥ jj0 \\+ nnode ⋅ space_ndim"):
        po.mat_10_bdb(np.zeros((4, 4)), args, f)
    args.jj0 = 0
    with pytest.raises(po.OracleError, match="calc_gradient requires that geo_ndim = space_ndim"):
        po.mat_10_bdb(np.zeros((4, 4)), args, f)
    with pytest.raises(po.OracleError, match="calc_gradient requires that geo_ndim = space_ndim"):
        po.mat_01_nsn(np.zeros((2, 2)), args, f)
    args.ii0 = 1
    with pytest.raises(po.OracleError, match="nrow\\(K\\) must be ≥ ii0 \\+ nnode$"):
        po.mat_01_nsn(np.zeros((2, 2)), args, f)
    with pytest.raises(po.OracleError, match="a.len\\(\\) must be ≥ ii0 \\+ nnode"):
        po.vec_01_ns(np.zeros(2), args, f)
    with pytest.raises(po.OracleError, match="c.len\\(\\) must be ≥ ii0 \\+ nnode"):
        po.vec_03_bv(np.zeros(2), args, f)
    with pytest.raises(po.OracleError, match="d.len\\(\\) must be ≥ ii0 \\+ nnode ⋅ space_ndim"):
        po.vec_04_bt(np.zeros(4), args, f)

    # callback error propagates ("stop")
    pad = po.Scratchpad.new(2, "tri3").set_coords([[3.0, 4.0], [8.0, 4.0], [5.0, 9.0]])
    args = po.CommonArgs(pad, po.Gauss.new("tri3"))

    def stop(out, p, nn, bb):
        raise StopIteration

    with pytest.raises(po.OracleError, match="stop"):
        po.mat_10_bdb(np.zeros((6, 6)), args, stop)

    # axisymmetric flag needs 2D
    pad = po.Scratchpad.new(3, "tet4").set_coords([[2, 3, 4], [6, 3, 2], [2, 5, 1], [4, 3, 6]])
    args = po.CommonArgs(pad, po.Gauss.new("tet4"))
    args.axisymmetric = True
    with pytest.raises(po.OracleError, match="axisymmetric requires space_ndim = 2"):
        po.mat_10_bdb(np.zeros((12, 12)), args, f)

    # coordinates not set
    pad = po.Scratchpad.new(2, "tri3")
    assert not pad.ok_xxt
    with pytest.raises(po.OracleError, match="all components of the coordinates matrix must be set first"):
        pad.calc_jacobian([0.0, 0.0])
    pad.set_xx(2, 1, 0.0)
    assert pad.ok_xxt  # flips when the LAST entry is set (scratchpad.rs:262-269)
    pad.set_xx(0, 0, 0.0)
    assert not pad.ok_xxt

    # zero determinant
    pad = po.Scratchpad.new(2, "tri3").set_coords([[0, 0], [1, 0], [2, 0]])
    with pytest.raises(po.OracleError, match="cannot compute inverse due to zero determinant"):
        pad.calc_jacobian([0.0, 0.0])

    # gauss rule not available
    with pytest.raises(po.OracleError, match="not available for Hex class"):
        po.Gauss("hex8", 7)


def test_cable_and_shell_jacobian():
    """src/shapes/scratchpad_calc_jacobian.rs:248-276"""
    pad = po.Scratchpad.new(2, "lin2").set_coords([[0.0, 0.0], [3.0, 4.0]])
    assert abs(pad.calc_jacobian([0.0]) - 2.5) < 1e-15  # |dx/dksi| = L/2
    pad = po.Scratchpad.new(3, "qua4").set_coords([[0, 0, 0], [1, 0, 0], [1, 1, 0], [0, 1, 0]])
    assert pad.calc_jacobian([0.0, 0.0]) == -1.0

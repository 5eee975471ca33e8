"""The oracle's two-pad coupling matrices (integ::mat_04_nsb .. mat_07_bsn), boundary integrals (*_bry) and calc_normal_vector
against the reference's own known answers (tests/golden/coupled_and_boundary.py, transcribed with file:line)."""
import sys
from pathlib import Path

import numpy as np
import pytest

sys.path.insert(0, str(Path(__file__).parent / "golden"))
import coupled_and_boundary as gold  # noqa: E402

from oracle import pyoracle as po  # noqa: E402


def one_cell(kind, coords):
    n = po.kind_nnode(kind)
    return (np.asarray(coords, dtype=float), np.array([po.kind_id(kind)], dtype=np.uint8), np.array([0, n], dtype=np.uint64),
            np.arange(n, dtype=np.uint64))


COUPLED = gold.coupled_cases()
BOUNDARY = gold.boundary_cases()


@pytest.mark.parametrize("case", COUPLED, ids=[c[0] for c in COUPLED])
def test_coupled_known_answers(case):
    name, ref, op, kind, kind_b, ndim, coords, rec, expect, rules, tol = case
    c, k, off, conn = one_cell(kind, coords)
    for n in rules:
        got = po.mesh_integ_ext(op, ndim, c, k, off, conn, kind_b=po.kind_id(kind_b), ngauss=n, coef=np.asarray(rec, dtype=float))
        got = got.reshape(expect.shape, order="F")
        assert np.max(np.abs(got - expect)) <= tol, f"{name} rule {n} ({ref})"


@pytest.mark.parametrize("case", BOUNDARY, ids=[c[0] for c in BOUNDARY])
def test_boundary_known_answers(case):
    name, ref, op, kind, ndim, coords, rec, expect, rule, tol = case
    c, k, off, conn = one_cell(kind, coords)
    got = po.mesh_integ_ext(op, ndim, c, k, off, conn, ngauss=rule, coef=np.asarray(rec, dtype=float))
    got = got.reshape(expect.shape, order="F") if expect.ndim == 2 else got
    assert np.max(np.abs(got - expect)) <= tol, f"{name} ({ref})"


def test_normal_vectors():
    """Scratchpad::calc_normal_vector (src/shapes/scratchpad_calc_normal_vector.rs:111-176): a line in 2D (n = e3 x g1) and a flat
    Qua4 in 3D (n = g1 x g2); magnitudes are the length / area Jacobians."""
    c, k, off, conn = one_cell("lin2", [[1.0, 2.0], [4.0, 6.0]])
    out = po.mesh_integ_ext("normal", 2, c, k, off, conn, ngauss=2).reshape(2, 3)
    np.testing.assert_allclose(out[:, :2], [[-0.8, 0.6]] * 2, rtol=1e-15)  # g1 = (1.5, 2): n = (-2, 1.5) / 2.5
    np.testing.assert_allclose(out[:, 2], [2.5, 2.5], rtol=1e-15)          # |dx/dksi| = L / 2
    quad = [[0.0, 0.0, 1.0], [2.0, 0.0, 1.0], [2.0, 3.0, 1.0], [0.0, 3.0, 1.0]]
    c, k, off, conn = one_cell("qua4", quad)
    out = po.mesh_integ_ext("normal", 3, c, k, off, conn, ngauss=4).reshape(4, 4)
    np.testing.assert_allclose(out[:, :3], [[0.0, 0.0, 1.0]] * 4, atol=1e-15)
    np.testing.assert_allclose(out[:, 3], [1.5] * 4, rtol=1e-15)  # area / 4
    with pytest.raises(po.OracleError, match="calc_normal_vector requires geo_ndim = 1 in 2D"):
        c, k, off, conn = one_cell("tri3", [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]])
        po.mesh_integ_ext("normal", 2, c, k, off, conn)


def test_error_strings_of_the_new_cases():
    c, k, off, conn = one_cell("tri3", [[3.0, 4.0], [8.0, 4.0], [5.0, 9.0]])
    with pytest.raises(po.OracleError, match=r"in 2D, geometry ndim must be equal to 1 \(a line\)"):
        po.mesh_integ_ext("mat_01_nsn_bry", 2, c, k, off, conn, coef=np.array([1.0, 0.0, 0.0]))
    c, k, off, conn = one_cell("tet4", gold.PAD_TET4)
    with pytest.raises(po.OracleError, match=r"in 3D, geometry ndim must be equal to 2 \(a surface\)"):
        po.mesh_integ_ext("vec_02_nv_bry", 3, c, k, off, conn, coef=np.array([0.0, 0.0, 0.0, 1.0]))

"""The reference-element tables shipped with the library (gemlab_b200/csrc/shape_tables.inc, generated data) against
(1) the oracle's node-by-node restatement of the reference formulas -- bit for bit, and (2) the library's own,
independently written family-wise formulas (gl_shape_eval) -- to a few ulp.  CPU only."""
import re
from pathlib import Path

import numpy as np
import pytest

import gemlab_b200  # noqa: F401  (loads the C ABI)
from gemlab_b200 import GeoKind
from gemlab_b200._lib import lib
from oracle import pyoracle as po

INC = Path(__file__).resolve().parent.parent / "gemlab_b200" / "csrc" / "shape_tables.inc"


def parse_tables():
    text = INC.read_text()
    out = {}
    for m in re.finditer(r"static const double SHAPE_([A-Z0-9]+)_(\d+)\[(\d+)\] = \{(.*?)\};", text, re.S):
        vals = [float.fromhex(v) for v in m.group(4).replace("\n", " ").split(",") if v.strip()]
        assert len(vals) == int(m.group(3))
        out[(m.group(1).lower(), int(m.group(2)))] = np.array(vals)
    return out


TABLES = parse_tables()
REF_COORDS = __import__("json").loads((Path(__file__).resolve().parent / "golden" / "kind_reference_coords.json").read_text())


def test_every_supported_kind_and_rule_is_tabulated():
    L = lib()
    for k in range(20):
        if not L.gl_kind_supported(k):
            continue
        name = L.gl_kind_name(k).decode()
        for ng in range(1, 70):
            try:
                po.Gauss.new_sized(po.kind_class(name), ng)
            except Exception:
                continue
            assert (name, ng) in TABLES, (name, ng)


@pytest.mark.parametrize("key", sorted(TABLES), ids=[f"{k}-{n}" for k, n in sorted(TABLES)])
def test_table_matches_oracle_bitwise_and_library_formulas(key):
    kind, ng = key
    g = po.Gauss.new_sized(po.kind_class(kind), ng)
    nn, gd = po.kind_nnode(kind), po.kind_geo_ndim(kind)
    tab = TABLES[key]
    N = tab[:ng * nn].reshape(ng, nn)
    Lt = tab[ng * nn:].reshape(ng, nn, gd)
    kid = int(GeoKind.from_str(kind))
    for p in range(ng):
        ksi = np.array(list(g.coords(p))[:3], dtype=float)
        assert np.array_equal(N[p], np.asarray(po.interp(kind, ksi)))          # bit for bit
        assert np.array_equal(Lt[p], np.asarray(po.deriv(kind, ksi)).reshape(nn, gd))
        n2, l2 = np.zeros(nn), np.zeros(nn * gd)
        # independent evaluation at an arbitrary point: family-wise formulas, and for the cubic / quartic kinds the nodal basis
        # built from the node coordinates and the polynomial space alone (gl_shapes.cpp) -- for ALL 20 kinds
        assert lib().gl_shape_eval(kid, ksi.ctypes.data, n2.ctypes.data, l2.ctypes.data) == 0
        assert np.max(np.abs(n2 - N[p])) <= 1e-14
        assert np.max(np.abs(l2.reshape(nn, gd) - Lt[p])) <= 1e-14
    # partition of unity and zero-sum derivatives (shapes.rs tests of the reference)
    assert np.max(np.abs(N.sum(axis=1) - 1.0)) <= 1e-14
    assert np.max(np.abs(Lt.sum(axis=1))) <= 1e-13
    # isoparametric consistency: sum_m N_m xi_m = ksi and sum_m xi_m (x) dN_m = I at every Gauss point, with xi_m the
    # reference's node coordinates (complete polynomial bases reproduce linear fields) -- an independent check of the
    # tables of the kinds that exist as tables only
    xi = np.array(REF_COORDS[kind])
    for p in range(ng):
        ksi = np.array(list(g.coords(p))[:gd], dtype=float)
        assert np.max(np.abs(N[p] @ xi - ksi)) <= 1e-14
        assert np.max(np.abs(xi.T @ Lt[p] - np.eye(gd))) <= 1e-13

#!/usr/bin/env python3
"""Writes tests/golden/known_answers.json: the known-answer vectors that pin the integration path.

The reference (cpmech/gemlab v2.4.0) is Rust and cannot be executed in the authoring container
(no cargo/rustc, russell_* not vendored), so these vectors are TRANSCRIBED from the reference's own
tests -- literal matrices from the books it cites and the closed-form element matrices of its
`Analytical*` helpers evaluated here in numpy -- with the reference file:line of each case:

  src/integ/mat_10_bdb.rs:69-117      Tet4 12x12 stiffness (doc-test, E=480, nu=1/3)
  src/integ/mat_10_bdb.rs:413-501     Tri3 plane-stress (Bhatti) + AnalyticalTri3::mat_10_bdb
  src/integ/mat_10_bdb.rs:504-546     Tet4 vs AnalyticalTet4::mat_10_bdb
  src/integ/mat_10_bdb.rs:549-676     axisymmetric Qua4 (Felippa A-FEM 12-9), rules 1, 4, 9, 16
  src/integ/mat_01_nsn.rs:150-229     Tri3/Qua4/Qua8/Tet4 vs analytical
  src/integ/mat_03_btb.rs:180-277     Tri3/Qua4/Qua8/Tet4 vs analytical
  src/integ/vec_01_ns.rs:58-82,158-279, vec_03_bv.rs:50-83,171-265, vec_02_nv.rs, vec_04_bt.rs
  tests/test_integ_heat_axis_1.rs     Lewis et al. axisymmetric heat (conduction part)
  src/integ/scalar_field.rs:110-230   rotated square / slanted Hex8 integrals
  src/shapes/scratchpad_calc_jacobian.rs:57-99, scratchpad_calc_gradient.rs:40-76  doc-tests

Run: python tests/golden/make_goldens.py   (no access to the reference tree is needed).
"""
import json
from pathlib import Path

import numpy as np

SQRT_2 = np.sqrt(2.0)
HERE = Path(__file__).resolve().parent

# element fixtures: src/integ/testing.rs:10-84
PAD_TRI3 = [[3.0, 4.0], [8.0, 4.0], [5.0, 9.0]]
PAD_TET4 = [[2.0, 3.0, 4.0], [6.0, 3.0, 2.0], [2.0, 5.0, 1.0], [4.0, 3.0, 6.0]]


def pad_lin2(length):
    return [[3.0, 4.0], [3.0 + length / SQRT_2, 4.0 + length / SQRT_2]]


def pad_qua4(xc, yc, a, b):
    return [[xc - a, yc - b], [xc + a, yc - b], [xc + a, yc + b], [xc - a, yc + b]]


def pad_qua8(xc, yc, a, b):
    return pad_qua4(xc, yc, a, b) + [[xc, yc - b], [xc + a, yc], [xc, yc + b], [xc - a, yc]]


def elastic_mandel(young, poisson, two_dim, plane_stress):
    """LinElasticity::get_modulus().matrix() (russell_tensor)."""
    ns = 4 if two_dim else 6
    d = np.zeros((ns, ns))
    if two_dim and plane_stress:
        c = young / (1.0 - poisson * poisson)
        d[0, 0] = d[1, 1] = c
        d[0, 1] = d[1, 0] = c * poisson
        d[3, 3] = c * (1.0 - poisson)
    else:
        c = young / ((1.0 + poisson) * (1.0 - 2.0 * poisson))
        for a in range(3):
            for b in range(3):
                d[a, b] = c * (1.0 - poisson) if a == b else c * poisson
        for q in range(3, ns):
            d[q, q] = c * (1.0 - 2.0 * poisson)
    return d


class AnaTri3:
    """src/integ/analytical_tri3.rs:39-104"""

    def __init__(self, xx):
        (x0, y0), (x1, y1), (x2, y2) = xx
        a0, a1, a2 = y1 - y2, y2 - y0, y0 - y1
        b0, b1, b2 = x2 - x1, x0 - x2, x1 - x0
        f0, f1, f2 = x1 * y2 - x2 * y1, x2 * y0 - x0 * y2, x0 * y1 - x1 * y0
        self.area = (f0 + f1 + f2) / 2.0
        r = 2.0 * self.area
        s = r * SQRT_2
        self.bb = np.array([[a0 / r, b0 / r], [a1 / r, b1 / r], [a2 / r, b2 / r]])
        self.bbe = np.array([
            [a0 / r, 0.0, a1 / r, 0.0, a2 / r, 0.0],
            [0.0, b0 / r, 0.0, b1 / r, 0.0, b2 / r],
            [0.0, 0.0, 0.0, 0.0, 0.0, 0.0],
            [b0 / s, a0 / s, b1 / s, a1 / s, b2 / s, a2 / s],
        ])
        self.x = (x0, x1, x2)
        self.y = (y0, y1, y2)

    def vec_01_ns(self, s, axisymmetric):  # :106-118
        x0, x1, x2 = self.x
        if axisymmetric:
            c = s * self.area / 12.0
            return np.array([c * (2 * x0 + x1 + x2), c * (x0 + 2 * x1 + x2), c * (x0 + x1 + 2 * x2)])
        c = s * self.area / 3.0
        return np.array([c, c, c])

    def vec_02_nv(self, v0, v1):  # :139-148
        return np.array([v0, v1, v0, v1, v0, v1]) * self.area / 3.0

    def vec_03_bv(self, w0, w1):  # :150-156
        return (w0 * self.bb[:, 0] + w1 * self.bb[:, 1]) * self.area

    def vec_03_bv_bilinear(self):  # :159-167
        sx, sy = sum(self.x), sum(self.y)
        return (sx * self.bb[:, 0] + sy * self.bb[:, 1]) * self.area / 3.0

    def vec_04_bt(self, t, axisymmetric):  # :170-197, t = Mandel vector (4)
        x0, x1, x2 = self.x
        t0, t1, t2, t3 = t
        out = np.zeros(6)
        for m in range(3):
            bm0, bm1 = self.bb[m]
            if axisymmetric:
                c = self.area / 3.0
                out[2 * m] = c * (bm0 * t0 + bm1 * t3 / SQRT_2) * (x0 + x1 + x2) + c * t2
                out[2 * m + 1] = c * (bm1 * t1 + bm0 * t3 / SQRT_2) * (x0 + x1 + x2)
            else:
                c = self.area
                out[2 * m] = c * (bm0 * t0 + bm1 * t3 / SQRT_2)
                out[2 * m + 1] = c * (bm1 * t1 + bm0 * t3 / SQRT_2)
        return out

    def mat_01_nsn(self, s, th):  # :199-202
        c = th * s * self.area / 12.0
        return np.array([[2 * c, c, c], [c, 2 * c, c], [c, c, 2 * c]])

    def mat_02_bvn(self, v0, v1):  # :253-261
        row = self.area * (self.bb[:, 0] * v0 + self.bb[:, 1] * v1) / 3.0
        return np.repeat(row[:, None], 3, axis=1)

    def mat_08_ntn(self, rho, th):  # :300-310
        c = th * rho * self.area / 12.0
        return c * np.kron(np.array([[2.0, 1.0, 1.0], [1.0, 2.0, 1.0], [1.0, 1.0, 2.0]]), np.eye(2))

    def mat_09_nvb(self, v0, v1):  # :314-327: K[2m+i][2n+j] = (area/3) v_i B[n][j]
        c = self.area / 3.0
        k = np.zeros((6, 6))
        for m in range(3):
            for n in range(3):
                for i, vi in enumerate((v0, v1)):
                    for j in range(2):
                        k[2 * m + i, 2 * n + j] = c * self.bb[n, j] * vi
        return k

    def mat_03_btb(self, kx, ky, axisymmetric):  # :257-273
        c = self.area * sum(self.x) / 3.0 if axisymmetric else self.area
        b = self.bb
        k = np.zeros((3, 3))
        for m in range(3):
            for n in range(3):
                k[m, n] = c * (b[m, 0] * b[n, 0] * kx + b[m, 1] * b[n, 1] * ky)
        return k

    def mat_10_bdb(self, young, poisson, plane_stress, th):  # :304-315
        dd = elastic_mandel(young, poisson, True, plane_stress)
        return th * self.area * (self.bbe.T @ dd @ self.bbe)


class AnaTet4:
    """src/integ/analytical_tet4.rs:31-139"""

    def __init__(self, xx):
        (x1, y1, z1), (x2, y2, z2), (x3, y3, z3), (x4, y4, z4) = xx
        x12, x13, x14, x23, x24, x34 = x1 - x2, x1 - x3, x1 - x4, x2 - x3, x2 - x4, x3 - x4
        x21, x31, x32, x42, x43 = -x12, -x13, -x23, -x24, -x34
        y12, y13, y14, y23, y24, y34 = y1 - y2, y1 - y3, y1 - y4, y2 - y3, y2 - y4, y3 - y4
        y21, y31, y32, y42, y43 = -y12, -y13, -y23, -y24, -y34
        z12, z13, z14, z23, z24, z34 = z1 - z2, z1 - z3, z1 - z4, z2 - z3, z2 - z4, z3 - z4
        z21, z31, z32, z42, z43 = -z12, -z13, -z23, -z24, -z34
        jj_det = x21 * (y23 * z34 - y34 * z23) + x32 * (y34 * z12 - y12 * z34) + x43 * (y12 * z23 - y23 * z12)
        a1 = y42 * z32 - y32 * z42
        b1 = x32 * z42 - x42 * z32
        c1 = x42 * y32 - x32 * y42
        a2 = y31 * z43 - y34 * z13
        b2 = x43 * z31 - x13 * z34
        c2 = x31 * y43 - x34 * y13
        a3 = y24 * z14 - y14 * z24
        b3 = x14 * z24 - x24 * z14
        c3 = x24 * y14 - x14 * y24
        a4 = y13 * z21 - y12 * z31
        b4 = x21 * z13 - x31 * z12
        c4 = x13 * y21 - x12 * y31
        r = jj_det
        s = r * SQRT_2
        self.volume = jj_det / 6.0
        self.bb = np.array([[a1, b1, c1], [a2, b2, c2], [a3, b3, c3], [a4, b4, c4]]) / r
        a = [a1, a2, a3, a4]
        b = [b1, b2, b3, b4]
        c = [c1, c2, c3, c4]
        bbe = np.zeros((6, 12))
        for m in range(4):
            bbe[0, 3 * m + 0] = a[m] / r
            bbe[1, 3 * m + 1] = b[m] / r
            bbe[2, 3 * m + 2] = c[m] / r
            bbe[3, 3 * m + 0] = b[m] / s
            bbe[3, 3 * m + 1] = a[m] / s
            bbe[4, 3 * m + 1] = c[m] / s
            bbe[4, 3 * m + 2] = b[m] / s
            bbe[5, 3 * m + 0] = c[m] / s
            bbe[5, 3 * m + 2] = a[m] / s
        self.bbe = bbe
        self.z = (z1, z2, z3, z4)

    def vec_01_ns(self, cs):  # :141-148
        return np.full(4, cs * self.volume / 4.0)

    def vec_01_ns_linear_along_z(self):  # :151-164
        z1, z2, z3, z4 = self.z
        v = self.volume
        return np.array([v * (2 * z1 + z2 + z3 + z4), v * (z1 + 2 * z2 + z3 + z4), v * (z1 + z2 + 2 * z3 + z4),
                         v * (z1 + z2 + z3 + 2 * z4)]) / 20.0

    def vec_02_nv(self, v0, v1, v2):  # :167-182
        return np.tile([v0, v1, v2], 4) * self.volume / 4.0

    def vec_03_bv(self, w0, w1, w2):  # :185-192
        return (w0 * self.bb[:, 0] + w1 * self.bb[:, 1] + w2 * self.bb[:, 2]) * self.volume

    def vec_04_bt(self, sig):  # :195-215 (sig = full 3x3 symmetric matrix)
        return (self.volume * (self.bb @ sig)).reshape(-1)

    def mat_01_nsn(self, s):  # :218-226
        c = self.volume / 20.0
        k = np.full((4, 4), s * c)
        np.fill_diagonal(k, 2.0 * s * c)
        return k

    def mat_02_bvn(self, v0, v1, v2):  # :262-275
        row = self.volume / 4.0 * (self.bb @ np.array([v0, v1, v2]))
        return np.repeat(row[:, None], 4, axis=1)

    def mat_08_ntn(self, sig):  # :378-398: K[3m+i][3n+j] = sig[i][j] * V * (1/10 if m == n else 1/20)
        w = np.full((4, 4), 1.0 / 20.0) + np.eye(4) / 20.0
        return self.volume * np.kron(w, np.asarray(sig))

    def mat_09_nvb(self, v0, v1, v2):  # :402-425: K[3m+i][3n+j] = (V/4) v_i B[n][j]
        c = self.volume / 4.0
        k = np.zeros((12, 12))
        for m in range(4):
            for n in range(4):
                for i, vi in enumerate((v0, v1, v2)):
                    for j in range(3):
                        k[3 * m + i, 3 * n + j] = c * self.bb[n, j] * vi
        return k

    def mat_03_btb(self, tt):  # :250-266: K[m][n] = c * sum_ab B[n][a] * T[b][a] * B[m][b] (T symmetric)
        return self.volume * (self.bb @ tt @ self.bb.T)

    def mat_10_bdb(self, young, poisson):  # :431-
        dd = elastic_mandel(young, poisson, False, False)
        return self.volume * (self.bbe.T @ dd @ self.bbe)


def ana_qua4_mat_01_nsn(a, b, s, th):  # src/integ/analytical_qua4.rs:21-29
    c = th * s * a * b / 9.0
    return c * np.array([[4, 2, 1, 2], [2, 4, 2, 1], [1, 2, 4, 2], [2, 1, 2, 4]], dtype=float)


def ana_qua4_mat_03_btb(a, b, kx, ky):  # src/integ/analytical_qua4.rs:36-45
    p, q = b * kx, a * ky
    return np.array([
        [p / (3 * a) + q / (3 * b), -p / (3 * a) + q / (6 * b), -p / (6 * a) - q / (6 * b), p / (6 * a) - q / (3 * b)],
        [-p / (3 * a) + q / (6 * b), p / (3 * a) + q / (3 * b), p / (6 * a) - q / (3 * b), -p / (6 * a) - q / (6 * b)],
        [-p / (6 * a) - q / (6 * b), p / (6 * a) - q / (3 * b), p / (3 * a) + q / (3 * b), -p / (3 * a) + q / (6 * b)],
        [p / (6 * a) - q / (3 * b), -p / (6 * a) - q / (6 * b), -p / (3 * a) + q / (6 * b), p / (3 * a) + q / (3 * b)],
    ])


def ana_qua8_mat_01_nsn(a, b, s, th):  # src/integ/analytical_qua8.rs:23-37
    c = -th * s * a * b
    m215, m245, m115 = -(2.0 / 15.0) * c, -(2.0 / 45.0) * c, -(1.0 / 15.0) * c
    p215, p845 = (2.0 * c) / 15.0, (8.0 * c) / 45.0
    m3245, m49, m1645 = -(32.0 / 45.0) * c, -(4.0 / 9.0) * c, -(16.0 / 45.0) * c
    return np.array([
        [m215, m245, m115, m245, p215, p845, p845, p215],
        [m245, m215, m245, m115, p215, p215, p845, p845],
        [m115, m245, m215, m245, p845, p215, p215, p845],
        [m245, m115, m245, m215, p845, p845, p215, p215],
        [p215, p215, p845, p845, m3245, m49, m1645, m49],
        [p845, p215, p215, p845, m49, m3245, m49, m1645],
        [p845, p845, p215, p215, m1645, m49, m3245, m49],
        [p215, p845, p845, p215, m49, m1645, m49, m3245],
    ])


def ana_qua8_mat_03_btb(a, b, kx, ky):  # src/integ/analytical_qua8.rs:43-57
    p, q = b * kx / a, a * ky / b
    e = [
        26 * p / 45 + 26 * q / 45,  # 0
        14 * p / 45 + 17 * q / 90,  # 1
        23 * p / 90 + 23 * q / 90,  # 2
        17 * p / 90 + 14 * q / 45,  # 3
        -8 * p / 9 + q / 15,        # 4
        -p / 15 - 4 * q / 9,        # 5
        -4 * p / 9 - q / 15,        # 6
        p / 15 - 8 * q / 9,         # 7
        16 * p / 9 + 8 * q / 15,    # 8
        8 * p / 9 - 8 * q / 15,     # 9
        8 * p / 15 + 16 * q / 9,    # 10
        -8 * p / 15 + 8 * q / 9,    # 11
    ]
    return np.array([
        [e[0], e[1], e[2], e[3], e[4], e[5], e[6], e[7]],
        [e[1], e[0], e[3], e[2], e[4], e[7], e[6], e[5]],
        [e[2], e[3], e[0], e[1], e[6], e[7], e[4], e[5]],
        [e[3], e[2], e[1], e[0], e[6], e[5], e[4], e[7]],
        [e[4], e[4], e[6], e[6], e[8], 0.0, e[9], 0.0],
        [e[5], e[7], e[7], e[5], 0.0, e[10], 0.0, e[11]],
        [e[6], e[6], e[4], e[4], e[9], 0.0, e[8], 0.0],
        [e[7], e[5], e[5], e[7], 0.0, e[11], 0.0, e[10]],
    ])


def L(x):
    return np.asarray(x, dtype=float).tolist()


def main():
    cases = []

    def add(name, ref, op, kind, ndim, xx, expect, tol, ngauss, **kw):
        cases.append(dict(name=name, ref=ref, op=op, kind=kind, ndim=ndim, xx=L(xx), expect=L(expect), tol=tol,
                          ngauss=list(ngauss), **kw))

    # ---- mat_10_bdb -----------------------------------------------------------------------------
    tet4_doc = [
        [745, 540, 120, -5, 30, 60, -270, -240, 0, -470, -330, -180],
        [540, 1720, 270, -120, 520, 210, -120, -1080, -60, -300, -1160, -420],
        [120, 270, 565, 0, 150, 175, 0, -120, -270, -120, -300, -470],
        [-5, -120, 0, 145, -90, -60, -90, 120, 0, -50, 90, 60],
        [30, 520, 150, -90, 220, 90, 60, -360, -60, 0, -380, -180],
        [60, 210, 175, -60, 90, 145, 0, -120, -90, 0, -180, -230],
        [-270, -120, 0, -90, 60, 0, 180, 0, 0, 180, 60, 0],
        [-240, -1080, -120, 120, -360, -120, 0, 720, 0, 120, 720, 240],
        [0, -60, -270, 0, -60, -90, 0, 0, 180, 0, 120, 180],
        [-470, -300, -120, -50, 0, 0, 180, 120, 0, 340, 180, 120],
        [-330, -1160, -300, 90, -380, -180, 60, 720, 120, 180, 820, 360],
        [-180, -420, -470, 60, -180, -230, 0, 240, 180, 120, 360, 520],
    ]
    dd3 = elastic_mandel(480.0, 1.0 / 3.0, False, False)
    add("mat_10_bdb_tet4_doctest", "src/integ/mat_10_bdb.rs:69-117 (printed with {:.0})", "mat_10_bdb", "tet4", 3,
        PAD_TET4, tet4_doc, 0.5, [0], dd=L(dd3), young=480.0, poisson=1.0 / 3.0, two_dim=False, plane_stress=False,
        note="doc-test compares format!({:.0}); tolerance 0.5 = rounding to integers")
    ana = AnaTet4(PAD_TET4)
    add("mat_10_bdb_tet4_analytical", "src/integ/mat_10_bdb.rs:504-546", "mat_10_bdb", "tet4", 3, PAD_TET4,
        ana.mat_10_bdb(480.0, 1.0 / 3.0), 1e-12, [1, 4, 5, 8, 14, 15, 24], dd=L(dd3), young=480.0, poisson=1.0 / 3.0,
        two_dim=False, plane_stress=False, alt_tol=1e-11)

    bhatti_xx = [[0.0, 0.0], [2.0, 0.0], [2.0, 1.5]]
    kk_bhatti = [
        [9.765625000000001e+02, 0.000000000000000e+00, -9.765625000000001e+02, 2.604166666666667e+02, 0.000000000000000e+00, -2.604166666666667e+02],
        [0.000000000000000e+00, 3.906250000000000e+02, 5.208333333333334e+02, -3.906250000000000e+02, -5.208333333333334e+02, 0.000000000000000e+00],
        [-9.765625000000001e+02, 5.208333333333334e+02, 1.671006944444445e+03, -7.812500000000000e+02, -6.944444444444445e+02, 2.604166666666667e+02],
        [2.604166666666667e+02, -3.906250000000000e+02, -7.812500000000000e+02, 2.126736111111111e+03, 5.208333333333334e+02, -1.736111111111111e+03],
        [0.000000000000000e+00, -5.208333333333334e+02, -6.944444444444445e+02, 5.208333333333334e+02, 6.944444444444445e+02, 0.000000000000000e+00],
        [-2.604166666666667e+02, 0.000000000000000e+00, 2.604166666666667e+02, -1.736111111111111e+03, 0.000000000000000e+00, 1.736111111111111e+03],
    ]
    dd_ps = elastic_mandel(10_000.0, 0.2, True, True)
    add("mat_10_bdb_tri3_bhatti", "src/integ/mat_10_bdb.rs:413-470", "mat_10_bdb", "tri3", 2, bhatti_xx, kk_bhatti,
        1e-12, [1], dd=L(dd_ps), alpha=0.25, young=10_000.0, poisson=0.2, two_dim=True, plane_stress=True)
    ana = AnaTri3(bhatti_xx)
    add("mat_10_bdb_tri3_analytical", "src/integ/mat_10_bdb.rs:472-500", "mat_10_bdb", "tri3", 2, bhatti_xx,
        ana.mat_10_bdb(10_000.0, 0.2, True, 0.25), [1e-12, 1e-12, 1e-12, 1e-11, 1e-12], [1, 3, 4, 12, 16], dd=L(dd_ps),
        alpha=0.25, young=10_000.0, poisson=0.2, two_dim=True, plane_stress=True, alt_tol=1e-12)

    felippa = {
        1: [[72.0, 18.0, 36.0, -18.0, -36.0, -18.0, 0.0, 18.0],
            [18.0, 153.0, -54.0, 135.0, -90.0, -153.0, -18.0, -135.0],
            [36.0, -54.0, 144.0, -90.0, 72.0, 54.0, -36.0, 90.0],
            [-18.0, 135.0, -90.0, 153.0, -54.0, -135.0, 18.0, -153.0],
            [-36.0, -90.0, 72.0, -54.0, 144.0, 90.0, 36.0, 54.0],
            [-18.0, -153.0, 54.0, -135.0, 90.0, 153.0, 18.0, 135.0],
            [0.0, -18.0, -36.0, 18.0, 36.0, 18.0, 72.0, -18.0],
            [18.0, -135.0, 90.0, -153.0, 54.0, 135.0, -18.0, 153.0]],
        4: [[168.0, -12.0, 24.0, 12.0, -24.0, -36.0, 48.0, 36.0],
            [-12.0, 108.0, -24.0, 84.0, -72.0, -102.0, -36.0, -90.0],
            [24.0, -24.0, 216.0, -120.0, 0.0, 72.0, -24.0, 72.0],
            [12.0, 84.0, -120.0, 300.0, -72.0, -282.0, 36.0, -102.0],
            [-24.0, -72.0, 0.0, -72.0, 216.0, 120.0, 24.0, 24.0],
            [-36.0, -102.0, 72.0, -282.0, 120.0, 300.0, -12.0, 84.0],
            [48.0, -36.0, -24.0, 36.0, 24.0, -12.0, 168.0, 12.0],
            [36.0, -90.0, 72.0, -102.0, 24.0, 84.0, 12.0, 108.0]],
    }
    # the 3x3 and 4x4 rules only change the (0,0), (0,6), (6,0), (6,6) entries w.r.t. 2x2 (mat_10_bdb.rs:590-610)
    f9 = [row[:] for row in felippa[4]]
    f9[0][0] = f9[6][6] = 232.0
    f9[0][6] = f9[6][0] = 80.0
    f16 = [row[:] for row in felippa[4]]
    f16[0][0] = f16[6][6] = 280.0
    f16[0][6] = f16[6][0] = 104.0
    felippa[9], felippa[16] = f9, f16
    dd_pe = elastic_mandel(96.0, 1.0 / 3.0, True, False)
    fel_tol = {1: 1e-13, 4: 1e-13, 9: 1e-13, 16: 1e-12}
    fel_alt = {1: 1e-14, 4: 1e-13, 9: 1e-12, 16: 1e-12}
    for n, mat in felippa.items():
        add(f"mat_10_bdb_axisym_qua4_felippa_{n}", "src/integ/mat_10_bdb.rs:549-676", "mat_10_bdb", "qua4", 2,
            pad_qua4(2.0, 1.0, 2.0, 1.0), mat, fel_tol[n], [n], dd=L(dd_pe), axisymmetric=True, young=96.0,
            poisson=1.0 / 3.0, two_dim=True, plane_stress=False, alt_tol=fel_alt[n])

    # ---- mat_01_nsn -----------------------------------------------------------------------------
    ana = AnaTri3(PAD_TRI3)
    add("mat_01_nsn_tri3", "src/integ/mat_01_nsn.rs:150-168", "mat_01_nsn", "tri3", 2, PAD_TRI3, ana.mat_01_nsn(12.0, 1.0),
        [8.34, 1e-14, 1e-14, 1e-14, 1e-12, 1e-13], [1, 3, 6, 7, 12, 16], s=12.0)
    add("mat_01_nsn_qua4", "src/integ/mat_01_nsn.rs:170-190", "mat_01_nsn", "qua4", 2, pad_qua4(2.0, 1.0, 2.0, 1.5),
        ana_qua4_mat_01_nsn(2.0, 1.5, 12.0, 1.0), [7.01, 1e-14, 1e-14, 1e-14], [1, 4, 9, 16], s=12.0)
    add("mat_01_nsn_qua8", "src/integ/mat_01_nsn.rs:192-209", "mat_01_nsn", "qua8", 2, pad_qua8(2.0, 1.0, 2.0, 1.5),
        ana_qua8_mat_01_nsn(2.0, 1.5, 3.0, 1.0), [1e-14, 1e-14], [9, 16], s=3.0)
    anat = AnaTet4(PAD_TET4)
    add("mat_01_nsn_tet4", "src/integ/mat_01_nsn.rs:211-229", "mat_01_nsn", "tet4", 3, PAD_TET4, anat.mat_01_nsn(3.0),
        [1e-15], [4], s=3.0)

    # ---- mat_03_btb -----------------------------------------------------------------------------
    kx, ky = 2.5, 3.8
    add("mat_03_btb_tri3", "src/integ/mat_03_btb.rs:180-200", "mat_03_btb", "tri3", 2, PAD_TRI3,
        ana.mat_03_btb(kx, ky, False), [1e-15, 1e-15, 1e-15], [1, 3, 7], tt=[kx, ky, 0.0, 0.0])
    add("mat_03_btb_qua4", "src/integ/mat_03_btb.rs:202-223", "mat_03_btb", "qua4", 2, pad_qua4(2.0, 1.0, 2.0, 1.5),
        ana_qua4_mat_03_btb(2.0, 1.5, kx, ky), [1e-15], [4], tt=[kx, ky, 0.0, 0.0])
    add("mat_03_btb_qua8", "src/integ/mat_03_btb.rs:225-246", "mat_03_btb", "qua8", 2, pad_qua8(2.0, 1.0, 2.0, 1.5),
        ana_qua8_mat_03_btb(2.0, 1.5, kx, ky), [1e-14, 1e-14], [9, 16], tt=[kx, ky, 0.0, 0.0])
    sig = np.array([[1.1, 1.2, 1.3], [1.2, 2.2, 2.3], [1.3, 2.3, 3.3]])
    sig_mandel = [1.1, 2.2, 3.3, 1.2 * SQRT_2, 2.3 * SQRT_2, 1.3 * SQRT_2]
    add("mat_03_btb_tet4", "src/integ/mat_03_btb.rs:248-277", "mat_03_btb", "tet4", 3, PAD_TET4, anat.mat_03_btb(sig),
        [1e-14], [4], tt=sig_mandel)
    # Lewis et al. example 5.6.1 (tests/test_integ_heat_axis_1.rs:57-91): K_lewis = K_cond + K_conv(side 0);
    # K_conv = AnalyticalTri3::mat_01_nsn_bry(0, h=1.2, axisym) = [[90,40],[40,70]] on element nodes (1,0),
    # so the conduction part (the mat_03_btb call) is K_lewis - K_conv:
    lewis_xx = [[15.0, 10.0], [25.0, 10.0], [20.0, 12.0]]
    k_cond = [[29.0, 21.0, -50.0], [21.0, 29.0, -50.0], [-50.0, -50.0, 100.0]]
    add("mat_03_btb_axisym_tri3_lewis", "tests/test_integ_heat_axis_1.rs:57-91", "mat_03_btb", "tri3", 2, lewis_xx, k_cond,
        [1e-13], [3], tt=[2.0, 2.0, 0.0, 0.0], axisymmetric=True)
    anal = AnaTri3(lewis_xx)
    add("mat_03_btb_axisym_tri3_analytical", "tests/test_integ_heat_axis_1.rs:57-70", "mat_03_btb", "tri3", 2, lewis_xx,
        anal.mat_03_btb(2.0, 2.0, True), [1e-13], [3], tt=[2.0, 2.0, 0.0, 0.0], axisymmetric=True)

    # ---- mat_02_bvn, mat_08_ntn, mat_09_nvb (single-pad cases of SURVEY.md 8(f4)) ----------------------
    ana3 = AnaTri3(PAD_TRI3)
    add("mat_02_bvn_tri3", "src/integ/mat_02_bvn.rs:170-191", "mat_02_bvn", "tri3", 2, PAD_TRI3, ana3.mat_02_bvn(2.0, 3.0),
        [1e-15], [3], v=[2.0, 3.0])
    add("mat_02_bvn_tet4", "src/integ/mat_02_bvn.rs:211-235", "mat_02_bvn", "tet4", 3, PAD_TET4, anat.mat_02_bvn(2.0, 3.0, 4.0),
        [1e-15], [4], v=[2.0, 3.0, 4.0])
    add("mat_08_ntn_tri3", "src/integ/mat_08_ntn.rs:185-209", "mat_08_ntn", "tri3", 2, PAD_TRI3, ana3.mat_08_ntn(2.7, 1.0),
        [1e-15, 1e-15], [3, 6], tt=[2.7, 2.7, 0.0, 0.0])
    add("mat_08_ntn_tet4", "src/integ/mat_08_ntn.rs:211-237", "mat_08_ntn", "tet4", 3, PAD_TET4, anat.mat_08_ntn(sig),
        [1e-15], [4], tt=sig_mandel)
    add("mat_09_nvb_tri3", "src/integ/mat_09_nvb.rs:170-194", "mat_09_nvb", "tri3", 2, PAD_TRI3, ana3.mat_09_nvb(2.0, 3.0),
        [1e-15], [3], v=[2.0, 3.0])
    add("mat_09_nvb_tet4", "src/integ/mat_09_nvb.rs:211-235", "mat_09_nvb", "tet4", 3, PAD_TET4, anat.mat_09_nvb(2.0, 3.0, 4.0),
        [1e-15], [4], v=[2.0, 3.0, 4.0])

    # ---- vec_01_ns ------------------------------------------------------------------------------
    add("vec_01_ns_doctest", "src/integ/vec_01_ns.rs:58-82", "vec_01_ns", "tri3", 2, [[2.0, 3.0], [6.0, 3.0], [2.0, 6.0]],
        [10.0, 10.0, 10.0], [1e-14], [0], s=5.0)
    add("vec_01_ns_tri3_constant", "src/integ/vec_01_ns.rs:192-215", "vec_01_ns", "tri3", 2, PAD_TRI3,
        ana.vec_01_ns(3.0, False), [1e-14, 1e-14, 1e-14, 1e-13, 1e-14], [1, 3, 4, 12, 16], s=3.0)
    add("vec_01_ns_tet4_constant", "src/integ/vec_01_ns.rs:217-243", "vec_01_ns", "tet4", 3, PAD_TET4, anat.vec_01_ns(120.0),
        [1e-13, 1e-13], [1, 4], s=120.0)
    add("vec_01_ns_tet4_linear_z", "src/integ/vec_01_ns.rs:245-279", "vec_01_ns", "tet4", 3, PAD_TET4,
        anat.vec_01_ns_linear_along_z(), [0.56, 1e-15, 1e-14, 1e-15, 1e-15, 1e-15, 1e-15], [1, 4, 5, 8, 14, 15, 24],
        s_field="x2")
    length = 6.0
    xa, xb = pad_lin2(length)[0][0], pad_lin2(length)[1][0]
    add("vec_01_ns_lin2_linear", "src/integ/vec_01_ns.rs:157-190", "vec_01_ns", "lin2", 2, pad_lin2(length),
        [length / 6.0 * (2 * xa + xb), length / 6.0 * (xa + 2 * xb)], [1e-15, 1e-14, 1e-14, 1e-14], [2, 3, 4, 5],
        s_field="x0")
    add("vec_01_ns_axisym_tri3_lewis", "tests/test_integ_heat_axis_1.rs:93-100", "vec_01_ns", "tri3", 2, lewis_xx,
        anal.vec_01_ns(1.2, True), [1e-13], [3], s=1.2, axisymmetric=True)

    # ---- vec_02_nv ------------------------------------------------------------------------------
    add("vec_02_nv_tri3_constant", "src/integ/vec_02_nv.rs:211-241", "vec_02_nv", "tri3", 2,
        PAD_TRI3, ana.vec_02_nv(-3.0, 8.0), [1e-14, 1e-14], [1, 3], v=[-3.0, 8.0])
    add("vec_02_nv_tet4_constant", "src/integ/vec_02_nv.rs:243-275", "vec_02_nv", "tet4", 3,
        PAD_TET4, anat.vec_02_nv(2.0, 3.0, 4.0), [1e-15, 1e-15], [1, 4], v=[2.0, 3.0, 4.0])
    cf = length / 6.0
    add("vec_02_nv_lin2_linear", "src/integ/vec_02_nv.rs:172-209", "vec_02_nv", "lin2", 2, pad_lin2(length),
        [cf * (2 * xa + xb), cf * (2 * xa + xb), cf * (xa + 2 * xb), cf * (xa + 2 * xb)], [1e-14, 1e-14], [2, 3],
        v_field="x0x0")

    # ---- vec_03_bv ------------------------------------------------------------------------------
    add("vec_03_bv_doctest", "src/integ/vec_03_bv.rs:50-83", "vec_03_bv", "tri3", 2, [[2.0, 3.0], [6.0, 3.0], [2.0, 6.0]],
        [-5.5, 1.5, 4.0], [1e-14], [0], v=[1.0, 2.0])
    add("vec_03_bv_tri3_constant", "src/integ/vec_03_bv.rs:171-199", "vec_03_bv", "tri3", 2, PAD_TRI3,
        ana.vec_03_bv(2.0, 3.0), [1e-14, 1e-14], [1, 3], v=[2.0, 3.0])
    add("vec_03_bv_tri3_bilinear", "src/integ/vec_03_bv.rs:201-230", "vec_03_bv", "tri3", 2, PAD_TRI3,
        ana.vec_03_bv_bilinear(), [1e-14, 1e-14], [1, 3], v_field="x")
    add("vec_03_bv_tet4_constant", "src/integ/vec_03_bv.rs:232-265", "vec_03_bv", "tet4", 3, PAD_TET4,
        anat.vec_03_bv(2.0, 3.0, 4.0), [1e-14] * 7, [1, 4, 5, 8, 14, 15, 24], v=[2.0, 3.0, 4.0])

    # ---- vec_04_bt ------------------------------------------------------------------------------
    t2d = [2.0, 3.0, 4.0, 5.0 * SQRT_2]  # Mandel vector of sigma = {s00=2, s11=3, s22=4, s01=5}
    add("vec_04_bt_tri3", "src/integ/vec_04_bt.rs:206-248", "vec_04_bt", "tri3", 2, PAD_TRI3,
        ana.vec_04_bt(t2d, False), [1e-14, 1e-14, 1e-14, 1e-13, 1e-14], [1, 3, 4, 12, 16], tt=t2d)
    add("vec_04_bt_axisym_tri3", "src/integ/vec_04_bt.rs:250-291", "vec_04_bt", "tri3", 2,
        PAD_TRI3, ana.vec_04_bt(t2d, True), [1e-13, 1e-13], [1, 3], tt=t2d, axisymmetric=True)
    sig4 = np.array([[2.0, 5.0, 7.0], [5.0, 3.0, 6.0], [7.0, 6.0, 4.0]])
    sig4_mandel = [2.0, 3.0, 4.0, 5.0 * SQRT_2, 6.0 * SQRT_2, 7.0 * SQRT_2]
    add("vec_04_bt_tet4", "src/integ/vec_04_bt.rs:293-331", "vec_04_bt", "tet4", 3, PAD_TET4,
        anat.vec_04_bt(sig4), [1e-14, 1e-14, 1e-13, 1e-14, 1e-13, 1e-13, 1e-13], [1, 4, 5, 8, 14, 15, 24], tt=sig4_mandel)

    # ---- scalar_field ---------------------------------------------------------------------------
    scalar = [
        dict(name="rotated_square", ref="src/integ/scalar_field.rs:66-150", kind="qua4", ndim=2,
             xx=[[0.0, 0.0], [1.0, -1.0], [2.0, 0.0], [1.0, 1.0]], ngauss=[1, 4, 9, 16], area=2.0, tol=1e-14),
        dict(name="slanted_hex8", ref="src/integ/scalar_field.rs:153-230", kind="hex8", ndim=3,
             xx=[[0, 0, 0], [1, 0, 0], [2, 1, 0], [1, 1, 0], [0, 0, 1], [1, 0, 1], [2, 1, 1], [1, 1, 1]],
             ngauss=[6, 8, 14, 27, 64], volume=1.0, int_x2=11.0 / 6.0, int_x3=2.0, tol=1e-14),
    ]

    # ---- Scratchpad doc-tests -------------------------------------------------------------------
    a = 3.0
    jac = dict(ref="src/shapes/scratchpad_calc_jacobian.rs:57-99, scratchpad_calc_gradient.rs:40-76", kind="qua4", ndim=2,
               xx=[[0.0, 0.0], [2 * a, 0.0], [2 * a, a], [0.0, a]], ksi=[0.0, 0.0], det=a * a / 2.0,
               jacobian=[[a, 0.0], [0.0, a / 2.0]],
               gradient=[[-1 / (4 * a), -1 / (2 * a)], [1 / (4 * a), -1 / (2 * a)], [1 / (4 * a), 1 / (2 * a)],
                         [-1 / (4 * a), 1 / (2 * a)]], tol=1e-15)

    out = dict(
        about="known-answer vectors transcribed from cpmech/gemlab v2.4.0 tests; see make_goldens.py",
        noise=1234.56,
        cases=cases,
        scalar_field=scalar,
        jacobian_doctest=jac,
    )
    (HERE / "known_answers.json").write_text(json.dumps(out, indent=1))
    print(f"wrote {len(cases)} cases")


if __name__ == "__main__":
    main()

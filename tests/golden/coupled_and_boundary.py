"""Known answers of the reference's tests for the two-pad coupling matrices and the boundary integrals, transcribed with file:line
(paths relative to the reference tree).  Imported by tests/test_oracle_coupled.py and tests/test_gpu_coupled.py."""
import numpy as np

SQRT_2 = np.sqrt(2.0)

# src/integ/testing.rs:32-65 (gen_pad_qua4 / gen_pad_qua8: a, b are half widths) and :68-83 (gen_pad_tet4)
def pad_qua8(a, b):
    return np.array([[-a, -b], [a, -b], [a, b], [-a, b], [0, -b], [a, 0], [0, b], [-a, 0]], dtype=float)


PAD_TET4 = np.array([[2, 3, 4], [6, 3, 2], [2, 5, 1], [4, 3, 6]], dtype=float)


def qua8_mat_04_nsb(a, b, s):  # src/integ/analytical_qua8.rs:61-72
    c = s / 18.0
    return np.array([
        [-7*b*c, -7*a*c, -b*c, -2*a*c, -2*b*c, -2*a*c, -2*b*c, -a*c, 8*b*c, -6*a*c, 6*b*c, 4*a*c, 4*b*c, 6*a*c, -6*b*c, 8*a*c],
        [b*c, -2*a*c, 7*b*c, -7*a*c, 2*b*c, -a*c, 2*b*c, -2*a*c, -8*b*c, -6*a*c, 6*b*c, 8*a*c, -4*b*c, 6*a*c, -6*b*c, 4*a*c],
        [2*b*c, 2*a*c, 2*b*c, a*c, 7*b*c, 7*a*c, b*c, 2*a*c, -4*b*c, -6*a*c, 6*b*c, -8*a*c, -8*b*c, 6*a*c, -6*b*c, -4*a*c],
        [-2*b*c, a*c, -2*b*c, 2*a*c, -b*c, 2*a*c, -7*b*c, 7*a*c, 4*b*c, -6*a*c, 6*b*c, -4*a*c, 8*b*c, 6*a*c, -6*b*c, -8*a*c]])


def qua8_mat_06_nvn(a, b, v0, v1):  # src/integ/analytical_qua8.rs:90-111
    c = a * b / 9.0
    rows = []
    pat = [[0, -1, -1, -1], [-1, 0, -1, -1], [-1, -1, 0, -1], [-1, -1, -1, 0], [4, 4, 2, 2], [2, 4, 4, 2], [2, 2, 4, 4], [4, 2, 2, 4]]
    for r in pat:
        rows.append([x * v0 * c for x in r])
        rows.append([x * v1 * c for x in r])
    return np.array(rows)


def qua8_mat_07_bsn(a, b, s):  # src/integ/analytical_qua8.rs:114-136
    c = s / 18.0
    return np.array([
        [-7*b*c, b*c, 2*b*c, -2*b*c], [-7*a*c, -2*a*c, 2*a*c, a*c], [-b*c, 7*b*c, 2*b*c, -2*b*c], [-2*a*c, -7*a*c, a*c, 2*a*c],
        [-2*b*c, 2*b*c, 7*b*c, -b*c], [-2*a*c, -a*c, 7*a*c, 2*a*c], [-2*b*c, 2*b*c, b*c, -7*b*c], [-a*c, -2*a*c, 2*a*c, 7*a*c],
        [8*b*c, -8*b*c, -4*b*c, 4*b*c], [-6*a*c, -6*a*c, -6*a*c, -6*a*c], [6*b*c, 6*b*c, 6*b*c, 6*b*c], [4*a*c, 8*a*c, -8*a*c, -4*a*c],
        [4*b*c, -4*b*c, -8*b*c, 8*b*c], [6*a*c, 6*a*c, 6*a*c, 6*a*c], [-6*b*c, -6*b*c, -6*b*c, -6*b*c], [8*a*c, 4*a*c, -4*a*c, -8*a*c]])


def qua8_mat_05_btn(a, b, t00, t01, t10, t11):  # src/integ/analytical_qua8.rs:75-87
    def row(c0, c1):  # every entry pairs (coefficient of b t0j, coefficient of a t1j) with a divisor
        out = []
        for (cb, ca, div) in zip(c0, c1, [18, 18, 9, 18, 9, 9, 9, 9]):
            out += [(cb * b * t00 + ca * a * t10) / div, (cb * b * t01 + ca * a * t11) / div]
        return out
    return np.array([
        row([1, 1, 1, 2, -4, -3, -2, -3], [1, 2, 1, 1, -3, -2, -3, -4]),
        [(-(b*t00) + 2*a*t10)/18, (-(b*t01) + 2*a*t11)/18, (-(b*t00) + a*t10)/18, (-(b*t01) + a*t11)/18, (-2*b*t00 + a*t10)/18, (-2*b*t01 + a*t11)/18,
         (-(b*t00) + a*t10)/9, (-(b*t01) + a*t11)/9, (4*b*t00 - 3*a*t10)/9, (4*b*t01 - 3*a*t11)/9, (3*b*t00 - 4*a*t10)/9, (3*b*t01 - 4*a*t11)/9,
         (2*b*t00 - 3*a*t10)/9, (2*b*t01 - 3*a*t11)/9, (3*b*t00 - 2*a*t10)/9, (3*b*t01 - 2*a*t11)/9],
        [(-(b*t00) - a*t10)/9, (-(b*t01) - a*t11)/9, (-2*b*t00 - a*t10)/18, (-2*b*t01 - a*t11)/18, (-(b*t00) - a*t10)/18, (-(b*t01) - a*t11)/18,
         (-(b*t00) - 2*a*t10)/18, (-(b*t01) - 2*a*t11)/18, (2*b*t00 + 3*a*t10)/9, (2*b*t01 + 3*a*t11)/9, (3*b*t00 + 4*a*t10)/9, (3*b*t01 + 4*a*t11)/9,
         (4*b*t00 + 3*a*t10)/9, (4*b*t01 + 3*a*t11)/9, (3*b*t00 + 2*a*t10)/9, (3*b*t01 + 2*a*t11)/9],
        [(2*b*t00 - a*t10)/18, (2*b*t01 - a*t11)/18, (b*t00 - a*t10)/9, (b*t01 - a*t11)/9, (b*t00 - 2*a*t10)/18, (b*t01 - 2*a*t11)/18,
         (b*t00 - a*t10)/18, (b*t01 - a*t11)/18, (-2*b*t00 + 3*a*t10)/9, (-2*b*t01 + 3*a*t11)/9, (-3*b*t00 + 2*a*t10)/9, (-3*b*t01 + 2*a*t11)/9,
         (-4*b*t00 + 3*a*t10)/9, (-4*b*t01 + 3*a*t11)/9, (-3*b*t00 + 4*a*t10)/9, (-3*b*t01 + 4*a*t11)/9]])


class AnaTet4:
    """volume and gradients of src/integ/analytical_tet4.rs:31-139 (constant B of the linear tetrahedron)"""

    def __init__(self, xx):
        xx = np.asarray(xx, dtype=float)
        jac = (xx[1:] - xx[0]).T  # columns dx/dksi
        self.volume = np.linalg.det(jac) / 6.0
        ll = np.array([[-1, -1, -1], [1, 0, 0], [0, 1, 0], [0, 0, 1]], dtype=float)
        self.bb = ll @ np.linalg.inv(jac)

    def mat_04_nsb(self, s):  # :298-311: K[m][3n+j] = (V/4) B[n][j] s
        return np.tile((self.volume / 4.0 * s * self.bb).reshape(1, 12), (4, 1))

    def mat_05_btn(self, t):  # :314-331: K[m][3n+j] = (V/4) sum_a B[m][a] T[a][j]
        return np.tile(self.volume / 4.0 * (self.bb @ t), (1, 4))

    def mat_06_nvn(self, v):  # :334-351: K[3m+i][n] = (V/20) v_i (2 if m == n else 1)
        w = np.ones((4, 4)) + np.eye(4)
        return self.volume / 20.0 * np.kron(w, np.asarray(v, dtype=float).reshape(3, 1))

    def mat_07_bsn(self, s):  # :354-374: K[3m+i][n] = (V/4) B[m][i] s
        return np.tile((self.volume / 4.0 * s * self.bb).reshape(12, 1), (1, 4))


def mandel(t):
    """Tensor2::vector() of a symmetric matrix: (00, 11, 22, 01 sqrt2, 12 sqrt2, 02 sqrt2); 2D keeps the first four"""
    t = np.asarray(t, dtype=float)
    return np.array([t[0, 0], t[1, 1], t[2, 2], t[0, 1] * SQRT_2, t[1, 2] * SQRT_2, t[0, 2] * SQRT_2])


def coupled_cases():
    """(name, ref, op, kind, kind_b, ndim, coords, coefficient record, expected matrix [row][col], rules, tolerance)"""
    a, b = 2.0, 3.0
    tt2 = np.array([[2.0, 5.0, 0.0], [5.0, 3.0, 0.0], [0.0, 0.0, 4.0]])
    tt3 = np.array([[1.0, 4.0, 6.0], [4.0, 2.0, 5.0], [6.0, 5.0, 3.0]])
    ana = AnaTet4(PAD_TET4)
    q8 = pad_qua8(a, b)
    return [
        ("mat_04_nsb_qua4_qua8", "src/integ/mat_04_nsb.rs:170-190", "mat_04_nsb", "qua8", "qua4", 2, q8, [9.0], qua8_mat_04_nsb(a, b, 9.0), [4, 9], 1e-14),
        ("mat_04_nsb_tet4_tet4", "src/integ/mat_04_nsb.rs:192-211", "mat_04_nsb", "tet4", "tet4", 3, PAD_TET4, [9.0], ana.mat_04_nsb(9.0), [4], 1e-14),
        ("mat_05_btn_qua4_qua8", "src/integ/mat_05_btn.rs:qua4_qua8_works", "mat_05_btn", "qua8", "qua4", 2, q8, mandel(tt2)[:4],
         qua8_mat_05_btn(a, b, 2.0, 5.0, 5.0, 3.0), [4, 9], 1e-14),
        ("mat_05_btn_tet4_tet4", "src/integ/mat_05_btn.rs:tet4_tet4_works", "mat_05_btn", "tet4", "tet4", 3, PAD_TET4, mandel(tt3), ana.mat_05_btn(tt3), [4], 1e-14),
        ("mat_06_nvn_qua4_qua8", "src/integ/mat_06_nvn.rs:qua4_qua8_works", "mat_06_nvn", "qua8", "qua4", 2, q8, [4.0, 5.0], qua8_mat_06_nvn(a, b, 4.0, 5.0), [4, 9], 1e-14),
        ("mat_06_nvn_tet4_tet4", "src/integ/mat_06_nvn.rs:tet4_tet4_works", "mat_06_nvn", "tet4", "tet4", 3, PAD_TET4, [4.0, 5.0, 6.0], ana.mat_06_nvn([4.0, 5.0, 6.0]), [4], 1e-15),
        ("mat_07_bsn_qua4_qua8", "src/integ/mat_07_bsn.rs:qua4_qua8_works", "mat_07_bsn", "qua8", "qua4", 2, q8, [9.0], qua8_mat_07_bsn(a, b, 9.0), [4, 9], 1e-14),
        ("mat_07_bsn_tet4_tet4", "src/integ/mat_07_bsn.rs:tet4_tet4_works", "mat_07_bsn", "tet4", "tet4", 3, PAD_TET4, [9.0], ana.mat_07_bsn(9.0), [4], 1e-14),
    ]


def boundary_cases():
    """(name, ref, op, kind, ndim, coords, record, expected, rule, tolerance); records: (s0, w..) with s = s0 + w.un, and for
    vec_02_nv_bry (v.., qn) with v = v + qn un"""
    tri = np.array([[3.0, 4.0], [8.0, 4.0], [5.0, 9.0]])  # gen_pad_tri3, src/integ/testing.rs:20-29
    out = []
    for side, (i, j) in enumerate([(1, 0), (2, 1), (0, 2)]):  # Tri3 edge_node_id: side 0 = (1, 0), 1 = (2, 1), 2 = (0, 2)  (src/shapes/tri3.rs)
        xe = tri[[i, j]]
        ll = np.linalg.norm(xe[1] - xe[0])
        c = 12.0 * ll / 6.0
        out.append((f"mat_01_nsn_bry_tri3_side{side}", "src/integ/mat_01_nsn_bry.rs:184-207, analytical_tri3.rs:232-249", "mat_01_nsn_bry", "lin2", 2, xe,
                    [12.0, 0.0, 0.0], np.array([[2 * c, c], [c, 2 * c]]), 2, 1e-14))
    xa, ya, xb, yb = 1.0, 2.0, 4.0, 6.0
    w0, w1 = 3.0, 5.0
    dx, dy = xb - xa, yb - ya
    out.append(("vec_01_ns_bry_lin2_constant_w", "src/integ/vec_01_ns_bry.rs:164-191", "vec_01_ns_bry", "lin2", 2, np.array([[xa, ya], [xb, yb]]),
                [0.0, w0, w1], np.array([(w1 * dx - w0 * dy) / 2.0] * 2), 2, 1e-15))
    ll = 4.0
    out.append(("vec_02_nv_bry_lin2_uniform_load", "src/integ/vec_02_nv_bry.rs:179-195", "vec_02_nv_bry", "lin2", 2, np.array([[0.0, 0.0], [ll, 0.0]]),
                [0.0, -1.0, 0.0], np.array([0.0, -2.0, 0.0, -2.0]), 0, 1e-15))
    ll = 3.0
    out.append(("vec_02_nv_bry_lin3_uniform_load", "src/integ/vec_02_nv_bry.rs:212-232", "vec_02_nv_bry", "lin3", 2, np.array([[0.0, 0.0], [ll, 0.0], [ll / 2, 0.0]]),
                [0.0, -1.0, 0.0], np.array([0.0, -0.5, 0.0, -0.5, 0.0, -2.0]), 0, 1e-15))
    return out

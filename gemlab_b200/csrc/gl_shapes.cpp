// gl_shapes.cpp -- host-side reference-element evaluation and Gauss-rule lookup of libgemlab_cuda.
//
// On the device N(ksi) and L = dN/dksi are never evaluated per cell: they depend only on (GeoKind, rule), so the
// library tabulates them ONCE per (kind, rule) here and ships the table to the GPU.  The reference instead calls
// `(pad.fn_interp)` / `(pad.fn_deriv)` at every Gauss point of every cell (src/integ/mat_10_bdb.rs:153-154 via
// GeoKind::functions, src/shapes/enums.rs:1124-1152).
//
// The formulas are written family-wise from the node reference coordinates (serendipity / Lagrange products for
// Qua/Hex, barycentric polynomials for Tri/Tet) rather than node by node as in src/shapes/{hex8,hex20,...}.rs; the node
// ORDER and reference coordinates are the reference's (NODE_REFERENCE_COORDS in each src/shapes/<kind>.rs), because the
// order is part of the connectivity contract.  tests/test_shapes_tables.py checks these against the oracle's
// node-by-node restatement.
#include <cmath>
#include <mutex>
#include <vector>

#include "gl_internal.hpp"

#include <cmath>
#include <cstring>

namespace gl {

#include "gauss_tables.inc"
#include "shape_tables.inc"
#include "ref_coords.inc"

namespace {

const int NNODE[GL_NKIND] = {2, 3, 4, 5, 3, 6, 10, 15, 4, 8, 9, 12, 16, 17, 4, 10, 20, 8, 20, 32};
const int GNDIM[GL_NKIND] = {1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3};
const int KCLASS[GL_NKIND] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4};
const char *NAMES[GL_NKIND] = {"lin2", "lin3", "lin4", "lin5", "tri3", "tri6", "tri10", "tri15", "qua4", "qua8",
                               "qua9", "qua12", "qua16", "qua17", "tet4", "tet10", "tet20", "hex8", "hex20", "hex32"};

// node reference coordinates of the tensor-product families (corner, then mid-side, then centre nodes)
const double QUA_XI[9][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {0, 0}};
const double HEX_XI[20][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1},
                              {-1, 1, 1},   {0, -1, -1}, {1, 0, -1}, {0, 1, -1},  {-1, 0, -1}, {0, -1, 1}, {1, 0, 1},
                              {0, 1, 1},    {-1, 0, 1},  {-1, -1, 0}, {1, -1, 0}, {1, 1, 0},  {-1, 1, 0}};

// 1D quadratic Lagrange basis on {-1, 0, +1} and its derivative, selected by the node coordinate a
inline double lag2(double a, double x) { return a == 0.0 ? 1.0 - x * x : 0.5 * x * (x + a); }
inline double dlag2(double a, double x) { return a == 0.0 ? -2.0 * x : x + 0.5 * a; }

void eval_lin(int nnode, const double *ksi, double *nn, double *ll) {
    const double r = ksi[0];
    if (nnode == 2) {
        nn[0] = 0.5 * (1.0 - r);
        nn[1] = 0.5 * (1.0 + r);
        ll[0] = -0.5;
        ll[1] = 0.5;
    } else { // lin3: nodes at -1, +1, 0
        const double a[3] = {-1.0, 1.0, 0.0};
        for (int m = 0; m < 3; m++) {
            nn[m] = lag2(a[m], r);
            ll[m] = dlag2(a[m], r);
        }
    }
}

// barycentric quadratic family: vertices l(2l-1), edge mid-nodes 4 l_a l_b
void eval_simplex(int gd, int nnode, const double *ksi, double *nn, double *ll) {
    // barycentric coordinates lam[0] = 1 - sum(ksi), lam[1..gd] = ksi; d lam[0]/d ksi_j = -1, d lam[a]/d ksi_j = delta
    double lam[4];
    lam[0] = 1.0;
    for (int a = 0; a < gd; a++) {
        lam[a + 1] = ksi[a];
        lam[0] -= ksi[a];
    }
    auto dlam = [&](int a, int j) { return a == 0 ? -1.0 : (a - 1 == j ? 1.0 : 0.0); };
    const int nv = gd + 1;
    if (nnode == nv) { // linear
        for (int a = 0; a < nv; a++) {
            nn[a] = lam[a];
            for (int j = 0; j < gd; j++) ll[a * gd + j] = dlam(a, j);
        }
        return;
    }
    // quadratic: vertex nodes first
    for (int a = 0; a < nv; a++) {
        nn[a] = lam[a] * (2.0 * lam[a] - 1.0);
        for (int j = 0; j < gd; j++) ll[a * gd + j] = (4.0 * lam[a] - 1.0) * dlam(a, j);
    }
    // edge nodes in the reference's order: tri6 (0-1, 1-2, 2-0); tet10 (0-1, 1-2, 2-0, 0-3, 1-3, 2-3)
    static const int TRI_E[3][2] = {{0, 1}, {1, 2}, {2, 0}};
    static const int TET_E[6][2] = {{0, 1}, {1, 2}, {2, 0}, {0, 3}, {1, 3}, {2, 3}};
    const int ne = (gd == 2) ? 3 : 6;
    for (int e = 0; e < ne; e++) {
        const int a = (gd == 2) ? TRI_E[e][0] : TET_E[e][0];
        const int b = (gd == 2) ? TRI_E[e][1] : TET_E[e][1];
        const int m = nv + e;
        nn[m] = 4.0 * lam[a] * lam[b];
        for (int j = 0; j < gd; j++) ll[m * gd + j] = 4.0 * (dlam(a, j) * lam[b] + lam[a] * dlam(b, j));
    }
}

void eval_qua(int nnode, const double *ksi, double *nn, double *ll) {
    const double r = ksi[0], s = ksi[1];
    for (int m = 0; m < nnode; m++) {
        const double a = QUA_XI[m][0], b = QUA_XI[m][1];
        double n, dr, ds;
        if (nnode == 9) { // Lagrange product
            n = lag2(a, r) * lag2(b, s);
            dr = dlag2(a, r) * lag2(b, s);
            ds = lag2(a, r) * dlag2(b, s);
        } else if (nnode == 4) { // bilinear
            n = 0.25 * (1.0 + a * r) * (1.0 + b * s);
            dr = 0.25 * a * (1.0 + b * s);
            ds = 0.25 * b * (1.0 + a * r);
        } else if (m < 4) { // serendipity corner
            const double fr = 1.0 + a * r, fs = 1.0 + b * s, g = a * r + b * s - 1.0;
            n = 0.25 * fr * fs * g;
            dr = 0.25 * a * fs * (g + fr);
            ds = 0.25 * b * fr * (g + fs);
        } else if (a == 0.0) { // mid-side on an s = +-1 edge
            n = 0.5 * (1.0 - r * r) * (1.0 + b * s);
            dr = -r * (1.0 + b * s);
            ds = 0.5 * b * (1.0 - r * r);
        } else { // mid-side on an r = +-1 edge
            n = 0.5 * (1.0 + a * r) * (1.0 - s * s);
            dr = 0.5 * a * (1.0 - s * s);
            ds = -s * (1.0 + a * r);
        }
        nn[m] = n;
        ll[m * 2 + 0] = dr;
        ll[m * 2 + 1] = ds;
    }
}

void eval_hex(int nnode, const double *ksi, double *nn, double *ll) {
    const double r = ksi[0], s = ksi[1], t = ksi[2];
    for (int m = 0; m < nnode; m++) {
        const double a = HEX_XI[m][0], b = HEX_XI[m][1], c = HEX_XI[m][2];
        const double fr = 1.0 + a * r, fs = 1.0 + b * s, ft = 1.0 + c * t;
        double n, dr, ds, dt;
        if (nnode == 8) { // trilinear
            n = 0.125 * fr * fs * ft;
            dr = 0.125 * a * fs * ft;
            ds = 0.125 * b * fr * ft;
            dt = 0.125 * c * fr * fs;
        } else if (m < 8) { // serendipity corner
            const double g = a * r + b * s + c * t - 2.0;
            n = 0.125 * fr * fs * ft * g;
            dr = 0.125 * a * fs * ft * (g + fr);
            ds = 0.125 * b * fr * ft * (g + fs);
            dt = 0.125 * c * fr * fs * (g + ft);
        } else if (a == 0.0) { // edge parallel to r
            n = 0.25 * (1.0 - r * r) * fs * ft;
            dr = -0.5 * r * fs * ft;
            ds = 0.25 * b * (1.0 - r * r) * ft;
            dt = 0.25 * c * (1.0 - r * r) * fs;
        } else if (b == 0.0) { // edge parallel to s
            n = 0.25 * fr * (1.0 - s * s) * ft;
            dr = 0.25 * a * (1.0 - s * s) * ft;
            ds = -0.5 * s * fr * ft;
            dt = 0.25 * c * fr * (1.0 - s * s);
        } else { // edge parallel to t
            n = 0.25 * fr * fs * (1.0 - t * t);
            dr = 0.25 * a * fs * (1.0 - t * t);
            ds = 0.25 * b * fr * (1.0 - t * t);
            dt = -0.5 * t * fr * fs;
        }
        nn[m] = n;
        ll[m * 3 + 0] = dr;
        ll[m * 3 + 1] = ds;
        ll[m * 3 + 2] = dt;
    }
}

} // namespace

// N | L of a (kind, rule) pair, bit-identical to the reference's expressions at its Gauss points (generated data, see
// tools/gen_shape_tables.py); nullptr when the pair is not tabulated (the caller then evaluates shape_eval)
const double *shape_table(int kind, int ng) {
    for (int i = 0; i < N_SHAPE_TABLES; i++)
        if (SHAPE_TABLES[i].kind == kind && SHAPE_TABLES[i].ng == ng) return SHAPE_TABLES[i].data;
    return nullptr;
}

int kind_nnode(int kind) { return (kind >= 0 && kind < GL_NKIND) ? NNODE[kind] : -1; }
int kind_geo_ndim(int kind) { return (kind >= 0 && kind < GL_NKIND) ? GNDIM[kind] : -1; }
int kind_class(int kind) { return (kind >= 0 && kind < GL_NKIND) ? KCLASS[kind] : -1; }
const char *kind_name(int kind) { return (kind >= 0 && kind < GL_NKIND) ? NAMES[kind] : ""; }
int kind_from_name(const char *name) {
    for (int k = 0; k < GL_NKIND; k++)
        if (std::strcmp(name, NAMES[k]) == 0) return k;
    return -1;
}
// every GeoKind integrates: the kernels read N and L from the (kind, rule) tables of shape_tables.inc
bool kind_supported(int kind) { return kind >= 0 && kind < GL_NKIND; }

// kinds with closed family-wise formulas (eval_lin / eval_simplex / eval_qua / eval_hex above); the cubic and quartic kinds go
// through their NODAL BASIS, built from the node coordinates alone (below)
bool kind_has_formulas(int kind) {
    switch (kind) {
    case GL_LIN2: case GL_LIN3: case GL_TRI3: case GL_TRI6: case GL_QUA4: case GL_QUA8: case GL_QUA9:
    case GL_TET4: case GL_TET10: case GL_HEX8: case GL_HEX20:
        return true;
    default:
        return false;
    }
}

// GeoKind::reference_coords (src/shapes/enums.rs:1049): natural coordinates of node m
int kind_reference_coords(int kind, int m, double *ksi) {
    if (kind < 0 || kind >= GL_NKIND || m < 0 || m >= NNODE[kind]) return GL_ERR_KIND_UNSUPPORTED;
    for (int j = 0; j < GNDIM[kind]; j++) ksi[j] = REF_COORDS[REF_COORDS_OFFSET[kind] + m][j];
    return GL_OK;
}

namespace {

// Shape functions of Lin4/5, Tri10/15, Qua12/16/17, Tet20 and Hex32 at an ARBITRARY point.  A nodal finite element is fixed by
// its nodes and its polynomial space: N_m is the member of the space that is 1 at node m and 0 at the others, i.e.
// N(ksi) = C phi(ksi) with C = inv(Phi), Phi[k][m] = phi_k(node m), for any basis phi of the space.  The spaces are the
// classical ones -- complete polynomials P_k on simplices and lines, the tensor space Q_3 for Qua16, the serendipity spaces
// S_k (monomials of superlinear degree <= k) for Qua12, Qua17 and Hex32 -- and the result agrees with the reference's
// node-by-node expressions (src/shapes/tri10.rs:94-120, hex32.rs:216-330, ...) to 1e-14 at every Gauss point of every rule
// (tests/test_shape_tables.py), without restating them.  Built once per kind, in long double.
struct NodalBasis {
    int n = 0, gd = 0;
    std::vector<int> expo;       // [n][3] monomial exponents
    std::vector<long double> cc; // [n][n]: N_m = sum_k cc[m * n + k] phi_k
    bool ok = false;
};

int superlinear_degree(int a, int b, int c) { return (a >= 2 ? a : 0) + (b >= 2 ? b : 0) + (c >= 2 ? c : 0); }

void monomials_of(int kind, std::vector<int> &e) {
    auto push = [&](int a, int b, int c) { e.push_back(a); e.push_back(b); e.push_back(c); };
    switch (kind) {
    case GL_LIN4: for (int a = 0; a < 4; a++) push(a, 0, 0); break;
    case GL_LIN5: for (int a = 0; a < 5; a++) push(a, 0, 0); break;
    case GL_TRI10: case GL_TRI15: {
        const int k = kind == GL_TRI10 ? 3 : 4;
        for (int a = 0; a <= k; a++)
            for (int b = 0; a + b <= k; b++) push(a, b, 0);
    } break;
    case GL_QUA16:
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++) push(a, b, 0);
        break;
    case GL_QUA12: case GL_QUA17: {
        const int k = kind == GL_QUA12 ? 3 : 4;
        for (int a = 0; a <= k; a++)
            for (int b = 0; a + b <= k; b++) push(a, b, 0);
        push(k, 1, 0);
        push(1, k, 0);
    } break;
    case GL_TET20:
        for (int a = 0; a <= 3; a++)
            for (int b = 0; a + b <= 3; b++)
                for (int c = 0; a + b + c <= 3; c++) push(a, b, c);
        break;
    case GL_HEX32:
        for (int a = 0; a < 4; a++)
            for (int b = 0; b < 4; b++)
                for (int c = 0; c < 4; c++)
                    if (superlinear_degree(a, b, c) <= 3) push(a, b, c);
        break;
    default: break;
    }
}

long double ipow(long double x, int k) {
    long double r = 1.0L;
    for (int i = 0; i < k; i++) r *= x;
    return r;
}

// inverse of an n x n matrix (row-major) by Gauss-Jordan with partial pivoting; false if singular
template <typename T>
bool invert_dense(std::vector<T> &a, int n) {
    std::vector<T> inv((size_t)n * n, (T)0);
    for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = (T)1;
    for (int col = 0; col < n; col++) {
        int piv = col;
        for (int r = col + 1; r < n; r++)
            if (std::fabs((double)a[(size_t)r * n + col]) > std::fabs((double)a[(size_t)piv * n + col])) piv = r;
        if (std::fabs((double)a[(size_t)piv * n + col]) < 1e-13) return false;
        if (piv != col)
            for (int k = 0; k < n; k++) {
                std::swap(a[(size_t)piv * n + k], a[(size_t)col * n + k]);
                std::swap(inv[(size_t)piv * n + k], inv[(size_t)col * n + k]);
            }
        const T d = a[(size_t)col * n + col];
        for (int k = 0; k < n; k++) {
            a[(size_t)col * n + k] /= d;
            inv[(size_t)col * n + k] /= d;
        }
        for (int r = 0; r < n; r++) {
            if (r == col) continue;
            const T f = a[(size_t)r * n + col];
            if (f == (T)0) continue;
            for (int k = 0; k < n; k++) {
                a[(size_t)r * n + k] -= f * a[(size_t)col * n + k];
                inv[(size_t)r * n + k] -= f * inv[(size_t)col * n + k];
            }
        }
    }
    a.swap(inv);
    return true;
}

const NodalBasis &nodal_basis(int kind) {
    static NodalBasis table[GL_NKIND];
    static std::once_flag once[GL_NKIND];
    std::call_once(once[kind], [kind]() {
        NodalBasis &nb = table[kind];
        nb.n = NNODE[kind];
        nb.gd = GNDIM[kind];
        monomials_of(kind, nb.expo);
        if ((int)nb.expo.size() != 3 * nb.n) return;
        const int n = nb.n;
        std::vector<long double> phi((size_t)n * n); // [k][m] = phi_k(node m)
        for (int k = 0; k < n; k++)
            for (int m = 0; m < n; m++) {
                const double *x = REF_COORDS[REF_COORDS_OFFSET[kind] + m];
                phi[(size_t)k * n + m] = ipow(x[0], nb.expo[3 * k]) * ipow(x[1], nb.expo[3 * k + 1]) * ipow(x[2], nb.expo[3 * k + 2]);
            }
        if (!invert_dense(phi, n)) return; // phi := inv(Phi): [m][k]
        nb.cc.swap(phi);
        nb.ok = true;
    });
    return table[kind];
}

int eval_nodal(int kind, const double *ksi, double *nn, double *ll) {
    const NodalBasis &nb = nodal_basis(kind);
    if (!nb.ok) return GL_ERR_KIND_UNSUPPORTED;
    const int n = nb.n, gd = nb.gd;
    long double x[3] = {ksi[0], gd > 1 ? ksi[1] : 0.0, gd > 2 ? ksi[2] : 0.0};
    std::vector<long double> phi(n), dphi((size_t)n * 3, 0.0L);
    for (int k = 0; k < n; k++) {
        const int *e = &nb.expo[3 * k];
        phi[k] = ipow(x[0], e[0]) * ipow(x[1], e[1]) * ipow(x[2], e[2]);
        for (int j = 0; j < gd; j++) {
            if (e[j] == 0) continue;
            long double v = (long double)e[j];
            for (int i = 0; i < 3; i++) v *= ipow(x[i], i == j ? e[i] - 1 : e[i]);
            dphi[(size_t)k * 3 + j] = v;
        }
    }
    for (int m = 0; m < n; m++) {
        long double s = 0.0L, d[3] = {0.0L, 0.0L, 0.0L};
        for (int k = 0; k < n; k++) {
            const long double c = nb.cc[(size_t)m * n + k];
            s += c * phi[k];
            for (int j = 0; j < gd; j++) d[j] += c * dphi[(size_t)k * 3 + j];
        }
        nn[m] = (double)s;
        for (int j = 0; j < gd; j++) ll[m * gd + j] = (double)d[j];
    }
    return GL_OK;
}

} // namespace

int shape_eval(int kind, const double *ksi, double *nn, double *ll) {
    if (kind < 0 || kind >= GL_NKIND) return GL_ERR_KIND_UNSUPPORTED;
    if (!kind_has_formulas(kind)) return eval_nodal(kind, ksi, nn, ll);
    const int nnode = NNODE[kind];
    switch (KCLASS[kind]) {
    case 0: eval_lin(nnode, ksi, nn, ll); break;
    case 1: eval_simplex(2, nnode, ksi, nn, ll); break;
    case 2: eval_qua(nnode, ksi, nn, ll); break;
    case 3: eval_simplex(3, nnode, ksi, nn, ll); break;
    default: eval_hex(nnode, ksi, nn, ll); break;
    }
    return GL_OK;
}

// recovery::get_extrap_matrix (src/recovery/get_extrap_matrix.rs:129-200): E (nnode x npoint, COLUMN-major in `ee`) maps values at
// the Gauss points of the rule to nodal values.  P[p][m] = N_m(ksi_p) (get_interp_matrix.rs):  npoint = nnode: E = inv(P);
// npoint > nnode or npoint = 1: E = pinv(P);  otherwise E = hat_xi_nodes B + pinv(P) (I - hat_xi B) with B = pinv(hat_xi), hat_xi =
// [ksi_p, 1] (the constant tables TR_PINV_HXI_* of the reference are B^T).
int extrap_matrix(int kind, int ngauss, double *ee) {
    if (kind < 0 || kind >= GL_NKIND) return GL_ERR_KIND_UNSUPPORTED;
    int np = 0;
    const double *g = nullptr;
    int st = gauss_rule(kind, ngauss, &np, &g);
    if (st) return st;
    const int nv = NNODE[kind], gd = GNDIM[kind], nh = gd + 1;
    std::vector<long double> pp((size_t)np * nv); // [p][m]
    {
        std::vector<double> nn(nv), ll((size_t)nv * gd);
        for (int p = 0; p < np; p++) {
            st = shape_eval(kind, g + 4 * p, nn.data(), ll.data());
            if (st) return st;
            for (int m = 0; m < nv; m++) pp[(size_t)p * nv + m] = nn[m];
        }
    }
    // Moore-Penrose pseudo-inverse of an r x c matrix a (row-major) -> c x r, by one-sided (Hestenes) Jacobi SVD in long double;
    // singular values below eps max(r, c) sigma_max count as zero (the sampled nodal basis is rank deficient for some rules,
    // e.g. Qua12 at 3 x 3 points), as LAPACK-based pseudo-inverses do
    auto pinv = [](const std::vector<long double> &a_in, int r_in, int c_in, std::vector<long double> &out) -> bool {
        const bool tr = r_in < c_in; // work on the tall orientation
        const int r = tr ? c_in : r_in, c = tr ? r_in : c_in;
        std::vector<long double> a((size_t)r * c), v((size_t)c * c, 0.0L);
        for (int i = 0; i < r; i++)
            for (int j = 0; j < c; j++) a[(size_t)i * c + j] = tr ? a_in[(size_t)j * c_in + i] : a_in[(size_t)i * c_in + j];
        for (int j = 0; j < c; j++) v[(size_t)j * c + j] = 1.0L;
        for (int sweep = 0; sweep < 60; sweep++) {
            long double off = 0.0L;
            for (int p = 0; p < c - 1; p++)
                for (int q = p + 1; q < c; q++) {
                    long double alpha = 0.0L, beta = 0.0L, gamma = 0.0L;
                    for (int i = 0; i < r; i++) {
                        alpha += a[(size_t)i * c + p] * a[(size_t)i * c + p];
                        beta += a[(size_t)i * c + q] * a[(size_t)i * c + q];
                        gamma += a[(size_t)i * c + p] * a[(size_t)i * c + q];
                    }
                    if (gamma == 0.0L || alpha == 0.0L || beta == 0.0L) continue;
                    off = std::max(off, fabsl(gamma) / sqrtl(alpha * beta));
                    const long double zeta = (beta - alpha) / (2.0L * gamma);
                    const long double t = (zeta >= 0.0L ? 1.0L : -1.0L) / (fabsl(zeta) + sqrtl(1.0L + zeta * zeta));
                    const long double cs = 1.0L / sqrtl(1.0L + t * t), sn = cs * t;
                    for (int i = 0; i < r; i++) {
                        const long double x = a[(size_t)i * c + p], y = a[(size_t)i * c + q];
                        a[(size_t)i * c + p] = cs * x - sn * y;
                        a[(size_t)i * c + q] = sn * x + cs * y;
                    }
                    for (int i = 0; i < c; i++) {
                        const long double x = v[(size_t)i * c + p], y = v[(size_t)i * c + q];
                        v[(size_t)i * c + p] = cs * x - sn * y;
                        v[(size_t)i * c + q] = sn * x + cs * y;
                    }
                }
            if (off < 1e-19L) break;
        }
        std::vector<long double> sig(c);
        long double smax = 0.0L;
        for (int j = 0; j < c; j++) {
            long double n2 = 0.0L;
            for (int i = 0; i < r; i++) n2 += a[(size_t)i * c + j] * a[(size_t)i * c + j];
            sig[j] = sqrtl(n2);
            smax = std::max(smax, sig[j]);
        }
        const long double cutoff = 2.220446049250313e-16L * (long double)std::max(r, c) * smax;
        // pinv(A) = V diag(1 / sigma) U^T with U = A_rotated / sigma: (c x r); entry (i, k) = sum_j v[i][j] a[k][j] / sigma_j^2
        std::vector<long double> pi((size_t)c * r, 0.0L);
        for (int j = 0; j < c; j++) {
            if (sig[j] <= cutoff) continue;
            const long double w = 1.0L / (sig[j] * sig[j]);
            for (int i = 0; i < c; i++)
                for (int k = 0; k < r; k++) pi[(size_t)i * r + k] += v[(size_t)i * c + j] * a[(size_t)k * c + j] * w;
        }
        // back to the caller's orientation: pinv(A^T) = pinv(A)^T
        out.assign((size_t)c_in * r_in, 0.0L);
        for (int i = 0; i < c_in; i++)
            for (int k = 0; k < r_in; k++) out[(size_t)i * r_in + k] = tr ? pi[(size_t)k * r + i] : pi[(size_t)i * r + k];
        return smax > 0.0L;
    };
    std::vector<long double> e((size_t)nv * np, 0.0L); // [m][p]
    if (np == nv) {
        std::vector<long double> inv = pp; // np x nv square
        if (!invert_dense(inv, nv)) return GL_ERR_ZERO_DET;
        e = inv; // inv(P): [m][p]
    } else if (np > nv || np == 1) {
        if (!pinv(pp, np, nv, e)) return GL_ERR_ZERO_DET;
    } else {
        std::vector<long double> ppi, hx((size_t)np * nh), bb;
        if (!pinv(pp, np, nv, ppi)) return GL_ERR_ZERO_DET; // nv x np
        for (int p = 0; p < np; p++) {
            for (int k = 0; k < gd; k++) hx[(size_t)p * nh + k] = g[4 * p + k];
            hx[(size_t)p * nh + gd] = 1.0L;
        }
        if (!pinv(hx, np, nh, bb)) return GL_ERR_ZERO_DET; // nh x np
        std::vector<long double> aa((size_t)np * np);
        for (int i = 0; i < np; i++)
            for (int j = 0; j < np; j++) {
                long double sum = 0.0L;
                for (int k = 0; k < nh; k++) sum += hx[(size_t)i * nh + k] * bb[(size_t)k * np + j];
                aa[(size_t)i * np + j] = (i == j ? 1.0L : 0.0L) - sum;
            }
        for (int i = 0; i < nv; i++) {
            const double *rc = REF_COORDS[REF_COORDS_OFFSET[kind] + i];
            for (int j = 0; j < np; j++) {
                long double v = bb[(size_t)gd * np + j];
                for (int k = 0; k < gd; k++) v += (long double)rc[k] * bb[(size_t)k * np + j];
                for (int k = 0; k < np; k++) v += ppi[(size_t)i * np + k] * aa[(size_t)k * np + j];
                e[(size_t)i * np + j] = v;
            }
        }
    }
    for (int m = 0; m < nv; m++)
        for (int p = 0; p < np; p++) ee[m + (size_t)p * nv] = (double)e[(size_t)m * np + p];
    return GL_OK;
}

#define RULE(tab) do { *data = &tab[0][0]; *npoint = (int)(sizeof(tab) / sizeof(tab[0])); return GL_OK; } while (0)

// Gauss::new (src/integ/gauss.rs:87-121) when ngauss == 0, else Gauss::new_sized (gauss.rs:189-246)
int gauss_rule(int kind, int ngauss, int *npoint, const double **data) {
    if (kind < 0 || kind >= GL_NKIND) return GL_ERR_KIND_UNSUPPORTED;
    if (ngauss == 0) {
        static const int DEFAULT_N[GL_NKIND] = {2, 3, 4, 5, 3, 7, 12, 16, 4, 9, 9, 16, 16, 16, 4, 14, 24, 8, 27, 64};
        ngauss = DEFAULT_N[kind];
    }
    switch (KCLASS[kind]) {
    case 0:
        switch (ngauss) {
        case 1: RULE(GL_IP_LIN_LEGENDRE_1);
        case 2: RULE(GL_IP_LIN_LEGENDRE_2);
        case 3: RULE(GL_IP_LIN_LEGENDRE_3);
        case 4: RULE(GL_IP_LIN_LEGENDRE_4);
        case 5: RULE(GL_IP_LIN_LEGENDRE_5);
        default: return GL_ERR_GAUSS_LIN;
        }
    case 1:
        switch (ngauss) {
        case 1: RULE(GL_IP_TRI_INTERNAL_1);
        case 3: RULE(GL_IP_TRI_INTERNAL_3);
        case 4: RULE(GL_IP_TRI_INTERNAL_4);
        case 6: RULE(GL_IP_TRI_FELIPPA_6);
        case 7: RULE(GL_IP_TRI_FELIPPA_7);
        case 12: RULE(GL_IP_TRI_INTERNAL_12);
        case 16: RULE(GL_IP_TRI_INTERNAL_16);
        default: return GL_ERR_GAUSS_TRI;
        }
    case 2:
        switch (ngauss) {
        case 1: RULE(GL_IP_QUA_LEGENDRE_1);
        case 4: RULE(GL_IP_QUA_LEGENDRE_4);
        case 9: RULE(GL_IP_QUA_LEGENDRE_9);
        case 16: RULE(GL_IP_QUA_LEGENDRE_16);
        default: return GL_ERR_GAUSS_QUA;
        }
    case 3:
        switch (ngauss) {
        case 1: RULE(GL_IP_TET_INTERNAL_1);
        case 4: RULE(GL_IP_TET_INTERNAL_4);
        case 5: RULE(GL_IP_TET_INTERNAL_5);
        case 8: RULE(GL_IP_TET_FELIPPA_8);
        case 14: RULE(GL_IP_TET_FELIPPA_14);
        case 15: RULE(GL_IP_TET_FELIPPA_15);
        case 24: RULE(GL_IP_TET_FELIPPA_24);
        default: return GL_ERR_GAUSS_TET;
        }
    default:
        switch (ngauss) {
        case 6: RULE(GL_IP_HEX_IRONS_6);
        case 8: RULE(GL_IP_HEX_LEGENDRE_8);
        case 14: RULE(GL_IP_HEX_IRONS_14);
        case 27: RULE(GL_IP_HEX_LEGENDRE_27);
        case 64: RULE(GL_IP_HEX_LEGENDRE_64);
        default: return GL_ERR_GAUSS_HEX;
        }
    }
}

const char *status_string(int status) {
    switch (status) {
    case GL_OK: return "";
    case GL_ERR_XXT_NOT_SET: return "all components of the coordinates matrix must be set first";
    case GL_ERR_ZERO_DET: return "cannot compute inverse due to zero determinant";
    case GL_ERR_GRADIENT_NDIM: return "calc_gradient requires that geo_ndim = space_ndim";
    case GL_ERR_NROW_KK_NDOF: return "nrow(K) must be \xe2\x89\xa5 ii0 + nnode \xe2\x8b\x85 space_ndim";
    case GL_ERR_NCOL_KK_NDOF: return "ncol(K) must be \xe2\x89\xa5 jj0 + nnode \xe2\x8b\x85 space_ndim";
    case GL_ERR_AXISYM_NDIM: return "axisymmetric requires space_ndim = 2";
    case GL_ERR_NROW_KK_NNODE: return "nrow(K) must be \xe2\x89\xa5 ii0 + nnode";
    case GL_ERR_NCOL_KK_NNODE: return "ncol(K) must be \xe2\x89\xa5 jj0 + nnode";
    case GL_ERR_A_LEN: return "a.len() must be \xe2\x89\xa5 ii0 + nnode";
    case GL_ERR_B_LEN: return "b.len() must be \xe2\x89\xa5 ii0 + nnode \xe2\x8b\x85 space_ndim";
    case GL_ERR_C_LEN: return "c.len() must be \xe2\x89\xa5 ii0 + nnode";
    case GL_ERR_D_LEN: return "d.len() must be \xe2\x89\xa5 ii0 + nnode \xe2\x8b\x85 space_ndim";
    case GL_ERR_GAUSS_LIN: return "requested number of integration points is not available for Lin class";
    case GL_ERR_GAUSS_TRI: return "requested number of integration points is not available for Tri class";
    case GL_ERR_GAUSS_QUA: return "requested number of integration points is not available for Qua class";
    case GL_ERR_GAUSS_TET: return "requested number of integration points is not available for Tet class";
    case GL_ERR_GAUSS_HEX: return "requested number of integration points is not available for Hex class";
    case GL_ERR_KIND_UNSUPPORTED: return "GeoKind is not supported by libgemlab_cuda yet";
    case GL_ERR_SPACE_NDIM: return "space_ndim must be 2 or 3";
    case GL_ERR_NROW_KK_NNODE_B: return "nrow(K) must be \xe2\x89\xa5 ii0 + pad_b.nnode";
    case GL_ERR_NCOL_KK_NDOF_PAD: return "ncol(K) must be \xe2\x89\xa5 jj0 + pad.nnode \xe2\x8b\x85 space_ndim";
    case GL_ERR_NROW_KK_NDOF_PAD: return "nrow(K) must be \xe2\x89\xa5 ii0 + pad.nnode \xe2\x8b\x85 space_ndim";
    case GL_ERR_NCOL_KK_NNODE_B: return "ncol(K) must be \xe2\x89\xa5 jj0 + pad_b.nnode";
    case GL_ERR_BRY_2D: return "in 2D, geometry ndim must be equal to 1 (a line)";
    case GL_ERR_BRY_3D: return "in 3D, geometry ndim must be equal to 2 (a surface)";
    case GL_ERR_NORMAL_2D: return "calc_normal_vector requires geo_ndim = 1 in 2D (CABLE in 2D)";
    case GL_ERR_NORMAL_3D: return "calc_normal_vector requires geo_ndim = 2 in 3D (SHELL in 3D)";
    case GL_ERR_INVALID_ARG: return "invalid argument";
    case GL_ERR_CELL_RANGE: return "cell range is out of bounds";
    case GL_ERR_OUT_CAPACITY: return "output buffer is too small for the requested cell range";
    case GL_ERR_COEF: return "invalid coefficient specification";
    case GL_ERR_MESH: return "invalid mesh data";
    case GL_ERR_CUDA: return "CUDA error";
    case GL_ERR_NO_DEVICE: return "no usable sm_100 CUDA device (libgemlab_cuda has no CPU fallback)";
    default: return "unknown error";
    }
}

// russell_tensor::LinElasticity (Mandel basis); see SURVEY.md Appendix A.4
void lin_elasticity(double young, double poisson, int two_dim, int plane_stress, double *dd) {
    const int ns = two_dim ? 4 : 6;
    for (int i = 0; i < ns * ns; i++) dd[i] = 0.0;
    auto D = [&](int a, int b) -> double & { return dd[a + b * ns]; };
    if (two_dim && plane_stress) {
        const double c = young / (1.0 - poisson * poisson);
        D(0, 0) = c;
        D(1, 1) = c;
        D(0, 1) = c * poisson;
        D(1, 0) = c * poisson;
        D(3, 3) = c * (1.0 - poisson);
    } else {
        const double c = young / ((1.0 + poisson) * (1.0 - 2.0 * poisson));
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) D(a, b) = (a == b) ? c * (1.0 - poisson) : c * poisson;
        for (int q = 3; q < ns; q++) D(q, q) = c * (1.0 - 2.0 * poisson);
    }
}

} // namespace gl

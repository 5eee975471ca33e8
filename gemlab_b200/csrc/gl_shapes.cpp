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
#include "gl_internal.hpp"

#include <cmath>
#include <cstring>

namespace gl {

#include "gauss_tables.inc"
#include "shape_tables.inc"

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

// kinds whose shape functions this library can also evaluate at an ARBITRARY point (gl_shape_eval, GL_OP_JACOBIAN); the
// cubic / quartic kinds exist as tables at the reference's Gauss points only
bool kind_has_formulas(int kind) {
    switch (kind) {
    case GL_LIN2: case GL_LIN3: case GL_TRI3: case GL_TRI6: case GL_QUA4: case GL_QUA8: case GL_QUA9:
    case GL_TET4: case GL_TET10: case GL_HEX8: case GL_HEX20:
        return true;
    default:
        return false;
    }
}

int shape_eval(int kind, const double *ksi, double *nn, double *ll) {
    if (!kind_has_formulas(kind)) return GL_ERR_KIND_UNSUPPORTED;
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

/*
 * gemlab_oracle.c -- CPU restatement of gemlab's per-cell integration path (see gemlab_oracle.h).
 *
 * TEST INFRASTRUCTURE ONLY: parity oracle + timed CPU baseline.  Never linked into the product.
 * Every function cites the reference file:line (relative to the gemlab v2.4.0 tree) it follows.
 * The arithmetic keeps the reference's operation order so that the only differences against the
 * Rust original are the summation order inside russell_lab's dgemm calls.
 */
#include "gemlab_oracle.h"

#include <math.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#include "gauss_tables.inc"

#define SQRT_2 1.41421356237309504880168872420969808 /* russell_lab::math::SQRT_2 */
#define ZERO_DETERMINANT 1e-15                          /* russell_lab mat_inverse threshold */

/* ---------------------------------------------------------------------------------------------- */
/* errors                                                                                         */
/* ---------------------------------------------------------------------------------------------- */

const char *go_strerror(int code) {
    switch (code) {
    case GO_OK: return "";
    case GO_ERR_XXT_NOT_SET: return "all components of the coordinates matrix must be set first";
    case GO_ERR_ZERO_DET: return "cannot compute inverse due to zero determinant";
    case GO_ERR_GRADIENT_NDIM: return "calc_gradient requires that geo_ndim = space_ndim";
    case GO_ERR_NROW_KK_NDOF: return "nrow(K) must be \xe2\x89\xa5 ii0 + nnode \xe2\x8b\x85 space_ndim";
    case GO_ERR_NCOL_KK_NDOF: return "ncol(K) must be \xe2\x89\xa5 jj0 + nnode \xe2\x8b\x85 space_ndim";
    case GO_ERR_AXISYM_NDIM: return "axisymmetric requires space_ndim = 2";
    case GO_ERR_NROW_KK_NNODE: return "nrow(K) must be \xe2\x89\xa5 ii0 + nnode";
    case GO_ERR_NCOL_KK_NNODE: return "ncol(K) must be \xe2\x89\xa5 jj0 + nnode";
    case GO_ERR_A_LEN: return "a.len() must be \xe2\x89\xa5 ii0 + nnode";
    case GO_ERR_B_LEN: return "b.len() must be \xe2\x89\xa5 ii0 + nnode \xe2\x8b\x85 space_ndim";
    case GO_ERR_C_LEN: return "c.len() must be \xe2\x89\xa5 ii0 + nnode";
    case GO_ERR_D_LEN: return "d.len() must be \xe2\x89\xa5 ii0 + nnode \xe2\x8b\x85 space_ndim";
    case GO_ERR_GAUSS_LIN: return "requested number of integration points is not available for Lin class";
    case GO_ERR_GAUSS_TRI: return "requested number of integration points is not available for Tri class";
    case GO_ERR_GAUSS_QUA: return "requested number of integration points is not available for Qua class";
    case GO_ERR_GAUSS_TET: return "requested number of integration points is not available for Tet class";
    case GO_ERR_GAUSS_HEX: return "requested number of integration points is not available for Hex class";
    case GO_ERR_KIND_UNSUPPORTED: return "GeoKind is not restated by the oracle";
    case GO_ERR_CALLBACK_STOP: return "stop";
    case GO_ERR_SPACE_NDIM: return "space_ndim must be 2 or 3";
    case GO_ERR_NROW_KK_NNODE_B: return "nrow(K) must be \xe2\x89\xa5 ii0 + pad_b.nnode";
    case GO_ERR_NCOL_KK_NDOF_PAD: return "ncol(K) must be \xe2\x89\xa5 jj0 + pad.nnode \xe2\x8b\x85 space_ndim";
    case GO_ERR_NROW_KK_NDOF_PAD: return "nrow(K) must be \xe2\x89\xa5 ii0 + pad.nnode \xe2\x8b\x85 space_ndim";
    case GO_ERR_NCOL_KK_NNODE_B: return "ncol(K) must be \xe2\x89\xa5 jj0 + pad_b.nnode";
    case GO_ERR_BRY_2D: return "in 2D, geometry ndim must be equal to 1 (a line)";
    case GO_ERR_BRY_3D: return "in 3D, geometry ndim must be equal to 2 (a surface)";
    case GO_ERR_NORMAL_2D: return "calc_normal_vector requires geo_ndim = 1 in 2D (CABLE in 2D)";
    case GO_ERR_NORMAL_3D: return "calc_normal_vector requires geo_ndim = 2 in 3D (SHELL in 3D)";
    default: return "unknown error";
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* GeoKind tables (src/shapes/enums.rs:280-360, 217-244)                                          */
/* ---------------------------------------------------------------------------------------------- */

static const int KIND_NNODE[GO_NKIND] = {2, 3, 4, 5, 3, 6, 10, 15, 4, 8, 9, 12, 16, 17, 4, 10, 20, 8, 20, 32};
static const int KIND_NDIM[GO_NKIND] = {1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 3, 3, 3, 3, 3, 3};
static const int KIND_CLASS[GO_NKIND] = {0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 2, 2, 3, 3, 3, 4, 4, 4};
static const char *KIND_NAME[GO_NKIND] = {"lin2", "lin3", "lin4", "lin5", "tri3", "tri6", "tri10", "tri15", "qua4", "qua8",
                                          "qua9", "qua12", "qua16", "qua17", "tet4", "tet10", "tet20", "hex8", "hex20", "hex32"};

int go_kind_nnode(int kind) { return (kind >= 0 && kind < GO_NKIND) ? KIND_NNODE[kind] : -1; }
int go_kind_geo_ndim(int kind) { return (kind >= 0 && kind < GO_NKIND) ? KIND_NDIM[kind] : -1; }
int go_kind_class(int kind) { return (kind >= 0 && kind < GO_NKIND) ? KIND_CLASS[kind] : -1; }
const char *go_kind_name(int kind) { return (kind >= 0 && kind < GO_NKIND) ? KIND_NAME[kind] : ""; }
int go_kind_from_name(const char *name) {
    for (int k = 0; k < GO_NKIND; k++)
        if (strcmp(name, KIND_NAME[k]) == 0) return k;
    return -1;
}
int go_kind_supported(int kind) {
    switch (kind) {
    case GO_LIN2: case GO_LIN3: case GO_TRI3: case GO_TRI6: case GO_QUA4: case GO_QUA8: case GO_QUA9:
    case GO_TET4: case GO_TET10: case GO_HEX8: case GO_HEX20:
        return 1;
    default:
        return 0;
    }
}

/* NODE_REFERENCE_COORDS of each kind (lin2.rs:30, lin3.rs, tri3.rs:74, tri6.rs, qua4.rs:67, qua8.rs:68,
 * qua9.rs, tet4.rs:150, tet10.rs:150-161, hex8.rs:122-131, hex20.rs:123-148) */
static const double REF_LIN3[3][1] = {{-1.0}, {1.0}, {0.0}};
static const double REF_TRI6[6][2] = {{0, 0}, {1, 0}, {0, 1}, {0.5, 0}, {0.5, 0.5}, {0, 0.5}};
static const double REF_QUA9[9][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {0, 0}};
static const double REF_TET10[10][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0.5, 0, 0},
                                        {0.5, 0.5, 0}, {0, 0.5, 0}, {0, 0, 0.5}, {0.5, 0, 0.5}, {0, 0.5, 0.5}};
static const double REF_HEX20[20][3] = {
    {-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
    {0, -1, -1},  {1, 0, -1},  {0, 1, -1}, {-1, 0, -1}, {0, -1, 1},  {1, 0, 1},  {0, 1, 1}, {-1, 0, 1},
    {-1, -1, 0},  {1, -1, 0},  {1, 1, 0},  {-1, 1, 0}};

int go_kind_reference_coords(int kind, int m, double *ksi) {
    if (!go_kind_supported(kind) || m < 0 || m >= KIND_NNODE[kind]) return GO_ERR_KIND_UNSUPPORTED;
    switch (kind) {
    case GO_LIN2: ksi[0] = REF_LIN3[m][0]; break;
    case GO_LIN3: ksi[0] = REF_LIN3[m][0]; break;
    case GO_TRI3: case GO_TRI6: ksi[0] = REF_TRI6[m][0]; ksi[1] = REF_TRI6[m][1]; break;
    case GO_QUA4: case GO_QUA8: case GO_QUA9: ksi[0] = REF_QUA9[m][0]; ksi[1] = REF_QUA9[m][1]; break;
    case GO_TET4: case GO_TET10: ksi[0] = REF_TET10[m][0]; ksi[1] = REF_TET10[m][1]; ksi[2] = REF_TET10[m][2]; break;
    case GO_HEX8: case GO_HEX20: ksi[0] = REF_HEX20[m][0]; ksi[1] = REF_HEX20[m][1]; ksi[2] = REF_HEX20[m][2]; break;
    }
    return GO_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* shape functions N and derivatives L = dN/dksi                                                  */
/* ---------------------------------------------------------------------------------------------- */

#define L1(m, v) ll[(m)] = (v)
#define L2(m, j, v) ll[(m)*2 + (j)] = (v)
#define L3(m, j, v) ll[(m)*3 + (j)] = (v)

/* lin2.rs:43-62 */
static void lin2_interp(double *nn, const double *ksi) {
    double r = ksi[0];
    nn[0] = 0.5 * (1.0 - r);
    nn[1] = 0.5 * (1.0 + r);
}
static void lin2_deriv(double *ll, const double *ksi) {
    (void)ksi;
    L1(0, -0.5);
    L1(1, 0.5);
}

/* lin3.rs:49-73 */
static void lin3_interp(double *nn, const double *ksi) {
    double r = ksi[0];
    nn[0] = 0.5 * (r * r - r);
    nn[1] = 0.5 * (r * r + r);
    nn[2] = 1.0 - r * r;
}
static void lin3_deriv(double *ll, const double *ksi) {
    double r = ksi[0];
    L1(0, r - 0.5);
    L1(1, r + 0.5);
    L1(2, -2.0 * r);
}

/* tri3.rs:94-120 */
static void tri3_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    nn[0] = 1.0 - r - s;
    nn[1] = r;
    nn[2] = s;
}
static void tri3_deriv(double *ll, const double *ksi) {
    (void)ksi;
    L2(0, 0, -1.0); L2(1, 0, 1.0); L2(2, 0, 0.0);
    L2(0, 1, -1.0); L2(1, 1, 0.0); L2(2, 1, 1.0);
}

/* tri6.rs:101-138 */
static void tri6_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    nn[0] = 1.0 - (r + s) * (3.0 - 2.0 * (r + s));
    nn[1] = r * (2.0 * r - 1.0);
    nn[2] = s * (2.0 * s - 1.0);
    nn[3] = 4.0 * r * (1.0 - (r + s));
    nn[4] = 4.0 * r * s;
    nn[5] = 4.0 * s * (1.0 - (r + s));
}
static void tri6_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    L2(0, 0, -3.0 + 4.0 * (r + s));
    L2(1, 0, 4.0 * r - 1.0);
    L2(2, 0, 0.0);
    L2(3, 0, 4.0 - 8.0 * r - 4.0 * s);
    L2(4, 0, 4.0 * s);
    L2(5, 0, -4.0 * s);
    L2(0, 1, -3.0 + 4.0 * (r + s));
    L2(1, 1, 0.0);
    L2(2, 1, 4.0 * s - 1.0);
    L2(3, 1, -4.0 * r);
    L2(4, 1, 4.0 * r);
    L2(5, 1, 4.0 - 4.0 * r - 8.0 * s);
}

/* qua4.rs:89-122 */
static void qua4_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    nn[0] = (1.0 - r - s + r * s) / 4.0;
    nn[1] = (1.0 + r - s - r * s) / 4.0;
    nn[2] = (1.0 + r + s + r * s) / 4.0;
    nn[3] = (1.0 - r + s - r * s) / 4.0;
}
static void qua4_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    L2(0, 0, (-1.0 + s) / 4.0); L2(0, 1, (-1.0 + r) / 4.0);
    L2(1, 0, (1.0 - s) / 4.0);  L2(1, 1, (-1.0 - r) / 4.0);
    L2(2, 0, (1.0 + s) / 4.0);  L2(2, 1, (1.0 + r) / 4.0);
    L2(3, 0, (-1.0 - s) / 4.0); L2(3, 1, (1.0 - r) / 4.0);
}

/* qua8.rs:106-149 */
static void qua8_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    nn[0] = (1.0 - r) * (1.0 - s) * (-r - s - 1.0) / 4.0;
    nn[1] = (1.0 + r) * (1.0 - s) * (r - s - 1.0) / 4.0;
    nn[2] = (1.0 + r) * (1.0 + s) * (r + s - 1.0) / 4.0;
    nn[3] = (1.0 - r) * (1.0 + s) * (-r + s - 1.0) / 4.0;
    nn[4] = (1.0 - s) * (1.0 - r * r) / 2.0;
    nn[5] = (1.0 + r) * (1.0 - s * s) / 2.0;
    nn[6] = (1.0 + s) * (1.0 - r * r) / 2.0;
    nn[7] = (1.0 - r) * (1.0 - s * s) / 2.0;
}
static void qua8_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    L2(0, 0, -(1.0 - s) * (-r - r - s) / 4.0);
    L2(1, 0, (1.0 - s) * (r + r - s) / 4.0);
    L2(2, 0, (1.0 + s) * (r + r + s) / 4.0);
    L2(3, 0, -(1.0 + s) * (-r - r + s) / 4.0);
    L2(4, 0, -(1.0 - s) * r);
    L2(5, 0, (1.0 - s * s) / 2.0);
    L2(6, 0, -(1.0 + s) * r);
    L2(7, 0, -(1.0 - s * s) / 2.0);
    L2(0, 1, -(1.0 - r) * (-s - s - r) / 4.0);
    L2(1, 1, -(1.0 + r) * (-s - s + r) / 4.0);
    L2(2, 1, (1.0 + r) * (s + s + r) / 4.0);
    L2(3, 1, (1.0 - r) * (s + s - r) / 4.0);
    L2(4, 1, -(1.0 - r * r) / 2.0);
    L2(5, 1, -(1.0 + r) * s);
    L2(6, 1, (1.0 - r * r) / 2.0);
    L2(7, 1, -(1.0 - r) * s);
}

/* qua9.rs:104-156 */
static void qua9_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    nn[0] = r * (r - 1.0) * s * (s - 1.0) / 4.0;
    nn[1] = r * (r + 1.0) * s * (s - 1.0) / 4.0;
    nn[2] = r * (r + 1.0) * s * (s + 1.0) / 4.0;
    nn[3] = r * (r - 1.0) * s * (s + 1.0) / 4.0;
    nn[4] = -(r * r - 1.0) * s * (s - 1.0) / 2.0;
    nn[5] = -r * (r + 1.0) * (s * s - 1.0) / 2.0;
    nn[6] = -(r * r - 1.0) * s * (s + 1.0) / 2.0;
    nn[7] = -r * (r - 1.0) * (s * s - 1.0) / 2.0;
    nn[8] = (r * r - 1.0) * (s * s - 1.0);
}
static void qua9_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1];
    L2(0, 0, (r + r - 1.0) * s * (s - 1.0) / 4.0);
    L2(1, 0, (r + r + 1.0) * s * (s - 1.0) / 4.0);
    L2(2, 0, (r + r + 1.0) * s * (s + 1.0) / 4.0);
    L2(3, 0, (r + r - 1.0) * s * (s + 1.0) / 4.0);
    L2(0, 1, r * (r - 1.0) * (s + s - 1.0) / 4.0);
    L2(1, 1, r * (r + 1.0) * (s + s - 1.0) / 4.0);
    L2(2, 1, r * (r + 1.0) * (s + s + 1.0) / 4.0);
    L2(3, 1, r * (r - 1.0) * (s + s + 1.0) / 4.0);
    L2(4, 0, -(r + r) * s * (s - 1.0) / 2.0);
    L2(5, 0, -(r + r + 1.0) * (s * s - 1.0) / 2.0);
    L2(6, 0, -(r + r) * s * (s + 1.0) / 2.0);
    L2(7, 0, -(r + r - 1.0) * (s * s - 1.0) / 2.0);
    L2(4, 1, -(r * r - 1.0) * (s + s - 1.0) / 2.0);
    L2(5, 1, -r * (r + 1.0) * (s + s) / 2.0);
    L2(6, 1, -(r * r - 1.0) * (s + s + 1.0) / 2.0);
    L2(7, 1, -r * (r - 1.0) * (s + s) / 2.0);
    L2(8, 0, 2.0 * r * (s * s - 1.0));
    L2(8, 1, 2.0 * s * (r * r - 1.0));
}

/* tet4.rs:166-200 */
static void tet4_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    nn[0] = 1.0 - r - s - t;
    nn[1] = r;
    nn[2] = s;
    nn[3] = t;
}
static void tet4_deriv(double *ll, const double *ksi) {
    (void)ksi;
    L3(0, 0, -1.0); L3(1, 0, 1.0); L3(2, 0, 0.0); L3(3, 0, 0.0);
    L3(0, 1, -1.0); L3(1, 1, 0.0); L3(2, 1, 1.0); L3(3, 1, 0.0);
    L3(0, 2, -1.0); L3(1, 2, 0.0); L3(2, 2, 0.0); L3(3, 2, 1.0);
}

/* tet10.rs:172-234 */
static void tet10_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    double u = 1.0 - r - s - t;
    nn[0] = u * (2.0 * u - 1.0);
    nn[1] = r * (2.0 * r - 1.0);
    nn[2] = s * (2.0 * s - 1.0);
    nn[3] = t * (2.0 * t - 1.0);
    nn[4] = 4.0 * u * r;
    nn[5] = 4.0 * r * s;
    nn[6] = 4.0 * s * u;
    nn[7] = 4.0 * u * t;
    nn[8] = 4.0 * r * t;
    nn[9] = 4.0 * s * t;
}
static void tet10_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    L3(0, 0, 4.0 * (r + s + t) - 3.0);
    L3(1, 0, 4.0 * r - 1.0);
    L3(2, 0, 0.0);
    L3(3, 0, 0.0);
    L3(4, 0, 4.0 - 8.0 * r - 4.0 * s - 4.0 * t);
    L3(5, 0, 4.0 * s);
    L3(6, 0, -4.0 * s);
    L3(7, 0, -4.0 * t);
    L3(8, 0, 4.0 * t);
    L3(9, 0, 0.0);
    L3(0, 1, 4.0 * (r + s + t) - 3.0);
    L3(1, 1, 0.0);
    L3(2, 1, 4.0 * s - 1.0);
    L3(3, 1, 0.0);
    L3(4, 1, -4.0 * r);
    L3(5, 1, 4.0 * r);
    L3(6, 1, 4.0 - 4.0 * r - 8.0 * s - 4.0 * t);
    L3(7, 1, -4.0 * t);
    L3(8, 1, 0.0);
    L3(9, 1, 4.0 * t);
    L3(0, 2, 4.0 * (r + s + t) - 3.0);
    L3(1, 2, 0.0);
    L3(2, 2, 0.0);
    L3(3, 2, 4.0 * t - 1.0);
    L3(4, 2, -4.0 * r);
    L3(5, 2, 0.0);
    L3(6, 2, -4.0 * s);
    L3(7, 2, 4.0 - 4.0 * r - 4.0 * s - 8.0 * t);
    L3(8, 2, 4.0 * r);
    L3(9, 2, 4.0 * s);
}

/* hex8.rs:142-211 */
static void hex8_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    double rm = 1.0 - r, sm = 1.0 - s, tm = 1.0 - t;
    double rp = 1.0 + r, sp = 1.0 + s, tp = 1.0 + t;
    nn[0] = rm * sm * tm / 8.0;
    nn[1] = rp * sm * tm / 8.0;
    nn[2] = rp * sp * tm / 8.0;
    nn[3] = rm * sp * tm / 8.0;
    nn[4] = rm * sm * tp / 8.0;
    nn[5] = rp * sm * tp / 8.0;
    nn[6] = rp * sp * tp / 8.0;
    nn[7] = rm * sp * tp / 8.0;
}
static void hex8_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    double rm = 1.0 - r, sm = 1.0 - s, tm = 1.0 - t;
    double rp = 1.0 + r, sp = 1.0 + s, tp = 1.0 + t;
    L3(0, 0, -sm * tm / 8.0); L3(1, 0, sm * tm / 8.0);  L3(2, 0, sp * tm / 8.0); L3(3, 0, -sp * tm / 8.0);
    L3(4, 0, -sm * tp / 8.0); L3(5, 0, sm * tp / 8.0);  L3(6, 0, sp * tp / 8.0); L3(7, 0, -sp * tp / 8.0);
    L3(0, 1, -rm * tm / 8.0); L3(1, 1, -rp * tm / 8.0); L3(2, 1, rp * tm / 8.0); L3(3, 1, rm * tm / 8.0);
    L3(4, 1, -rm * tp / 8.0); L3(5, 1, -rp * tp / 8.0); L3(6, 1, rp * tp / 8.0); L3(7, 1, rm * tp / 8.0);
    L3(0, 2, -rm * sm / 8.0); L3(1, 2, -rp * sm / 8.0); L3(2, 2, -rp * sp / 8.0); L3(3, 2, -rm * sp / 8.0);
    L3(4, 2, rm * sm / 8.0);  L3(5, 2, rp * sm / 8.0);  L3(6, 2, rp * sp / 8.0);  L3(7, 2, rm * sp / 8.0);
}

/* hex20.rs:159-282 */
static void hex20_interp(double *nn, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    double rm = 1.0 - r, sm = 1.0 - s, tm = 1.0 - t;
    double rp = 1.0 + r, sp = 1.0 + s, tp = 1.0 + t;
    double rr = 1.0 - r * r, ss = 1.0 - s * s, tt = 1.0 - t * t;
    nn[0] = rm * sm * tm * (-2.0 - r - s - t) / 8.0;
    nn[1] = rp * sm * tm * (-2.0 + r - s - t) / 8.0;
    nn[2] = rp * sp * tm * (-2.0 + r + s - t) / 8.0;
    nn[3] = rm * sp * tm * (-2.0 - r + s - t) / 8.0;
    nn[4] = rm * sm * tp * (-2.0 - r - s + t) / 8.0;
    nn[5] = rp * sm * tp * (-2.0 + r - s + t) / 8.0;
    nn[6] = rp * sp * tp * (-2.0 + r + s + t) / 8.0;
    nn[7] = rm * sp * tp * (-2.0 - r + s + t) / 8.0;
    nn[8] = rr * sm * tm / 4.0;
    nn[9] = rp * ss * tm / 4.0;
    nn[10] = rr * sp * tm / 4.0;
    nn[11] = rm * ss * tm / 4.0;
    nn[12] = rr * sm * tp / 4.0;
    nn[13] = rp * ss * tp / 4.0;
    nn[14] = rr * sp * tp / 4.0;
    nn[15] = rm * ss * tp / 4.0;
    nn[16] = rm * sm * tt / 4.0;
    nn[17] = rp * sm * tt / 4.0;
    nn[18] = rp * sp * tt / 4.0;
    nn[19] = rm * sp * tt / 4.0;
}
static void hex20_deriv(double *ll, const double *ksi) {
    double r = ksi[0], s = ksi[1], t = ksi[2];
    double rm = 1.0 - r, sm = 1.0 - s, tm = 1.0 - t;
    double rp = 1.0 + r, sp = 1.0 + s, tp = 1.0 + t;
    double rr = 1.0 - r * r, ss = 1.0 - s * s, tt = 1.0 - t * t;
    /* with respect to r */
    L3(0, 0, -rm * sm * tm / 8.0 - sm * (-2.0 - r - s - t) * tm / 8.0);
    L3(1, 0, rp * sm * tm / 8.0 + sm * (-2.0 + r - s - t) * tm / 8.0);
    L3(2, 0, rp * sp * tm / 8.0 + sp * (-2.0 + r + s - t) * tm / 8.0);
    L3(3, 0, -rm * sp * tm / 8.0 - sp * (-2.0 - r + s - t) * tm / 8.0);
    L3(4, 0, -rm * sm * tp / 8.0 - sm * (-2.0 - r - s + t) * tp / 8.0);
    L3(5, 0, rp * sm * tp / 8.0 + sm * (-2.0 + r - s + t) * tp / 8.0);
    L3(6, 0, rp * sp * tp / 8.0 + sp * (-2.0 + r + s + t) * tp / 8.0);
    L3(7, 0, -rm * sp * tp / 8.0 - sp * (-2.0 - r + s + t) * tp / 8.0);
    L3(8, 0, -r * sm * tm / 2.0);
    L3(9, 0, ss * tm / 4.0);
    L3(10, 0, -r * sp * tm / 2.0);
    L3(11, 0, -ss * tm / 4.0);
    L3(12, 0, -r * sm * tp / 2.0);
    L3(13, 0, ss * tp / 4.0);
    L3(14, 0, -r * sp * tp / 2.0);
    L3(15, 0, -ss * tp / 4.0);
    L3(16, 0, -sm * tt / 4.0);
    L3(17, 0, sm * tt / 4.0);
    L3(18, 0, sp * tt / 4.0);
    L3(19, 0, -sp * tt / 4.0);
    /* with respect to s */
    L3(0, 1, -rm * sm * tm / 8.0 - rm * (-2.0 - r - s - t) * tm / 8.0);
    L3(1, 1, -rp * sm * tm / 8.0 - rp * (-2.0 + r - s - t) * tm / 8.0);
    L3(2, 1, rp * sp * tm / 8.0 + rp * (-2.0 + r + s - t) * tm / 8.0);
    L3(3, 1, rm * sp * tm / 8.0 + rm * (-2.0 - r + s - t) * tm / 8.0);
    L3(4, 1, -rm * sm * tp / 8.0 - rm * (-2.0 - r - s + t) * tp / 8.0);
    L3(5, 1, -rp * sm * tp / 8.0 - rp * (-2.0 + r - s + t) * tp / 8.0);
    L3(6, 1, rp * sp * tp / 8.0 + rp * (-2.0 + r + s + t) * tp / 8.0);
    L3(7, 1, rm * sp * tp / 8.0 + rm * (-2.0 - r + s + t) * tp / 8.0);
    L3(8, 1, -rr * tm / 4.0);
    L3(9, 1, -rp * s * tm / 2.0);
    L3(10, 1, rr * tm / 4.0);
    L3(11, 1, -rm * s * tm / 2.0);
    L3(12, 1, -rr * tp / 4.0);
    L3(13, 1, -rp * s * tp / 2.0);
    L3(14, 1, rr * tp / 4.0);
    L3(15, 1, -rm * s * tp / 2.0);
    L3(16, 1, -rm * tt / 4.0);
    L3(17, 1, -rp * tt / 4.0);
    L3(18, 1, rp * tt / 4.0);
    L3(19, 1, rm * tt / 4.0);
    /* with respect to t */
    L3(0, 2, -rm * sm * (-2.0 - r - s - t) / 8.0 - rm * sm * tm / 8.0);
    L3(1, 2, -rp * sm * (-2.0 + r - s - t) / 8.0 - rp * sm * tm / 8.0);
    L3(2, 2, -rp * sp * (-2.0 + r + s - t) / 8.0 - rp * sp * tm / 8.0);
    L3(3, 2, -rm * sp * (-2.0 - r + s - t) / 8.0 - rm * sp * tm / 8.0);
    L3(4, 2, rm * sm * (-2.0 - r - s + t) / 8.0 + rm * sm * tp / 8.0);
    L3(5, 2, rp * sm * (-2.0 + r - s + t) / 8.0 + rp * sm * tp / 8.0);
    L3(6, 2, rp * sp * (-2.0 + r + s + t) / 8.0 + rp * sp * tp / 8.0);
    L3(7, 2, rm * sp * (-2.0 - r + s + t) / 8.0 + rm * sp * tp / 8.0);
    L3(8, 2, -rr * sm / 4.0);
    L3(9, 2, -rp * ss / 4.0);
    L3(10, 2, -rr * sp / 4.0);
    L3(11, 2, -rm * ss / 4.0);
    L3(12, 2, rr * sm / 4.0);
    L3(13, 2, rp * ss / 4.0);
    L3(14, 2, rr * sp / 4.0);
    L3(15, 2, rm * ss / 4.0);
    L3(16, 2, -rm * sm * t / 2.0);
    L3(17, 2, -rp * sm * t / 2.0);
    L3(18, 2, -rp * sp * t / 2.0);
    L3(19, 2, -rm * sp * t / 2.0);
}

/* Lin4/5, Tri10/15, Qua12/16/17, Tet20, Hex32 are not restated node by node: their N and L at the reference's Gauss points
 * come from shape_tables_extra.inc, evaluated from the reference's own calc_interp / calc_deriv expressions by
 * tools/gen_shape_tables.py.  The integration loops only ever evaluate shapes at Gauss points, so a lookup by the exact
 * point serves them; any other ksi is GO_ERR_KIND_UNSUPPORTED for these kinds. */
#include "shape_tables_extra.inc"

int go_gauss(int kind, int ngauss, const double (**data)[4], int *npoint);

static const double *extra_shape_row(int kind, const double *ksi, int want_deriv) {
    const int nn = KIND_NNODE[kind], gd = KIND_NDIM[kind];
    for (int t = 0; t < XN_SHAPE_TABLES; t++) {
        if (XSHAPE_TABLES[t].kind != kind) continue;
        const double (*pts)[4];
        int np = 0;
        if (go_gauss(kind, XSHAPE_TABLES[t].ng, &pts, &np) != GO_OK || np != XSHAPE_TABLES[t].ng) continue;
        for (int p = 0; p < np; p++) {
            int same = 1;
            for (int j = 0; j < gd; j++) same = same && (pts[p][j] == ksi[j]);
            if (!same) continue;
            const double *base = XSHAPE_TABLES[t].data;
            return want_deriv ? base + (size_t)np * nn + (size_t)p * nn * gd : base + (size_t)p * nn;
        }
    }
    return NULL;
}

/* GeoKind::functions (src/shapes/enums.rs:1124-1152) */
int go_interp(int kind, const double *ksi, double *nn) {
    switch (kind) {
    case GO_LIN2: lin2_interp(nn, ksi); break;
    case GO_LIN3: lin3_interp(nn, ksi); break;
    case GO_TRI3: tri3_interp(nn, ksi); break;
    case GO_TRI6: tri6_interp(nn, ksi); break;
    case GO_QUA4: qua4_interp(nn, ksi); break;
    case GO_QUA8: qua8_interp(nn, ksi); break;
    case GO_QUA9: qua9_interp(nn, ksi); break;
    case GO_TET4: tet4_interp(nn, ksi); break;
    case GO_TET10: tet10_interp(nn, ksi); break;
    case GO_HEX8: hex8_interp(nn, ksi); break;
    case GO_HEX20: hex20_interp(nn, ksi); break;
    default: {
        if (kind < 0 || kind >= GO_NKIND) return GO_ERR_KIND_UNSUPPORTED;
        const double *row = extra_shape_row(kind, ksi, 0);
        if (!row) return GO_ERR_KIND_UNSUPPORTED;
        memcpy(nn, row, sizeof(double) * (size_t)KIND_NNODE[kind]);
    }
    }
    return GO_OK;
}
int go_deriv(int kind, const double *ksi, double *ll) {
    switch (kind) {
    case GO_LIN2: lin2_deriv(ll, ksi); break;
    case GO_LIN3: lin3_deriv(ll, ksi); break;
    case GO_TRI3: tri3_deriv(ll, ksi); break;
    case GO_TRI6: tri6_deriv(ll, ksi); break;
    case GO_QUA4: qua4_deriv(ll, ksi); break;
    case GO_QUA8: qua8_deriv(ll, ksi); break;
    case GO_QUA9: qua9_deriv(ll, ksi); break;
    case GO_TET4: tet4_deriv(ll, ksi); break;
    case GO_TET10: tet10_deriv(ll, ksi); break;
    case GO_HEX8: hex8_deriv(ll, ksi); break;
    case GO_HEX20: hex20_deriv(ll, ksi); break;
    default: {
        if (kind < 0 || kind >= GO_NKIND) return GO_ERR_KIND_UNSUPPORTED;
        const double *row = extra_shape_row(kind, ksi, 1);
        if (!row) return GO_ERR_KIND_UNSUPPORTED;
        memcpy(ll, row, sizeof(double) * (size_t)KIND_NNODE[kind] * (size_t)KIND_NDIM[kind]);
    }
    }
    return GO_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* Gauss (src/integ/gauss.rs:87-121 default rule, :189-246 sized rule)                            */
/* ---------------------------------------------------------------------------------------------- */

#define RULE(tab) do { *data = tab; *npoint = (int)(sizeof(tab) / sizeof(tab[0])); return GO_OK; } while (0)

int go_gauss(int kind, int ngauss, const double (**data)[4], int *npoint) {
    if (kind < 0 || kind >= GO_NKIND) return GO_ERR_KIND_UNSUPPORTED;
    if (ngauss == 0) {
        switch (kind) {
        case GO_LIN2: RULE(GO_IP_LIN_LEGENDRE_2);
        case GO_LIN3: RULE(GO_IP_LIN_LEGENDRE_3);
        case GO_LIN4: RULE(GO_IP_LIN_LEGENDRE_4);
        case GO_LIN5: RULE(GO_IP_LIN_LEGENDRE_5);
        case GO_TRI3: RULE(GO_IP_TRI_INTERNAL_3);
        case GO_TRI6: RULE(GO_IP_TRI_FELIPPA_7);
        case GO_TRI10: RULE(GO_IP_TRI_INTERNAL_12);
        case GO_TRI15: RULE(GO_IP_TRI_INTERNAL_16);
        case GO_QUA4: RULE(GO_IP_QUA_LEGENDRE_4);
        case GO_QUA8: RULE(GO_IP_QUA_LEGENDRE_9);
        case GO_QUA9: RULE(GO_IP_QUA_LEGENDRE_9);
        case GO_QUA12: RULE(GO_IP_QUA_LEGENDRE_16);
        case GO_QUA16: RULE(GO_IP_QUA_LEGENDRE_16);
        case GO_QUA17: RULE(GO_IP_QUA_LEGENDRE_16);
        case GO_TET4: RULE(GO_IP_TET_INTERNAL_4);
        case GO_TET10: RULE(GO_IP_TET_FELIPPA_14);
        case GO_TET20: RULE(GO_IP_TET_FELIPPA_24);
        case GO_HEX8: RULE(GO_IP_HEX_LEGENDRE_8);
        case GO_HEX20: RULE(GO_IP_HEX_LEGENDRE_27);
        case GO_HEX32: RULE(GO_IP_HEX_LEGENDRE_64);
        }
    }
    switch (KIND_CLASS[kind]) {
    case GO_CLASS_LIN:
        switch (ngauss) {
        case 1: RULE(GO_IP_LIN_LEGENDRE_1);
        case 2: RULE(GO_IP_LIN_LEGENDRE_2);
        case 3: RULE(GO_IP_LIN_LEGENDRE_3);
        case 4: RULE(GO_IP_LIN_LEGENDRE_4);
        case 5: RULE(GO_IP_LIN_LEGENDRE_5);
        default: return GO_ERR_GAUSS_LIN;
        }
    case GO_CLASS_TRI:
        switch (ngauss) {
        case 1: RULE(GO_IP_TRI_INTERNAL_1);
        case 3: RULE(GO_IP_TRI_INTERNAL_3);
        case 4: RULE(GO_IP_TRI_INTERNAL_4);
        case 6: RULE(GO_IP_TRI_FELIPPA_6);
        case 7: RULE(GO_IP_TRI_FELIPPA_7);
        case 12: RULE(GO_IP_TRI_INTERNAL_12);
        case 16: RULE(GO_IP_TRI_INTERNAL_16);
        default: return GO_ERR_GAUSS_TRI;
        }
    case GO_CLASS_QUA:
        switch (ngauss) {
        case 1: RULE(GO_IP_QUA_LEGENDRE_1);
        case 4: RULE(GO_IP_QUA_LEGENDRE_4);
        case 9: RULE(GO_IP_QUA_LEGENDRE_9);
        case 16: RULE(GO_IP_QUA_LEGENDRE_16);
        default: return GO_ERR_GAUSS_QUA;
        }
    case GO_CLASS_TET:
        switch (ngauss) {
        case 1: RULE(GO_IP_TET_INTERNAL_1);
        case 4: RULE(GO_IP_TET_INTERNAL_4);
        case 5: RULE(GO_IP_TET_INTERNAL_5);
        case 8: RULE(GO_IP_TET_FELIPPA_8);
        case 14: RULE(GO_IP_TET_FELIPPA_14);
        case 15: RULE(GO_IP_TET_FELIPPA_15);
        case 24: RULE(GO_IP_TET_FELIPPA_24);
        default: return GO_ERR_GAUSS_TET;
        }
    default:
        switch (ngauss) {
        case 6: RULE(GO_IP_HEX_IRONS_6);
        case 8: RULE(GO_IP_HEX_LEGENDRE_8);
        case 14: RULE(GO_IP_HEX_IRONS_14);
        case 27: RULE(GO_IP_HEX_LEGENDRE_27);
        case 64: RULE(GO_IP_HEX_LEGENDRE_64);
        default: return GO_ERR_GAUSS_HEX;
        }
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* Scratchpad                                                                                     */
/* ---------------------------------------------------------------------------------------------- */

/* Scratchpad::new (src/shapes/scratchpad.rs:152-182) */
int go_pad_init(go_pad *pad, int space_ndim, int kind) {
    if (space_ndim < 2 || space_ndim > 3) return GO_ERR_SPACE_NDIM;
    if (kind < 0 || kind >= GO_NKIND) return GO_ERR_KIND_UNSUPPORTED; /* table-only kinds work at Gauss points (go_interp) */
    if (KIND_NDIM[kind] > space_ndim) return GO_ERR_SPACE_NDIM;
    memset(pad, 0, sizeof(*pad));
    pad->kind = kind;
    pad->space_ndim = space_ndim;
    pad->geo_ndim = KIND_NDIM[kind];
    pad->nnode = KIND_NNODE[kind];
    pad->ok_xxt = 0;
    return GO_OK;
}

/* Scratchpad::set_xx (src/shapes/scratchpad.rs:262-269): ok flag flips when the LAST entry is set */
void go_pad_set_xx(go_pad *pad, int m, int j, double value) {
    pad->ok_xxt = 0;
    pad->xxt[j * pad->nnode + m] = value;
    if (m == pad->nnode - 1 && j == pad->space_ndim - 1) pad->ok_xxt = 1;
}

void go_pad_calc_interp(go_pad *pad, const double *ksi) { go_interp(pad->kind, ksi, pad->interp); }

/* russell_lab::mat_inverse for 1x1, 2x2, 3x3 (closed form; returns det; fails when |det| <= 1e-15).
 * a, ai: n x n row-major. */
static int mat_inverse_small(double *ai, const double *a, int n, double *det_out) {
    if (n == 1) {
        double det = a[0];
        if (fabs(det) <= ZERO_DETERMINANT) return GO_ERR_ZERO_DET;
        ai[0] = 1.0 / det;
        *det_out = det;
        return GO_OK;
    }
    if (n == 2) {
        double det = a[0] * a[3] - a[1] * a[2];
        if (fabs(det) <= ZERO_DETERMINANT) return GO_ERR_ZERO_DET;
        ai[0] = a[3] / det;
        ai[1] = -a[1] / det;
        ai[2] = -a[2] / det;
        ai[3] = a[0] / det;
        *det_out = det;
        return GO_OK;
    }
    {
        double a00 = a[0], a01 = a[1], a02 = a[2];
        double a10 = a[3], a11 = a[4], a12 = a[5];
        double a20 = a[6], a21 = a[7], a22 = a[8];
        double det = a00 * (a11 * a22 - a12 * a21) - a01 * (a10 * a22 - a12 * a20) + a02 * (a10 * a21 - a11 * a20);
        if (fabs(det) <= ZERO_DETERMINANT) return GO_ERR_ZERO_DET;
        ai[0] = (a11 * a22 - a12 * a21) / det;
        ai[1] = (a02 * a21 - a01 * a22) / det;
        ai[2] = (a01 * a12 - a02 * a11) / det;
        ai[3] = (a12 * a20 - a10 * a22) / det;
        ai[4] = (a00 * a22 - a02 * a20) / det;
        ai[5] = (a02 * a10 - a00 * a12) / det;
        ai[6] = (a10 * a21 - a11 * a20) / det;
        ai[7] = (a01 * a20 - a00 * a21) / det;
        ai[8] = (a00 * a11 - a01 * a10) / det;
        *det_out = det;
        return GO_OK;
    }
}

/* Scratchpad::calc_jacobian (src/shapes/scratchpad_calc_jacobian.rs:101-131) */
int go_pad_calc_jacobian(go_pad *pad, const double *ksi, double *det_jac) {
    if (!pad->ok_xxt) return GO_ERR_XXT_NOT_SET;
    const int sd = pad->space_ndim, gd = pad->geo_ndim, nn = pad->nnode;
    /* matrix L: dN/dksi */
    go_deriv(pad->kind, ksi, pad->deriv);
    /* matrix J = Xt . L (mat_mat_mul -> dgemm in the reference) */
    for (int i = 0; i < sd; i++) {
        for (int j = 0; j < gd; j++) {
            double sum = 0.0;
            for (int m = 0; m < nn; m++) sum += pad->xxt[i * nn + m] * pad->deriv[m * gd + j];
            pad->jacobian[i * gd + j] = sum;
        }
    }
    if (gd == sd) {
        /* SOLID: inverse J (returns determinant) */
        return mat_inverse_small(pad->inv_jacobian, pad->jacobian, sd, det_jac);
    }
    if (gd == 1) {
        /* CABLE: norm of Jacobian vector */
        double norm_jac = 0.0;
        for (int i = 0; i < sd; i++) norm_jac += pad->jacobian[i] * pad->jacobian[i];
        *det_jac = sqrt(norm_jac);
        return GO_OK;
    }
    /* SHELL */
    *det_jac = -1.0; /* DET_JAC_NOT_AVAILABLE */
    return GO_OK;
}

/* Scratchpad::calc_gradient (src/shapes/scratchpad_calc_gradient.rs:78-91) */
int go_pad_calc_gradient(go_pad *pad, const double *ksi, double *det_jac) {
    const int sd = pad->space_ndim, gd = pad->geo_ndim, nn = pad->nnode;
    if (gd != sd) return GO_ERR_GRADIENT_NDIM;
    int err = go_pad_calc_jacobian(pad, ksi, det_jac);
    if (err) return err;
    /* B = L . inv(J) */
    for (int m = 0; m < nn; m++) {
        for (int j = 0; j < sd; j++) {
            double sum = 0.0;
            for (int k = 0; k < sd; k++) sum += pad->deriv[m * gd + k] * pad->inv_jacobian[k * sd + j];
            pad->gradient[m * sd + j] = sum;
        }
    }
    return GO_OK;
}

/* Scratchpad::calc_coords (src/shapes/scratchpad_calc_coords.rs:62-73): x = Xt . N */
void go_pad_calc_coords(go_pad *pad, const double *ksi, double *x) {
    go_pad_calc_interp(pad, ksi);
    for (int i = 0; i < pad->space_ndim; i++) {
        double sum = 0.0;
        for (int m = 0; m < pad->nnode; m++) sum += pad->xxt[i * pad->nnode + m] * pad->interp[m];
        x[i] = sum;
    }
}

/* ---------------------------------------------------------------------------------------------- */
/* integ                                                                                          */
/* ---------------------------------------------------------------------------------------------- */

#define B(m, j) (pad->gradient[(m)*sd + (j)])
#define D(a, b) (dd[(a) + (b)*ns])
#define KK(i, j) (kk[(size_t)(i) + (size_t)(j) * (size_t)nrow])
#define XXT(j, m) (pad->xxt[(j)*nnode + (m)])

/* add_to_kk (src/integ/mat_10_bdb.rs:179-208) */
static void add_to_kk(double *kk, int nrow, int sd, int nnode, double c, const double *dd, const go_args *args) {
    const double s = SQRT_2;
    const go_pad *pad = args->pad;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (sd == 2) {
        const int ns = 4;
        for (int m = 0; m < nnode; m++) {
            for (int n = 0; n < nnode; n++) {
                KK(ii0+0+m*2,jj0+0+n*2) += c * (B(m,1)*B(n,1)*D(3,3) + s*B(m,1)*B(n,0)*D(3,0) + s*B(m,0)*B(n,1)*D(0,3) + 2.0*B(m,0)*B(n,0)*D(0,0)) / 2.0;
                KK(ii0+0+m*2,jj0+1+n*2) += c * (B(m,1)*B(n,0)*D(3,3) + s*B(m,1)*B(n,1)*D(3,1) + s*B(m,0)*B(n,0)*D(0,3) + 2.0*B(m,0)*B(n,1)*D(0,1)) / 2.0;
                KK(ii0+1+m*2,jj0+0+n*2) += c * (B(m,0)*B(n,1)*D(3,3) + s*B(m,0)*B(n,0)*D(3,0) + s*B(m,1)*B(n,1)*D(1,3) + 2.0*B(m,1)*B(n,0)*D(1,0)) / 2.0;
                KK(ii0+1+m*2,jj0+1+n*2) += c * (B(m,0)*B(n,0)*D(3,3) + s*B(m,0)*B(n,1)*D(3,1) + s*B(m,1)*B(n,0)*D(1,3) + 2.0*B(m,1)*B(n,1)*D(1,1)) / 2.0;
            }
        }
    } else {
        const int ns = 6;
        for (int m = 0; m < nnode; m++) {
            for (int n = 0; n < nnode; n++) {
                KK(ii0+0+m*3,jj0+0+n*3) += c * (B(m,2)*B(n,2)*D(5,5) + B(m,2)*B(n,1)*D(5,3) + s*B(m,2)*B(n,0)*D(5,0) + B(m,1)*B(n,2)*D(3,5) + B(m,1)*B(n,1)*D(3,3) + s*B(m,1)*B(n,0)*D(3,0) + s*B(m,0)*B(n,2)*D(0,5) + s*B(m,0)*B(n,1)*D(0,3) + 2.0*B(m,0)*B(n,0)*D(0,0)) / 2.0;
                KK(ii0+0+m*3,jj0+1+n*3) += c * (B(m,2)*B(n,2)*D(5,4) + B(m,2)*B(n,0)*D(5,3) + s*B(m,2)*B(n,1)*D(5,1) + B(m,1)*B(n,2)*D(3,4) + B(m,1)*B(n,0)*D(3,3) + s*B(m,1)*B(n,1)*D(3,1) + s*B(m,0)*B(n,2)*D(0,4) + s*B(m,0)*B(n,0)*D(0,3) + 2.0*B(m,0)*B(n,1)*D(0,1)) / 2.0;
                KK(ii0+0+m*3,jj0+2+n*3) += c * (B(m,2)*B(n,0)*D(5,5) + B(m,2)*B(n,1)*D(5,4) + s*B(m,2)*B(n,2)*D(5,2) + B(m,1)*B(n,0)*D(3,5) + B(m,1)*B(n,1)*D(3,4) + s*B(m,1)*B(n,2)*D(3,2) + s*B(m,0)*B(n,0)*D(0,5) + s*B(m,0)*B(n,1)*D(0,4) + 2.0*B(m,0)*B(n,2)*D(0,2)) / 2.0;
                KK(ii0+1+m*3,jj0+0+n*3) += c * (B(m,2)*B(n,2)*D(4,5) + B(m,2)*B(n,1)*D(4,3) + s*B(m,2)*B(n,0)*D(4,0) + B(m,0)*B(n,2)*D(3,5) + B(m,0)*B(n,1)*D(3,3) + s*B(m,0)*B(n,0)*D(3,0) + s*B(m,1)*B(n,2)*D(1,5) + s*B(m,1)*B(n,1)*D(1,3) + 2.0*B(m,1)*B(n,0)*D(1,0)) / 2.0;
                KK(ii0+1+m*3,jj0+1+n*3) += c * (B(m,2)*B(n,2)*D(4,4) + B(m,2)*B(n,0)*D(4,3) + s*B(m,2)*B(n,1)*D(4,1) + B(m,0)*B(n,2)*D(3,4) + B(m,0)*B(n,0)*D(3,3) + s*B(m,0)*B(n,1)*D(3,1) + s*B(m,1)*B(n,2)*D(1,4) + s*B(m,1)*B(n,0)*D(1,3) + 2.0*B(m,1)*B(n,1)*D(1,1)) / 2.0;
                KK(ii0+1+m*3,jj0+2+n*3) += c * (B(m,2)*B(n,0)*D(4,5) + B(m,2)*B(n,1)*D(4,4) + s*B(m,2)*B(n,2)*D(4,2) + B(m,0)*B(n,0)*D(3,5) + B(m,0)*B(n,1)*D(3,4) + s*B(m,0)*B(n,2)*D(3,2) + s*B(m,1)*B(n,0)*D(1,5) + s*B(m,1)*B(n,1)*D(1,4) + 2.0*B(m,1)*B(n,2)*D(1,2)) / 2.0;
                KK(ii0+2+m*3,jj0+0+n*3) += c * (B(m,0)*B(n,2)*D(5,5) + B(m,0)*B(n,1)*D(5,3) + s*B(m,0)*B(n,0)*D(5,0) + B(m,1)*B(n,2)*D(4,5) + B(m,1)*B(n,1)*D(4,3) + s*B(m,1)*B(n,0)*D(4,0) + s*B(m,2)*B(n,2)*D(2,5) + s*B(m,2)*B(n,1)*D(2,3) + 2.0*B(m,2)*B(n,0)*D(2,0)) / 2.0;
                KK(ii0+2+m*3,jj0+1+n*3) += c * (B(m,0)*B(n,2)*D(5,4) + B(m,0)*B(n,0)*D(5,3) + s*B(m,0)*B(n,1)*D(5,1) + B(m,1)*B(n,2)*D(4,4) + B(m,1)*B(n,0)*D(4,3) + s*B(m,1)*B(n,1)*D(4,1) + s*B(m,2)*B(n,2)*D(2,4) + s*B(m,2)*B(n,0)*D(2,3) + 2.0*B(m,2)*B(n,1)*D(2,1)) / 2.0;
                KK(ii0+2+m*3,jj0+2+n*3) += c * (B(m,0)*B(n,0)*D(5,5) + B(m,0)*B(n,1)*D(5,4) + s*B(m,0)*B(n,2)*D(5,2) + B(m,1)*B(n,0)*D(4,5) + B(m,1)*B(n,1)*D(4,4) + s*B(m,1)*B(n,2)*D(4,2) + s*B(m,2)*B(n,0)*D(2,5) + s*B(m,2)*B(n,1)*D(2,4) + 2.0*B(m,2)*B(n,2)*D(2,2)) / 2.0;
            }
        }
    }
}

/* add_to_kk_axisymmetric (src/integ/mat_10_bdb.rs:213-233) */
static void add_to_kk_axisymmetric(double *kk, int nrow, int nnode, double c, double r, const double *dd, const go_args *args) {
    const double s = SQRT_2;
    const go_pad *pad = args->pad;
    const int sd = 2, ns = 4;
    const double *nn = pad->interp;
    const int ii0 = args->ii0, jj0 = args->jj0;
    for (int m = 0; m < nnode; m++) {
        for (int n = 0; n < nnode; n++) {
            KK(ii0+0+m*2,jj0+0+n*2) += c * r * (B(m,1)*B(n,1)*D(3,3) + s*B(m,1)*B(n,0)*D(3,0) + s*B(m,0)*B(n,1)*D(0,3) + 2.0*B(m,0)*B(n,0)*D(0,0)) / 2.0 +
                                       c * (r*nn[n]*(2.0*D(0,2)*B(m,0) + s*D(3,2)*B(m,1)) + nn[m]*(2.0*nn[n]*D(2,2) + 2.0*r*D(2,0)*B(n,0) + s*r*D(2,3)*B(n,1))) / (2.0*r);

            KK(ii0+0+m*2,jj0+1+n*2) += c * r * (B(m,1)*B(n,0)*D(3,3) + s*B(m,1)*B(n,1)*D(3,1) + s*B(m,0)*B(n,0)*D(0,3) + 2.0*B(m,0)*B(n,1)*D(0,1)) / 2.0 +
                                       c * (nn[m]*(s*D(2,3)*B(n,0) + 2.0*D(2,1)*B(n,1))) / 2.0;

            KK(ii0+1+m*2,jj0+0+n*2) += c * r * (B(m,0)*B(n,1)*D(3,3) + s*B(m,0)*B(n,0)*D(3,0) + s*B(m,1)*B(n,1)*D(1,3) + 2.0*B(m,1)*B(n,0)*D(1,0)) / 2.0 +
                                       c * (nn[n]*(s*D(3,2)*B(m,0) + 2.0*D(1,2)*B(m,1))) / 2.0;

            KK(ii0+1+m*2,jj0+1+n*2) += c * r * (B(m,0)*B(n,0)*D(3,3) + s*B(m,0)*B(n,1)*D(3,1) + s*B(m,1)*B(n,0)*D(1,3) + 2.0*B(m,1)*B(n,1)*D(1,1)) / 2.0;
        }
    }
}

static double radius_at_ip(const go_pad *pad) {
    /* r = sum_m N[m] * Xt[0][m]  (mat_10_bdb.rs:164-167 and siblings) */
    const int nnode = pad->nnode;
    double r = 0.0;
    for (int m = 0; m < nnode; m++) r += pad->interp[m] * XXT(0, m);
    return r;
}

/* integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-174) */
int go_mat_10_bdb(double *kk, int nrow, int ncol, go_args *args, go_fn_dd fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    if (nrow < args->ii0 + nnode * sd) return GO_ERR_NROW_KK_NDOF;
    if (ncol < args->jj0 + nnode * sd) return GO_ERR_NCOL_KK_NDOF;
    if (args->axisymmetric && sd != 2) return GO_ERR_AXISYM_NDIM;
    double dd[36];
    memset(dd, 0, sizeof(dd)); /* Tensor4::new */
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_gradient(pad, iota, &det_jac);
        if (err) return err;
        if (fn(dd, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        const double c = det_jac * weight * args->alpha;
        if (args->axisymmetric) {
            double r = radius_at_ip(pad);
            add_to_kk_axisymmetric(kk, nrow, nnode, c, r, dd, args);
        } else {
            add_to_kk(kk, nrow, sd, nnode, c, dd, args);
        }
    }
    return GO_OK;
}

/* calc_bb_matrix (src/integ/mat_10_bdb.rs:238-278); bb_mat is (ns x ndof) row-major, zero-initialised by the caller */
static double calc_bb_matrix(double *bb_mat, const go_pad *pad, int axisymmetric) {
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ndof = nnode * sd;
    double radius = 1.0;
#define BM(a, col) bb_mat[(a)*ndof + (col)]
    if (sd == 3) {
        for (int i = 0; i < nnode; i++) {
            BM(0, 0 + i * 3) = B(i, 0);
            BM(1, 1 + i * 3) = B(i, 1);
            BM(2, 2 + i * 3) = B(i, 2);
            BM(3, 0 + i * 3) = B(i, 1) / SQRT_2;
            BM(4, 1 + i * 3) = B(i, 2) / SQRT_2;
            BM(5, 2 + i * 3) = B(i, 0) / SQRT_2;
            BM(3, 1 + i * 3) = B(i, 0) / SQRT_2;
            BM(4, 2 + i * 3) = B(i, 1) / SQRT_2;
            BM(5, 0 + i * 3) = B(i, 2) / SQRT_2;
        }
    } else {
        if (axisymmetric) {
            radius = radius_at_ip(pad);
            for (int i = 0; i < nnode; i++) {
                BM(0, 0 + i * 2) = B(i, 0);
                BM(1, 1 + i * 2) = B(i, 1);
                BM(2, 0 + i * 2) = pad->interp[i] / radius;
                BM(3, 0 + i * 2) = B(i, 1) / SQRT_2;
                BM(3, 1 + i * 2) = B(i, 0) / SQRT_2;
            }
        } else {
            for (int i = 0; i < nnode; i++) {
                BM(0, 0 + i * 2) = B(i, 0);
                BM(1, 1 + i * 2) = B(i, 1);
                BM(3, 0 + i * 2) = B(i, 1) / SQRT_2;
                BM(3, 1 + i * 2) = B(i, 0) / SQRT_2;
            }
        }
    }
#undef BM
    return radius;
}

/* integ::mat_10_bdb_alt (src/integ/mat_10_bdb.rs:304-354) with add_to_stiff_mat (:288-301).
 * Note: the reference's add_to_stiff_mat asserts kk.dims() == (ndof, ndof); offsets are not applied. */
int go_mat_10_bdb_alt(double *kk, int nrow, int ncol, go_args *args, go_fn_dd fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ndof = nnode * sd;
    if (nrow < args->ii0 + ndof) return GO_ERR_NROW_KK_NDOF;
    if (ncol < args->jj0 + ndof) return GO_ERR_NCOL_KK_NDOF;
    if (args->axisymmetric && sd != 2) return GO_ERR_AXISYM_NDIM;
    const int ns = 2 * sd;
    double dd[36];
    double bb_mat[6 * 3 * GO_MAX_NNODE];
    memset(dd, 0, sizeof(dd));
    memset(bb_mat, 0, sizeof(bb_mat));
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_gradient(pad, iota, &det_jac);
        if (err) return err;
        if (fn(dd, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        const double radius = calc_bb_matrix(bb_mat, pad, args->axisymmetric);
        const double coef = det_jac * weight * args->alpha * radius;
        for (int i = 0; i < ndof; i++)
            for (int j = 0; j < ndof; j++)
                for (int k = 0; k < ns; k++)
                    for (int l = 0; l < ns; l++) KK(i, j) += coef * bb_mat[k * ndof + i] * D(k, l) * bb_mat[l * ndof + j];
    }
    return GO_OK;
}

/* integ::mat_01_nsn (src/integ/mat_01_nsn.rs:48-102) */
int go_mat_01_nsn(double *kk, int nrow, int ncol, go_args *args, go_fn_s fn, void *user) {
    go_pad *pad = args->pad;
    const int nnode = pad->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode) return GO_ERR_NROW_KK_NNODE;
    if (ncol < jj0 + nnode) return GO_ERR_NCOL_KK_NNODE;
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_gradient(pad, iota, &det_jac);
        if (err) return err;
        const double *nn = pad->interp;
        double s;
        if (fn(&s, p, nn, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        double c;
        if (args->axisymmetric) {
            double r = radius_at_ip(pad);
            c = s * det_jac * weight * args->alpha * r;
        } else {
            c = s * det_jac * weight * args->alpha;
        }
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode; n++) KK(ii0 + m, jj0 + n) += nn[m] * c * nn[n];
    }
    return GO_OK;
}

/* integ::mat_03_btb (src/integ/mat_03_btb.rs:50-131) */
int go_mat_03_btb(double *kk, int nrow, int ncol, go_args *args, go_fn_tt fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode) return GO_ERR_NROW_KK_NNODE;
    if (ncol < jj0 + nnode) return GO_ERR_NCOL_KK_NNODE;
    double t[6] = {0, 0, 0, 0, 0, 0}; /* Tensor2::new */
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    const double s = SQRT_2;
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_gradient(pad, iota, &det_jac);
        if (err) return err;
        if (fn(t, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        double c;
        if (args->axisymmetric) {
            double r = radius_at_ip(pad);
            c = det_jac * weight * args->alpha * r;
        } else {
            c = det_jac * weight * args->alpha;
        }
        if (sd == 2) {
            for (int m = 0; m < nnode; m++)
                for (int n = 0; n < nnode; n++)
                    KK(ii0 + m, jj0 + n) += c * (B(n, 1) * (t[1] * B(m, 1) + (t[3] * B(m, 0)) / s) + B(n, 0) * (t[0] * B(m, 0) + (t[3] * B(m, 1)) / s));
        } else {
            for (int m = 0; m < nnode; m++)
                for (int n = 0; n < nnode; n++)
                    KK(ii0 + m, jj0 + n) += c * (B(n, 2) * (t[2] * B(m, 2) + (t[5] * B(m, 0)) / s + (t[4] * B(m, 1)) / s) +
                                                B(n, 1) * (t[1] * B(m, 1) + (t[3] * B(m, 0)) / s + (t[4] * B(m, 2)) / s) +
                                                B(n, 0) * (t[0] * B(m, 0) + (t[3] * B(m, 1)) / s + (t[5] * B(m, 2)) / s));
        }
    }
    return GO_OK;
}

/* shared prologue of the single-pad matrix cases: N, B, det J and the coefficient c = det J * w * alpha [* r] */
static int matrix_case_point(go_args *args, int p, double *c) {
    go_pad *pad = args->pad;
    const double *iota = args->gauss[p];
    const double weight = args->gauss[p][3];
    go_pad_calc_interp(pad, iota);
    double det_jac;
    int err = go_pad_calc_gradient(pad, iota, &det_jac);
    if (err) return err;
    if (args->axisymmetric) {
        double r = radius_at_ip(pad);
        *c = det_jac * weight * args->alpha * r;
    } else {
        *c = det_jac * weight * args->alpha;
    }
    return GO_OK;
}

/* integ::mat_02_bvn (src/integ/mat_02_bvn.rs:48-120): K(m,n) = sum_p (B(m).v) N(n) c_p */
int go_mat_02_bvn(double *kk, int nrow, int ncol, go_args *args, go_fn_v fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode) return GO_ERR_NROW_KK_NNODE;
    if (ncol < jj0 + nnode) return GO_ERR_NCOL_KK_NNODE;
    double v[3] = {0, 0, 0};
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        double c;
        int err = matrix_case_point(args, p, &c);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(v, p, nn, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode; n++) {
                if (sd == 2) KK(ii0 + m, jj0 + n) += c * (B(m, 0) * v[0] + B(m, 1) * v[1]) * nn[n];
                else KK(ii0 + m, jj0 + n) += c * (B(m, 0) * v[0] + B(m, 1) * v[1] + B(m, 2) * v[2]) * nn[n];
            }
    }
    return GO_OK;
}

/* integ::mat_08_ntn (src/integ/mat_08_ntn.rs:56-135): K(m i, n j) = sum_p N(m) T(i,j) N(n) c_p, T from its Mandel vector */
int go_mat_08_ntn(double *kk, int nrow, int ncol, go_args *args, go_fn_tt fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode * sd) return GO_ERR_NROW_KK_NDOF;
    if (ncol < jj0 + nnode * sd) return GO_ERR_NCOL_KK_NDOF;
    double t[6] = {0, 0, 0, 0, 0, 0};
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    const double s = SQRT_2;
    for (int p = 0; p < args->ngauss; p++) {
        double c;
        int err = matrix_case_point(args, p, &c);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(t, p, nn, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode; n++) {
                if (sd == 2) {
                    KK(ii0 + 0 + m * 2, jj0 + 0 + n * 2) += c * nn[m] * t[0] * nn[n];
                    KK(ii0 + 0 + m * 2, jj0 + 1 + n * 2) += c * nn[m] * t[3] * nn[n] / s;
                    KK(ii0 + 1 + m * 2, jj0 + 0 + n * 2) += c * nn[m] * t[3] * nn[n] / s;
                    KK(ii0 + 1 + m * 2, jj0 + 1 + n * 2) += c * nn[m] * t[1] * nn[n];
                } else {
                    KK(ii0 + 0 + m * 3, jj0 + 0 + n * 3) += c * nn[m] * t[0] * nn[n];
                    KK(ii0 + 0 + m * 3, jj0 + 1 + n * 3) += c * nn[m] * t[3] * nn[n] / s;
                    KK(ii0 + 0 + m * 3, jj0 + 2 + n * 3) += c * nn[m] * t[5] * nn[n] / s;
                    KK(ii0 + 1 + m * 3, jj0 + 0 + n * 3) += c * nn[m] * t[3] * nn[n] / s;
                    KK(ii0 + 1 + m * 3, jj0 + 1 + n * 3) += c * nn[m] * t[1] * nn[n];
                    KK(ii0 + 1 + m * 3, jj0 + 2 + n * 3) += c * nn[m] * t[4] * nn[n] / s;
                    KK(ii0 + 2 + m * 3, jj0 + 0 + n * 3) += c * nn[m] * t[5] * nn[n] / s;
                    KK(ii0 + 2 + m * 3, jj0 + 1 + n * 3) += c * nn[m] * t[4] * nn[n] / s;
                    KK(ii0 + 2 + m * 3, jj0 + 2 + n * 3) += c * nn[m] * t[2] * nn[n];
                }
            }
    }
    return GO_OK;
}

/* integ::mat_09_nvb (src/integ/mat_09_nvb.rs:54-125): K(m i, n j) = sum_p N(m) v(i) B(n,j) c_p */
int go_mat_09_nvb(double *kk, int nrow, int ncol, go_args *args, go_fn_v fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode * sd) return GO_ERR_NROW_KK_NDOF;
    if (ncol < jj0 + nnode * sd) return GO_ERR_NCOL_KK_NDOF;
    double v[3] = {0, 0, 0};
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        double c;
        int err = matrix_case_point(args, p, &c);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(v, p, nn, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode; n++)
                for (int i = 0; i < sd; i++)
                    for (int j = 0; j < sd; j++) KK(ii0 + i + m * sd, jj0 + j + n * sd) += c * nn[m] * v[i] * B(n, j);
    }
    return GO_OK;
}

/* integ::vec_01_ns (src/integ/vec_01_ns.rs:83-130); works for CABLE cells too (calc_jacobian only) */
int go_vec_01_ns(double *a, int len, go_args *args, go_fn_s fn, void *user) {
    go_pad *pad = args->pad;
    const int nnode = pad->nnode;
    const int ii0 = args->ii0;
    if (len < ii0 + nnode) return GO_ERR_A_LEN;
    if (args->clear) memset(a, 0, sizeof(double) * (size_t)len);
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_jacobian(pad, iota, &det_jac);
        if (err) return err;
        const double *nn = pad->interp;
        double s;
        if (fn(&s, p, nn, NULL, user)) return GO_ERR_CALLBACK_STOP;
        double coef;
        if (args->axisymmetric) {
            double r = radius_at_ip(pad);
            coef = s * det_jac * weight * args->alpha * r;
        } else {
            coef = s * det_jac * weight * args->alpha;
        }
        for (int m = 0; m < nnode; m++) a[ii0 + m] += nn[m] * coef;
    }
    return GO_OK;
}

/* integ::vec_02_nv (src/integ/vec_02_nv.rs:85-140) */
int go_vec_02_nv(double *b, int len, go_args *args, go_fn_v fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0;
    if (len < ii0 + nnode * sd) return GO_ERR_B_LEN;
    double v[3] = {0, 0, 0};
    if (args->clear) memset(b, 0, sizeof(double) * (size_t)len);
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_jacobian(pad, iota, &det_jac);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(v, p, nn, NULL, user)) return GO_ERR_CALLBACK_STOP;
        double coef;
        if (args->axisymmetric) {
            double r = radius_at_ip(pad);
            coef = det_jac * weight * args->alpha * r;
        } else {
            coef = det_jac * weight * args->alpha;
        }
        if (sd == 2) {
            for (int m = 0; m < nnode; m++) {
                b[ii0 + 0 + m * 2] += coef * nn[m] * v[0];
                b[ii0 + 1 + m * 2] += coef * nn[m] * v[1];
            }
        } else {
            for (int m = 0; m < nnode; m++) {
                b[ii0 + 0 + m * 3] += coef * nn[m] * v[0];
                b[ii0 + 1 + m * 3] += coef * nn[m] * v[1];
                b[ii0 + 2 + m * 3] += coef * nn[m] * v[2];
            }
        }
    }
    return GO_OK;
}

/* integ::vec_03_bv (src/integ/vec_03_bv.rs:84-141) */
int go_vec_03_bv(double *c, int len, go_args *args, go_fn_v fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0;
    if (len < ii0 + nnode) return GO_ERR_C_LEN;
    double w[3] = {0, 0, 0};
    if (args->clear) memset(c, 0, sizeof(double) * (size_t)len);
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_gradient(pad, iota, &det_jac);
        if (err) return err;
        if (fn(w, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        double coef;
        if (args->axisymmetric) {
            double r = radius_at_ip(pad);
            coef = det_jac * weight * args->alpha * r;
        } else {
            coef = det_jac * weight * args->alpha;
        }
        if (sd == 2) {
            for (int m = 0; m < nnode; m++) c[ii0 + m] += coef * (w[0] * B(m, 0) + w[1] * B(m, 1));
        } else {
            for (int m = 0; m < nnode; m++) c[ii0 + m] += coef * (w[0] * B(m, 0) + w[1] * B(m, 1) + w[2] * B(m, 2));
        }
    }
    return GO_OK;
}

/* integ::vec_04_bt (src/integ/vec_04_bt.rs:93-170) */
int go_vec_04_bt(double *d, int len, go_args *args, go_fn_tt fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode;
    const int ii0 = args->ii0;
    if (len < ii0 + nnode * sd) return GO_ERR_D_LEN;
    if (args->axisymmetric && sd != 2) return GO_ERR_AXISYM_NDIM;
    double t[6] = {0, 0, 0, 0, 0, 0};
    if (args->clear) memset(d, 0, sizeof(double) * (size_t)len);
    const double s = SQRT_2;
    for (int p = 0; p < args->ngauss; p++) {
        const double *iota = args->gauss[p];
        const double weight = args->gauss[p][3];
        go_pad_calc_interp(pad, iota);
        double det_jac;
        int err = go_pad_calc_gradient(pad, iota, &det_jac);
        if (err) return err;
        if (fn(t, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        const double c = det_jac * weight * args->alpha;
        if (args->axisymmetric) {
            const double *nn = pad->interp;
            double r = radius_at_ip(pad);
            for (int m = 0; m < nnode; m++) {
                d[ii0 + 0 + m * 2] += c * r * (t[0] * B(m, 0) + t[3] * B(m, 1) / s) + c * nn[m] * t[2];
                d[ii0 + 1 + m * 2] += c * r * (t[3] * B(m, 0) / s + t[1] * B(m, 1));
            }
        } else if (sd == 2) {
            for (int m = 0; m < nnode; m++) {
                d[ii0 + 0 + m * 2] += c * (t[0] * B(m, 0) + t[3] * B(m, 1) / s);
                d[ii0 + 1 + m * 2] += c * (t[3] * B(m, 0) / s + t[1] * B(m, 1));
            }
        } else {
            for (int m = 0; m < nnode; m++) {
                d[ii0 + 0 + m * 3] += c * (t[0] * B(m, 0) + t[3] * B(m, 1) / s + t[5] * B(m, 2) / s);
                d[ii0 + 1 + m * 3] += c * (t[3] * B(m, 0) / s + t[1] * B(m, 1) + t[4] * B(m, 2) / s);
                d[ii0 + 2 + m * 3] += c * (t[5] * B(m, 0) / s + t[4] * B(m, 1) / s + t[2] * B(m, 2));
            }
        }
    }
    return GO_OK;
}

/* integ::scalar_field (src/integ/scalar_field.rs) */
int go_scalar_field(double *result, go_pad *pad, const double (*gauss)[4], int ngauss, go_fn_s fn, void *user) {
    double ii = 0.0;
    for (int p = 0; p < ngauss; p++) {
        const double *iota = gauss[p];
        const double weight = gauss[p][3];
        double det_jac;
        int err = go_pad_calc_jacobian(pad, iota, &det_jac);
        if (err) return err;
        double s;
        if (fn(&s, p, NULL, NULL, user)) return GO_ERR_CALLBACK_STOP;
        ii += s * det_jac * weight;
    }
    *result = ii;
    return GO_OK;
}

#undef B
#undef D
#undef KK
#undef XXT

/* ---------------------------------------------------------------------------------------------- */
/* russell_tensor::LinElasticity -> Mandel modulus (pinned by the Tet4,    */
/* Bhatti and Felippa known-answer tests of src/integ/mat_10_bdb.rs)                              */
/* ---------------------------------------------------------------------------------------------- */

void go_lin_elasticity(double young, double poisson, int two_dim, int plane_stress, double *dd) {
    const int ns = two_dim ? 4 : 6;
    memset(dd, 0, sizeof(double) * (size_t)(ns * ns));
#define DD(a, b) dd[(a) + (b)*ns]
    if (two_dim && plane_stress) {
        const double c = young / (1.0 - poisson * poisson);
        DD(0, 0) = c;
        DD(0, 1) = c * poisson;
        DD(1, 0) = c * poisson;
        DD(1, 1) = c;
        DD(3, 3) = c * (1.0 - poisson);
    } else {
        const double c = young / ((1.0 + poisson) * (1.0 - 2.0 * poisson));
        for (int a = 0; a < 3; a++)
            for (int b = 0; b < 3; b++) DD(a, b) = (a == b) ? c * (1.0 - poisson) : c * poisson;
        for (int q = 3; q < ns; q++) DD(q, q) = c * (1.0 - 2.0 * poisson);
    }
#undef DD
}

/* ---------------------------------------------------------------------------------------------- */
/* mesh-wide loop                                                                                 */
/* ---------------------------------------------------------------------------------------------- */

size_t go_op_out_size(int op, int kind, int ndim) {
    const size_t nnode = (size_t)go_kind_nnode(kind);
    const size_t ndof = nnode * (size_t)ndim;
    switch (op) {
    case GO_OP_MAT_10_BDB: case GO_OP_MAT_10_BDB_ALT: case GO_OP_MAT_08_NTN: case GO_OP_MAT_09_NVB: return ndof * ndof;
    case GO_OP_MAT_01_NSN: case GO_OP_MAT_03_BTB: case GO_OP_MAT_02_BVN: return nnode * nnode;
    case GO_OP_VEC_01_NS: case GO_OP_VEC_03_BV: return nnode;
    case GO_OP_VEC_02_NV: case GO_OP_VEC_04_BT: return ndof;
    case GO_OP_JACOBIAN: return 1;
    default: return 0;
    }
}

size_t go_op_coef_size(int op, int ndim) {
    const size_t ns = 2 * (size_t)ndim;
    switch (op) {
    case GO_OP_MAT_10_BDB: case GO_OP_MAT_10_BDB_ALT: return ns * ns;
    case GO_OP_MAT_01_NSN: case GO_OP_VEC_01_NS: return 1;
    case GO_OP_MAT_03_BTB: case GO_OP_VEC_04_BT: case GO_OP_MAT_08_NTN: return ns;
    case GO_OP_VEC_02_NV: case GO_OP_VEC_03_BV: case GO_OP_MAT_02_BVN: case GO_OP_MAT_09_NVB: return (size_t)ndim;
    default: return 0;
    }
}

typedef struct {
    const double *base;  /* coefficient of this cell at gauss point 0 */
    size_t gp_stride;
    size_t len;
} coef_view;

static int cb_copy(double *dst, int p, const double *nn, const double *bb, void *user) {
    (void)nn; (void)bb;
    const coef_view *cv = (const coef_view *)user;
    memcpy(dst, cv->base + (size_t)p * cv->gp_stride, sizeof(double) * cv->len);
    return 0;
}

static int integ_one_cell(int op, const go_mesh *mesh, size_t e, size_t cell_begin, int ngauss, const double *coef,
                          size_t coef_cell_stride, size_t coef_gp_stride, double alpha, int axisymmetric, const double *ksi,
                          double *out) {
    const int kind = mesh->kinds[e];
    go_pad pad;
    int err = go_pad_init(&pad, mesh->ndim, kind);
    if (err) return err;
    /* Mesh::set_pad (src/mesh/mesh.rs:448-456) */
    const uint64_t *points = mesh->conn + mesh->conn_off[e];
    for (int m = 0; m < pad.nnode; m++)
        for (int j = 0; j < mesh->ndim; j++) go_pad_set_xx(&pad, m, j, mesh->coords[points[m] * (size_t)mesh->ndim + (size_t)j]);
    if (op == GO_OP_JACOBIAN) return go_pad_calc_jacobian(&pad, ksi, out);
    go_args args;
    args.pad = &pad;
    err = go_gauss(kind, ngauss, &args.gauss, &args.ngauss);
    if (err) return err;
    args.clear = 1;
    args.axisymmetric = axisymmetric;
    args.alpha = alpha;
    args.ii0 = 0;
    args.jj0 = 0;
    coef_view cv;
    cv.base = coef + (e - cell_begin) * coef_cell_stride;
    cv.gp_stride = coef_gp_stride;
    cv.len = go_op_coef_size(op, mesh->ndim);
    const int nnode = pad.nnode, ndof = pad.nnode * mesh->ndim;
    switch (op) {
    case GO_OP_MAT_10_BDB: return go_mat_10_bdb(out, ndof, ndof, &args, cb_copy, &cv);
    case GO_OP_MAT_10_BDB_ALT: return go_mat_10_bdb_alt(out, ndof, ndof, &args, cb_copy, &cv);
    case GO_OP_MAT_01_NSN: return go_mat_01_nsn(out, nnode, nnode, &args, cb_copy, &cv);
    case GO_OP_MAT_03_BTB: return go_mat_03_btb(out, nnode, nnode, &args, cb_copy, &cv);
    case GO_OP_MAT_02_BVN: return go_mat_02_bvn(out, nnode, nnode, &args, cb_copy, &cv);
    case GO_OP_MAT_08_NTN: return go_mat_08_ntn(out, ndof, ndof, &args, cb_copy, &cv);
    case GO_OP_MAT_09_NVB: return go_mat_09_nvb(out, ndof, ndof, &args, cb_copy, &cv);
    case GO_OP_VEC_01_NS: return go_vec_01_ns(out, nnode, &args, cb_copy, &cv);
    case GO_OP_VEC_02_NV: return go_vec_02_nv(out, ndof, &args, cb_copy, &cv);
    case GO_OP_VEC_03_BV: return go_vec_03_bv(out, nnode, &args, cb_copy, &cv);
    case GO_OP_VEC_04_BT: return go_vec_04_bt(out, ndof, &args, cb_copy, &cv);
    default: return GO_ERR_KIND_UNSUPPORTED;
    }
}

int go_mesh_integ(int op, const go_mesh *mesh, size_t cell_begin, size_t cell_end, int ngauss, const double *coef,
                  size_t coef_cell_stride, size_t coef_gp_stride, double alpha, int axisymmetric, const double *ksi, double *out,
                  const uint64_t *out_off, int nthreads, size_t *err_cell) {
    int first_err = GO_OK;
    size_t first_cell = (size_t)-1;
    if (out_off == NULL) {
        /* homogeneous fast path requires equal sizes; otherwise the caller must pass offsets */
        const size_t sz = (cell_end > cell_begin) ? go_op_out_size(op, mesh->kinds[cell_begin], mesh->ndim) : 0;
        for (size_t e = cell_begin; e < cell_end; e++)
            if (go_op_out_size(op, mesh->kinds[e], mesh->ndim) != sz) return GO_ERR_KIND_UNSUPPORTED;
    }
    const long long n = (long long)(cell_end - cell_begin);
    if (nthreads <= 1) {
        for (long long i = 0; i < n; i++) {
            const size_t e = cell_begin + (size_t)i;
            const size_t sz = go_op_out_size(op, mesh->kinds[e], mesh->ndim);
            double *dst = out + (out_off ? out_off[i] : (size_t)i * sz);
            int err = integ_one_cell(op, mesh, e, cell_begin, ngauss, coef, coef_cell_stride, coef_gp_stride, alpha, axisymmetric, ksi, dst);
            if (err && e < first_cell) { first_err = err; first_cell = e; }
        }
    } else {
#ifdef _OPENMP
#pragma omp parallel for schedule(static) num_threads(nthreads)
#endif
        for (long long i = 0; i < n; i++) {
            const size_t e = cell_begin + (size_t)i;
            const size_t sz = go_op_out_size(op, mesh->kinds[e], mesh->ndim);
            double *dst = out + (out_off ? out_off[i] : (size_t)i * sz);
            int err = integ_one_cell(op, mesh, e, cell_begin, ngauss, coef, coef_cell_stride, coef_gp_stride, alpha, axisymmetric, ksi, dst);
            if (err) {
#ifdef _OPENMP
#pragma omp critical
#endif
                { if (e < first_cell) { first_err = err; first_cell = e; } }
            }
        }
    }
    if (err_cell) *err_cell = first_cell;
    return first_err;
}


/* ---------------------------------------------------------------------------------------------- */
/* two-pad coupling matrices and boundary integrals                                               */
/* ---------------------------------------------------------------------------------------------- */

#define B(m, j) (pad->gradient[(m)*sd + (j)])
#define KK(i, j) (kk[(size_t)(i) + (size_t)(j) * (size_t)nrow])
#define XXT(j, m) (pad->xxt[(j)*nnode + (m)])
#define BB(m, j) (pad_b->gradient[(m)*sd + (j)])

/* the coefficient c of the coupled cases: N, B and det J of the main pad, then c = det J * w * alpha [* r] */
static int coupled_point(go_args *args, int p, double *c) { return matrix_case_point(args, p, c); }

/* integ::mat_04_nsb (src/integ/mat_04_nsb.rs:66-137): K(m, n j) = sum_p Nb(m) s c_p B(n, j) */
int go_mat_04_nsb(double *kk, int nrow, int ncol, go_pad *pad_b, go_args *args, go_fn_s fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode, nnode_b = pad_b->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode_b) return GO_ERR_NROW_KK_NNODE_B;
    if (ncol < jj0 + nnode * sd) return GO_ERR_NCOL_KK_NDOF_PAD;
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        go_pad_calc_interp(pad_b, args->gauss[p]); /* Nb */
        double c0;
        int err = coupled_point(args, p, &c0);
        if (err) return err;
        double s = 0.0;
        if (fn(&s, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        const double c = s * c0; /* s * det_jac * weight * alpha [* r]: same product up to the order of the scalar factors */
        const double *nnb = pad_b->interp;
        for (int m = 0; m < nnode_b; m++)
            for (int n = 0; n < nnode; n++)
                for (int j = 0; j < sd; j++) KK(ii0 + m, jj0 + j + n * sd) += nnb[m] * c * B(n, j);
    }
    return GO_OK;
}

/* integ::mat_05_btn (src/integ/mat_05_btn.rs:68-168): K(m, n j) = sum_p (Bb(m) . T)_j N(n) c_p, T from its Mandel vector */
int go_mat_05_btn(double *kk, int nrow, int ncol, go_pad *pad_b, go_args *args, go_fn_tt fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode, nnode_b = pad_b->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode_b) return GO_ERR_NROW_KK_NNODE_B;
    if (ncol < jj0 + nnode * sd) return GO_ERR_NCOL_KK_NDOF_PAD;
    double t[6] = {0, 0, 0, 0, 0, 0};
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    const double s = SQRT_2;
    for (int p = 0; p < args->ngauss; p++) {
        double det_b;
        int err = go_pad_calc_gradient(pad_b, args->gauss[p], &det_b); /* Bb */
        if (err) return err;
        double c;
        err = coupled_point(args, p, &c);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(t, p, nn, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        for (int m = 0; m < nnode_b; m++)
            for (int n = 0; n < nnode; n++) {
                if (sd == 2) {
                    KK(ii0 + m, jj0 + 0 + n * 2) += c * (BB(m, 0) * t[0] + BB(m, 1) * t[3] / s) * nn[n];
                    KK(ii0 + m, jj0 + 1 + n * 2) += c * (BB(m, 0) * t[3] / s + BB(m, 1) * t[1]) * nn[n];
                } else {
                    KK(ii0 + m, jj0 + 0 + n * 3) += c * (BB(m, 0) * t[0] + BB(m, 1) * t[3] / s + BB(m, 2) * t[5] / s) * nn[n];
                    KK(ii0 + m, jj0 + 1 + n * 3) += c * (BB(m, 0) * t[3] / s + BB(m, 1) * t[1] + BB(m, 2) * t[4] / s) * nn[n];
                    KK(ii0 + m, jj0 + 2 + n * 3) += c * (BB(m, 0) * t[5] / s + BB(m, 1) * t[4] / s + BB(m, 2) * t[2]) * nn[n];
                }
            }
    }
    return GO_OK;
}

/* integ::mat_06_nvn (src/integ/mat_06_nvn.rs:69-140): K(m i, n) = sum_p N(m) v_i Nb(n) c_p */
int go_mat_06_nvn(double *kk, int nrow, int ncol, go_args *args, go_pad *pad_b, go_fn_v fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode, nnode_b = pad_b->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode * sd) return GO_ERR_NROW_KK_NDOF_PAD;
    if (ncol < jj0 + nnode_b) return GO_ERR_NCOL_KK_NNODE_B;
    double v[3] = {0, 0, 0};
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        double c;
        int err = coupled_point(args, p, &c);
        if (err) return err;
        go_pad_calc_interp(pad_b, args->gauss[p]); /* Nb */
        const double *nn = pad->interp, *nnb = pad_b->interp;
        if (fn(v, p, nn, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode_b; n++)
                for (int i = 0; i < sd; i++) KK(ii0 + i + m * sd, jj0 + n) += c * nn[m] * v[i] * nnb[n];
    }
    return GO_OK;
}

/* integ::mat_07_bsn (src/integ/mat_07_bsn.rs:69-135): K(m i, n) = sum_p B(m, i) s c_p Nb(n) */
int go_mat_07_bsn(double *kk, int nrow, int ncol, go_args *args, go_pad *pad_b, go_fn_s fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode, nnode_b = pad_b->nnode;
    const int ii0 = args->ii0, jj0 = args->jj0;
    if (nrow < ii0 + nnode * sd) return GO_ERR_NROW_KK_NDOF_PAD;
    if (ncol < jj0 + nnode_b) return GO_ERR_NCOL_KK_NNODE_B;
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        double c0;
        int err = coupled_point(args, p, &c0);
        if (err) return err;
        go_pad_calc_interp(pad_b, args->gauss[p]); /* Nb */
        double s = 0.0;
        if (fn(&s, p, pad->interp, pad->gradient, user)) return GO_ERR_CALLBACK_STOP;
        const double c = s * c0;
        const double *nnb = pad_b->interp;
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode_b; n++)
                for (int i = 0; i < sd; i++) KK(ii0 + i + m * sd, jj0 + n) += B(m, i) * c * nnb[n];
    }
    return GO_OK;
}

/* Scratchpad::calc_normal_vector (src/shapes/scratchpad_calc_normal_vector.rs:111-176) */
int go_pad_calc_normal_vector(go_pad *pad, double *un, const double *ksi, double *mag_n) {
    const int sd = pad->space_ndim, gd = pad->geo_ndim, nnode = pad->nnode;
    if (sd == 2 && gd != 1) return GO_ERR_NORMAL_2D;
    if (sd == 3 && gd != 2) return GO_ERR_NORMAL_3D;
    go_deriv(pad->kind, ksi, pad->deriv);
    /* J = Xt . L (space_ndim x geo_ndim) */
    double jj[3][2] = {{0, 0}, {0, 0}, {0, 0}};
    for (int i = 0; i < sd; i++)
        for (int j = 0; j < gd; j++) {
            double sum = 0.0;
            for (int m = 0; m < nnode; m++) sum += XXT(i, m) * pad->deriv[m * gd + j];
            jj[i][j] = sum;
        }
    if (sd == 2) { /* n = e3 x g1 */
        un[0] = -jj[1][0];
        un[1] = jj[0][0];
        const double mag = sqrt(un[0] * un[0] + un[1] * un[1]);
        if (mag > 0.0) {
            un[0] /= mag;
            un[1] /= mag;
        }
        *mag_n = mag;
        return GO_OK;
    }
    /* n = g1 x g2 */
    un[0] = jj[1][0] * jj[2][1] - jj[2][0] * jj[1][1];
    un[1] = jj[2][0] * jj[0][1] - jj[0][0] * jj[2][1];
    un[2] = jj[0][0] * jj[1][1] - jj[1][0] * jj[0][1];
    const double mag = sqrt(un[0] * un[0] + un[1] * un[1] + un[2] * un[2]);
    if (mag > 0.0) {
        un[0] /= mag;
        un[1] /= mag;
        un[2] /= mag;
    }
    *mag_n = mag;
    return GO_OK;
}

static int bry_checks(const go_pad *pad) {
    if (pad->space_ndim == 2 && pad->geo_ndim != 1) return GO_ERR_BRY_2D;
    if (pad->space_ndim == 3 && pad->geo_ndim != 2) return GO_ERR_BRY_3D;
    return GO_OK;
}

/* N, unit normal and c = mag_n * w * alpha [* r] at Gauss point p of a boundary cell */
static int bry_point(go_args *args, int p, double *un, double *c) {
    go_pad *pad = args->pad;
    go_pad_calc_interp(pad, args->gauss[p]);
    double mag_n;
    int err = go_pad_calc_normal_vector(pad, un, args->gauss[p], &mag_n);
    if (err) return err;
    const double weight = args->gauss[p][3];
    *c = args->axisymmetric ? mag_n * weight * args->alpha * radius_at_ip(pad) : mag_n * weight * args->alpha;
    return GO_OK;
}

/* integ::mat_01_nsn_bry (src/integ/mat_01_nsn_bry.rs:48-113) */
int go_mat_01_nsn_bry(double *kk, int nrow, int ncol, go_args *args, go_fn_bry fn, void *user) {
    go_pad *pad = args->pad;
    const int nnode = pad->nnode, ii0 = args->ii0, jj0 = args->jj0;
    int err = bry_checks(pad);
    if (err) return err;
    if (nrow < ii0 + nnode) return GO_ERR_NROW_KK_NNODE;
    if (ncol < jj0 + nnode) return GO_ERR_NCOL_KK_NNODE;
    if (args->clear) memset(kk, 0, sizeof(double) * (size_t)nrow * (size_t)ncol);
    for (int p = 0; p < args->ngauss; p++) {
        double un[3] = {0, 0, 0}, c0, s = 0.0;
        err = bry_point(args, p, un, &c0);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(&s, p, un, nn, user)) return GO_ERR_CALLBACK_STOP;
        const double c = s * c0;
        for (int m = 0; m < nnode; m++)
            for (int n = 0; n < nnode; n++) KK(ii0 + m, jj0 + n) += nn[m] * c * nn[n];
    }
    return GO_OK;
}

/* integ::vec_01_ns_bry (src/integ/vec_01_ns_bry.rs:54-115) */
int go_vec_01_ns_bry(double *a, int len, go_args *args, go_fn_bry fn, void *user) {
    go_pad *pad = args->pad;
    const int nnode = pad->nnode, ii0 = args->ii0;
    int err = bry_checks(pad);
    if (err) return err;
    if (len < ii0 + nnode) return GO_ERR_A_LEN;
    if (args->clear) memset(a, 0, sizeof(double) * (size_t)len);
    for (int p = 0; p < args->ngauss; p++) {
        double un[3] = {0, 0, 0}, c0, s = 0.0;
        err = bry_point(args, p, un, &c0);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(&s, p, un, nn, user)) return GO_ERR_CALLBACK_STOP;
        const double coef = s * c0;
        for (int m = 0; m < nnode; m++) a[ii0 + m] += nn[m] * coef;
    }
    return GO_OK;
}

/* integ::vec_02_nv_bry (src/integ/vec_02_nv_bry.rs:61-130) */
int go_vec_02_nv_bry(double *b, int len, go_args *args, go_fn_bry fn, void *user) {
    go_pad *pad = args->pad;
    const int sd = pad->space_ndim, nnode = pad->nnode, ii0 = args->ii0;
    int err = bry_checks(pad);
    if (err) return err;
    if (len < ii0 + nnode * sd) return GO_ERR_B_LEN;
    if (args->clear) memset(b, 0, sizeof(double) * (size_t)len);
    for (int p = 0; p < args->ngauss; p++) {
        double un[3] = {0, 0, 0}, coef, v[3] = {0, 0, 0};
        err = bry_point(args, p, un, &coef);
        if (err) return err;
        const double *nn = pad->interp;
        if (fn(v, p, un, nn, user)) return GO_ERR_CALLBACK_STOP;
        for (int m = 0; m < nnode; m++)
            for (int i = 0; i < sd; i++) b[ii0 + i + m * sd] += coef * nn[m] * v[i];
    }
    return GO_OK;
}

size_t go_op_out_size_ext(int op, int kind, int kind_b, int ndim, int ngauss) {
    const size_t nnode = (size_t)go_kind_nnode(kind), ndof = nnode * (size_t)ndim;
    const size_t nb = kind_b >= 0 ? (size_t)go_kind_nnode(kind_b) : 0;
    switch (op) {
    case GO_OP_MAT_04_NSB: case GO_OP_MAT_05_BTN: case GO_OP_MAT_06_NVN: case GO_OP_MAT_07_BSN: return nb * ndof;
    case GO_OP_MAT_01_NSN_BRY: return nnode * nnode;
    case GO_OP_VEC_01_NS_BRY: return nnode;
    case GO_OP_VEC_02_NV_BRY: return ndof;
    case GO_OP_NORMAL: {
        const double(*g)[4];
        int ng = 0;
        if (go_gauss(kind, ngauss, &g, &ng)) return 0;
        return (size_t)ng * (size_t)(ndim + 1);
    }
    default: return 0;
    }
}

size_t go_op_coef_size_ext(int op, int ndim) {
    switch (op) {
    case GO_OP_MAT_04_NSB: case GO_OP_MAT_07_BSN: return 1;
    case GO_OP_MAT_05_BTN: return 2 * (size_t)ndim;
    case GO_OP_MAT_06_NVN: return (size_t)ndim;
    case GO_OP_MAT_01_NSN_BRY: case GO_OP_VEC_01_NS_BRY: case GO_OP_VEC_02_NV_BRY: return (size_t)ndim + 1;
    default: return 0;
    }
}

/* data form of the boundary closures: s = s0 + w.un (scalar cases), v = v + qn un (vec_02_nv_bry) */
typedef struct {
    coef_view cv;
    int ndim, vector_case;
} bry_view;

static int cb_bry(double *out, int p, const double *un, const double *nn, void *user) {
    (void)nn;
    const bry_view *bv = (const bry_view *)user;
    const double *r = bv->cv.base + (size_t)p * bv->cv.gp_stride;
    if (bv->vector_case) {
        for (int i = 0; i < bv->ndim; i++) out[i] = r[i] + r[bv->ndim] * un[i];
    } else {
        double s = r[0];
        for (int i = 0; i < bv->ndim; i++) s += r[1 + i] * un[i];
        out[0] = s;
    }
    return 0;
}

static int integ_one_cell_ext(int op, const go_mesh *mesh, size_t e, size_t cell_begin, int ngauss, int kind_b, const double *coef,
                              size_t coef_cell_stride, size_t coef_gp_stride, double alpha, int axisymmetric, double *out) {
    const int kind = mesh->kinds[e], ndim = mesh->ndim;
    go_pad pad, pad_b;
    int err = go_pad_init(&pad, ndim, kind);
    if (err) return err;
    const uint64_t *points = mesh->conn + mesh->conn_off[e];
    for (int m = 0; m < pad.nnode; m++)
        for (int j = 0; j < ndim; j++) go_pad_set_xx(&pad, m, j, mesh->coords[points[m] * (size_t)ndim + (size_t)j]);
    go_args args;
    args.pad = &pad;
    err = go_gauss(kind, ngauss, &args.gauss, &args.ngauss);
    if (err) return err;
    args.clear = 1;
    args.axisymmetric = axisymmetric;
    args.alpha = alpha;
    args.ii0 = 0;
    args.jj0 = 0;
    const int nnode = pad.nnode, ndof = pad.nnode * ndim;
    if (op == GO_OP_NORMAL) {
        for (int p = 0; p < args.ngauss; p++) {
            double un[3] = {0, 0, 0}, mag;
            err = go_pad_calc_normal_vector(&pad, un, args.gauss[p], &mag);
            if (err) return err;
            for (int i = 0; i < ndim; i++) out[p * (ndim + 1) + i] = un[i];
            out[p * (ndim + 1) + ndim] = mag;
        }
        return GO_OK;
    }
    coef_view cv;
    cv.base = coef + (e - cell_begin) * coef_cell_stride;
    cv.gp_stride = coef_gp_stride;
    cv.len = go_op_coef_size_ext(op, ndim);
    if (op >= GO_OP_MAT_01_NSN_BRY) {
        bry_view bv;
        bv.cv = cv;
        bv.ndim = ndim;
        bv.vector_case = op == GO_OP_VEC_02_NV_BRY;
        switch (op) {
        case GO_OP_MAT_01_NSN_BRY: return go_mat_01_nsn_bry(out, nnode, nnode, &args, cb_bry, &bv);
        case GO_OP_VEC_01_NS_BRY: return go_vec_01_ns_bry(out, nnode, &args, cb_bry, &bv);
        default: return go_vec_02_nv_bry(out, ndof, &args, cb_bry, &bv);
        }
    }
    /* the second pad: kind_b on the first nnode_b nodes of the cell */
    err = go_pad_init(&pad_b, ndim, kind_b);
    if (err) return err;
    if (pad_b.nnode > pad.nnode || pad_b.geo_ndim != pad.geo_ndim) return GO_ERR_KIND_UNSUPPORTED;
    for (int m = 0; m < pad_b.nnode; m++)
        for (int j = 0; j < ndim; j++) go_pad_set_xx(&pad_b, m, j, mesh->coords[points[m] * (size_t)ndim + (size_t)j]);
    const int nb = pad_b.nnode;
    switch (op) {
    case GO_OP_MAT_04_NSB: return go_mat_04_nsb(out, nb, ndof, &pad_b, &args, cb_copy, &cv);
    case GO_OP_MAT_05_BTN: return go_mat_05_btn(out, nb, ndof, &pad_b, &args, cb_copy, &cv);
    case GO_OP_MAT_06_NVN: return go_mat_06_nvn(out, ndof, nb, &args, &pad_b, cb_copy, &cv);
    case GO_OP_MAT_07_BSN: return go_mat_07_bsn(out, ndof, nb, &args, &pad_b, cb_copy, &cv);
    default: return GO_ERR_KIND_UNSUPPORTED;
    }
}

int go_mesh_integ_ext(int op, const go_mesh *mesh, size_t cell_begin, size_t cell_end, int ngauss, int kind_b, const double *coef,
                      size_t coef_cell_stride, size_t coef_gp_stride, double alpha, int axisymmetric, double *out, int nthreads,
                      size_t *err_cell) {
    int first_err = GO_OK;
    size_t first_cell = (size_t)-1;
    (void)nthreads;
    size_t off = 0;
    for (size_t e = cell_begin; e < cell_end; e++) {
        const size_t sz = go_op_out_size_ext(op, mesh->kinds[e], kind_b, mesh->ndim, ngauss);
        int err = integ_one_cell_ext(op, mesh, e, cell_begin, ngauss, kind_b, coef, coef_cell_stride, coef_gp_stride, alpha, axisymmetric,
                                     out + off);
        if (err && e < first_cell) {
            first_err = err;
            first_cell = e;
        }
        off += sz;
    }
    if (err_cell) *err_cell = first_cell;
    return first_err;
}

/*
 * gemlab_oracle.h -- CPU restatement of gemlab's per-cell integration path.
 *
 * TEST INFRASTRUCTURE ONLY.  This library is the parity oracle and the timed CPU baseline for the
 * CUDA implementation in gemlab_b200/.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load it.  The product path never calls into it.
 *
 * It restates, in the reference's own operation order, the Rust code of cpmech/gemlab v2.4.0
 * (paths relative to the reference tree):
 *   shapes      src/shapes/{lin2,lin3,tri3,tri6,qua4,qua8,qua9,tet4,tet10,hex8,hex20}.rs
 *   Scratchpad  src/shapes/scratchpad.rs:152-299, scratchpad_calc_jacobian.rs:101-131,
 *               scratchpad_calc_gradient.rs:78-91, scratchpad_calc_coords.rs
 *   Gauss       src/integ/gauss.rs:87-246 (selection), :317-774 (tables, generated .inc)
 *   integ       src/integ/mat_10_bdb.rs:121-354, mat_01_nsn.rs:48-102, mat_03_btb.rs:50-131,
 *               vec_01_ns.rs:83-130, vec_02_nv.rs:85-140, vec_03_bv.rs:84-141, vec_04_bt.rs:93-170,
 *               scalar_field.rs
 *   Mesh        src/mesh/mesh.rs:448-456 (set_pad gather)
 * Third-party pieces that are not in the reference tree (russell_lab / russell_tensor ^2.8, no
 * Cargo.lock) are restated from their published algorithms: closed-form mat_inverse for 1x1/2x2/3x3
 * with the |det| <= 1e-15 failure, plain triple-loop mat_mat_mul, and the Mandel-basis elastic
 * modulus of LinElasticity.  Parity is pinned by the reference's own known-answer tests
 * (tests/golden/, tests/test_oracle_goldens.py); Hex8/Hex20/Tet10/Qua8-stiffness element matrices
 * are NOT pinned by any reference test ("parity unpinned" for those).
 *
 * Matrices crossing this interface are COLUMN-MAJOR (russell_lab::Matrix storage) unless noted.
 */
#ifndef GEMLAB_ORACLE_H
#define GEMLAB_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GO_MAX_NNODE 32

/* GeoKind ordinals follow GeoKind::VALUES (src/shapes/enums.rs:1467-1493) */
enum {
    GO_LIN2 = 0, GO_LIN3, GO_LIN4, GO_LIN5,
    GO_TRI3, GO_TRI6, GO_TRI10, GO_TRI15,
    GO_QUA4, GO_QUA8, GO_QUA9, GO_QUA12, GO_QUA16, GO_QUA17,
    GO_TET4, GO_TET10, GO_TET20,
    GO_HEX8, GO_HEX20, GO_HEX32,
    GO_NKIND
};
enum { GO_CLASS_LIN = 0, GO_CLASS_TRI, GO_CLASS_QUA, GO_CLASS_TET, GO_CLASS_HEX };

/* error codes; go_strerror returns the reference's StrError strings */
enum {
    GO_OK = 0,
    GO_ERR_XXT_NOT_SET = 1,        /* scratchpad_calc_jacobian.rs:104 */
    GO_ERR_ZERO_DET = 2,           /* russell_lab mat_inverse */
    GO_ERR_GRADIENT_NDIM = 3,      /* scratchpad_calc_gradient.rs:82 */
    GO_ERR_NROW_KK_NDOF = 4,       /* mat_10_bdb.rs:129 */
    GO_ERR_NCOL_KK_NDOF = 5,       /* mat_10_bdb.rs:132 */
    GO_ERR_AXISYM_NDIM = 6,        /* mat_10_bdb.rs:135 */
    GO_ERR_NROW_KK_NNODE = 7,      /* mat_01_nsn.rs:57 */
    GO_ERR_NCOL_KK_NNODE = 8,      /* mat_01_nsn.rs:60 */
    GO_ERR_A_LEN = 9,              /* vec_01_ns.rs:91 */
    GO_ERR_B_LEN = 10,             /* vec_02_nv.rs:93 */
    GO_ERR_C_LEN = 11,             /* vec_03_bv.rs:92 */
    GO_ERR_D_LEN = 12,             /* vec_04_bt.rs:100 */
    GO_ERR_GAUSS_LIN = 13, GO_ERR_GAUSS_TRI = 14, GO_ERR_GAUSS_QUA = 15,
    GO_ERR_GAUSS_TET = 16, GO_ERR_GAUSS_HEX = 17, /* gauss.rs:197-240 */
    GO_ERR_KIND_UNSUPPORTED = 18,  /* oracle does not restate this GeoKind yet */
    GO_ERR_CALLBACK_STOP = 19,     /* the user callback returned an error ("stop" in the tests) */
    GO_ERR_SPACE_NDIM = 20,        /* scratchpad.rs:153-161 */
    GO_ERR_NROW_KK_NNODE_B = 21,   /* mat_04_nsb.rs:78 "nrow(K) must be >= ii0 + pad_b.nnode" */
    GO_ERR_NCOL_KK_NDOF_PAD = 22,  /* mat_04_nsb.rs:81 "ncol(K) must be >= jj0 + pad.nnode . space_ndim" */
    GO_ERR_NROW_KK_NDOF_PAD = 23,  /* mat_06_nvn.rs:81 "nrow(K) must be >= ii0 + pad.nnode . space_ndim" */
    GO_ERR_NCOL_KK_NNODE_B = 24,   /* mat_06_nvn.rs:84 "ncol(K) must be >= jj0 + pad_b.nnode" */
    GO_ERR_BRY_2D = 25,            /* mat_01_nsn_bry.rs:56 "in 2D, geometry ndim must be equal to 1 (a line)" */
    GO_ERR_BRY_3D = 26,            /* mat_01_nsn_bry.rs:59 "in 3D, geometry ndim must be equal to 2 (a surface)" */
    GO_ERR_NORMAL_2D = 27,         /* scratchpad_calc_normal_vector.rs:115 */
    GO_ERR_NORMAL_3D = 28          /* scratchpad_calc_normal_vector.rs:118 */
};
const char *go_strerror(int code);

/* --- GeoKind (src/shapes/enums.rs) ------------------------------------------------------- */
int go_kind_nnode(int kind);
int go_kind_geo_ndim(int kind);
int go_kind_class(int kind);
const char *go_kind_name(int kind);
int go_kind_from_name(const char *name);
int go_kind_supported(int kind);
/* reference coordinates of local node m (geo_ndim values written to ksi) */
int go_kind_reference_coords(int kind, int m, double *ksi);
int go_interp(int kind, const double *ksi, double *nn);       /* N (nnode) */
int go_deriv(int kind, const double *ksi, double *ll);        /* L (nnode x geo_ndim), row-major */

/* --- Gauss (src/integ/gauss.rs) ---------------------------------------------------------- */
/* ngauss == 0 selects Gauss::new(kind); otherwise Gauss::new_sized(kind.class(), ngauss) */
int go_gauss(int kind, int ngauss, const double (**data)[4], int *npoint);

/* --- Scratchpad (src/shapes/scratchpad.rs:61-116) ----------------------------------------- */
typedef struct {
    int kind, space_ndim, geo_ndim, nnode, ok_xxt;
    double xxt[3 * GO_MAX_NNODE];          /* (space_ndim x nnode): xxt[j*nnode+m] */
    double interp[GO_MAX_NNODE];           /* N */
    double deriv[3 * GO_MAX_NNODE];        /* L (nnode x geo_ndim): deriv[m*geo_ndim+j] */
    double jacobian[9];                    /* J (space_ndim x geo_ndim): [i*geo_ndim+j] */
    double inv_jacobian[9];                /* (space_ndim x space_ndim): [i*space_ndim+j] */
    double gradient[3 * GO_MAX_NNODE];     /* B (nnode x space_ndim): [m*space_ndim+j] */
} go_pad;

int go_pad_init(go_pad *pad, int space_ndim, int kind);
void go_pad_set_xx(go_pad *pad, int m, int j, double value);
void go_pad_calc_interp(go_pad *pad, const double *ksi);
int go_pad_calc_jacobian(go_pad *pad, const double *ksi, double *det_jac);
int go_pad_calc_gradient(go_pad *pad, const double *ksi, double *det_jac);
void go_pad_calc_coords(go_pad *pad, const double *ksi, double *x);

/* --- CommonArgs (src/integ/common_args.rs) ------------------------------------------------ */
typedef struct {
    go_pad *pad;
    int kind_for_gauss; /* unused helper */
    const double (*gauss)[4];
    int ngauss;
    int clear;
    int axisymmetric;
    double alpha;
    int ii0, jj0;
} go_args;

/* callbacks return 0 on success, nonzero to abort (the integ function returns GO_ERR_CALLBACK_STOP) */
typedef int (*go_fn_dd)(double *dd /* ns x ns Mandel, column-major */, int p, const double *nn, const double *bb, void *user);
typedef int (*go_fn_tt)(double *tt /* ns Mandel vector */, int p, const double *nn, const double *bb, void *user);
typedef int (*go_fn_s)(double *s, int p, const double *nn, const double *bb, void *user);
typedef int (*go_fn_v)(double *v /* space_ndim */, int p, const double *nn, const double *bb, void *user);

/* kk is column-major with leading dimension nrow */
int go_mat_10_bdb(double *kk, int nrow, int ncol, go_args *args, go_fn_dd fn, void *user);
int go_mat_10_bdb_alt(double *kk, int nrow, int ncol, go_args *args, go_fn_dd fn, void *user);
int go_mat_01_nsn(double *kk, int nrow, int ncol, go_args *args, go_fn_s fn, void *user);
int go_mat_02_bvn(double *kk, int nrow, int ncol, go_args *args, go_fn_v fn, void *user);
int go_mat_08_ntn(double *kk, int nrow, int ncol, go_args *args, go_fn_tt fn, void *user);
int go_mat_09_nvb(double *kk, int nrow, int ncol, go_args *args, go_fn_v fn, void *user);
int go_mat_03_btb(double *kk, int nrow, int ncol, go_args *args, go_fn_tt fn, void *user);
int go_vec_01_ns(double *a, int len, go_args *args, go_fn_s fn, void *user);
int go_vec_02_nv(double *b, int len, go_args *args, go_fn_v fn, void *user);
int go_vec_03_bv(double *c, int len, go_args *args, go_fn_v fn, void *user);
int go_vec_04_bt(double *d, int len, go_args *args, go_fn_tt fn, void *user);
int go_scalar_field(double *result, go_pad *pad, const double (*gauss)[4], int ngauss, go_fn_s fn, void *user);

/* russell_tensor::LinElasticity::new(E, nu, two_dim, plane_stress).get_modulus().matrix():
 * ns x ns Mandel matrix (ns = 4 if two_dim else 6), column-major (symmetric anyway) */
void go_lin_elasticity(double young, double poisson, int two_dim, int plane_stress, double *dd);

/* --- mesh-wide loop (the user loop of examples/integ_mom_inertia_disk.rs:21-33) -------------
 * op codes */
enum {
    GO_OP_MAT_10_BDB = 0, GO_OP_MAT_01_NSN = 1, GO_OP_MAT_03_BTB = 2,
    GO_OP_VEC_01_NS = 3, GO_OP_VEC_03_BV = 4, GO_OP_VEC_02_NV = 5, GO_OP_VEC_04_BT = 6,
    GO_OP_MAT_10_BDB_ALT = 7, GO_OP_JACOBIAN = 8,
    GO_OP_MAT_02_BVN = 9, GO_OP_MAT_08_NTN = 10, GO_OP_MAT_09_NVB = 11,
    /* two-pad coupling matrices (second pad = kind_b on the first nnode_b nodes of the cell) */
    GO_OP_MAT_04_NSB = 12, GO_OP_MAT_05_BTN = 13, GO_OP_MAT_06_NVN = 14, GO_OP_MAT_07_BSN = 15,
    /* boundary integrals over cells with geo_ndim = space_ndim - 1 */
    GO_OP_MAT_01_NSN_BRY = 16, GO_OP_VEC_01_NS_BRY = 17, GO_OP_VEC_02_NV_BRY = 18,
    GO_OP_NORMAL = 19 /* calc_normal_vector at every Gauss point: ng x (un[space_ndim], mag_n) */
};

/* Scratchpad::calc_normal_vector (src/shapes/scratchpad_calc_normal_vector.rs:111-176): unit normal and its magnitude */
int go_pad_calc_normal_vector(go_pad *pad, double *un, const double *ksi, double *mag_n);

/* callbacks of the boundary cases receive the unit normal: fn(out, p, un, nn, user) */
typedef int (*go_fn_bry)(double *out, int p, const double *un, const double *nn, void *user);
/* integ::mat_04_nsb .. mat_07_bsn (src/integ/mat_04_nsb.rs:66, mat_05_btn.rs:68, mat_06_nvn.rs:69, mat_07_bsn.rs:69) */
int go_mat_04_nsb(double *kk, int nrow, int ncol, go_pad *pad_b, go_args *args, go_fn_s fn, void *user);
int go_mat_05_btn(double *kk, int nrow, int ncol, go_pad *pad_b, go_args *args, go_fn_tt fn, void *user);
int go_mat_06_nvn(double *kk, int nrow, int ncol, go_args *args, go_pad *pad_b, go_fn_v fn, void *user);
int go_mat_07_bsn(double *kk, int nrow, int ncol, go_args *args, go_pad *pad_b, go_fn_s fn, void *user);
/* integ::mat_01_nsn_bry, vec_01_ns_bry, vec_02_nv_bry (src/integ/mat_01_nsn_bry.rs:48, vec_01_ns_bry.rs:54, vec_02_nv_bry.rs:61) */
int go_mat_01_nsn_bry(double *kk, int nrow, int ncol, go_args *args, go_fn_bry fn, void *user);
int go_vec_01_ns_bry(double *a, int len, go_args *args, go_fn_bry fn, void *user);
int go_vec_02_nv_bry(double *b, int len, go_args *args, go_fn_bry fn, void *user);

/* per-element output size in doubles for op on a cell of `kind` in `ndim` space */
size_t go_op_out_size(int op, int kind, int ndim);
/* coefficient record length in doubles: ns*ns (mat_10), 1 (mat_01, vec_01), ns (mat_03, vec_04), ndim (vec_02/03) */
size_t go_op_coef_size(int op, int ndim);

typedef struct {
    int ndim;
    size_t npoint, ncell;
    const double *coords;      /* AoS npoint x ndim (Mesh.points[i].coords flattened) */
    const uint8_t *kinds;      /* ncell GeoKind ordinals */
    const uint64_t *conn_off;  /* ncell+1 */
    const uint64_t *conn;      /* point ids */
} go_mesh;

/* Integrate cells [cell_begin, cell_end): per cell Mesh::set_pad gather, then integ::<op>.
 * coef: base pointer; coefficient of (cell e, gauss point p) is
 *       coef + (e - cell_begin) * coef_cell_stride + p * coef_gp_stride   (strides in doubles; 0 = uniform)
 * out:  element e is written at out + out_off[e - cell_begin] (column-major, tight);
 *       if out_off == NULL elements are packed contiguously in cell order.
 * ksi:  only for GO_OP_JACOBIAN (det J at one reference point, e.g. check_jacobian's [1/3,1/3,1/3])
 * nthreads <= 1: sequential (the reference's execution model); otherwise OpenMP over cells.
 * Returns 0 or the first error code; *err_cell receives the smallest failing cell id. */
int go_mesh_integ(int op, const go_mesh *mesh, size_t cell_begin, size_t cell_end, int ngauss,
                  const double *coef, size_t coef_cell_stride, size_t coef_gp_stride, double alpha,
                  int axisymmetric, const double *ksi, double *out, const uint64_t *out_off,
                  int nthreads, size_t *err_cell);

/* the mesh loop for the cases above.  kind_b: GeoKind of the second pad (its nodes are the first nnode_b nodes of every cell);
 * coefficient records: mat_04 / mat_07 s; mat_05 Mandel vector; mat_06 v; the boundary cases take (s0, w[space_ndim]) with
 * s = s0 + w.un, and vec_02_nv_bry (v[space_ndim], qn) with v = v + qn un -- the data form of closures that read the normal */
int go_mesh_integ_ext(int op, const go_mesh *mesh, size_t cell_begin, size_t cell_end, int ngauss, int kind_b, const double *coef,
                      size_t coef_cell_stride, size_t coef_gp_stride, double alpha, int axisymmetric, double *out, int nthreads,
                      size_t *err_cell);
size_t go_op_out_size_ext(int op, int kind, int kind_b, int ndim, int ngauss);
size_t go_op_coef_size_ext(int op, int ndim);

#ifdef __cplusplus
}
#endif
#endif

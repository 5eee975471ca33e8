/*
 * gemlab_cuda.h -- C ABI of libgemlab_cuda: batched, B200-native (sm_100a) Gauss-quadrature
 * integration of element matrices and vectors over a cell range of a gemlab mesh.
 *
 * This is the drop-in boundary for gemlab's per-cell integration path.  The reference
 * (cpmech/gemlab v2.4.0, pure Rust) exposes no FFI for this path: its boundary is the generic Rust
 * signature `integ::X(out, &mut CommonArgs, closure)` called once per cell by user code
 * (examples/integ_mom_inertia_disk.rs:21-33, pmsim-style assemblers).  Each entry point below names the
 * reference item it replaces (paths relative to the reference tree).  A `gemlab-cuda` Rust crate binds these
 * symbols (gemlab-cuda/src/ffi.rs) and offers `integ::batch::*` taking `&Mesh` + a
 * cell range.
 *
 * Conventions
 *   - plain C types only; every pointer is either a HOST pointer or a raw CUDA DEVICE pointer, as
 *     stated by the accompanying `location` field.  No torch / C++ types cross this boundary.
 *   - all functions return 0 on success or a stable nonzero gl_status; gl_last_error(ctx) returns
 *     the reference's own StrError string for that status (tests compare the strings).
 *   - GeoKind ordinals follow GeoKind::VALUES (src/shapes/enums.rs:1467-1493).
 *   - matrices are COLUMN-MAJOR (russell_lab::Matrix storage): K(ii,jj) at ii + jj*nrow, with the
 *     reference's index layout ii = ii0 + i + m*space_ndim, jj = jj0 + j + n*space_ndim
 *     (src/integ/mat_10_bdb.rs:31-46).
 *   - a ctx belongs to ONE device and ONE host thread at a time (Send, not Sync); calls are
 *     stream-ordered on the ctx stream and asynchronous when every buffer is a device buffer;
 *     gl_ctx_sync() waits and reports device-side failures (zero Jacobian determinant).
 *   - there is no CPU fallback: every entry point fails with GL_ERR_CUDA when no sm_100 device is
 *     usable.
 */
#ifndef GEMLAB_CUDA_H
#define GEMLAB_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define GL_API
#else
#define GL_API __attribute__((visibility("default")))
#endif

typedef struct gl_ctx gl_ctx;   /* one per GPU: stream, reference-element tables, scratch buffers */
typedef struct gl_mesh gl_mesh; /* device-resident SoA mesh, bucketed by GeoKind */

typedef enum gl_status {
    GL_OK = 0,
    GL_ERR_XXT_NOT_SET = 1,   /* "all components of the coordinates matrix must be set first" (scratchpad_calc_jacobian.rs:104) */
    GL_ERR_ZERO_DET = 2,      /* "cannot compute inverse due to zero determinant" (russell_lab::mat_inverse) */
    GL_ERR_GRADIENT_NDIM = 3, /* "calc_gradient requires that geo_ndim = space_ndim" (scratchpad_calc_gradient.rs:82) */
    GL_ERR_NROW_KK_NDOF = 4,  /* "nrow(K) must be >= ii0 + nnode * space_ndim" (mat_10_bdb.rs:129) */
    GL_ERR_NCOL_KK_NDOF = 5,  /* "ncol(K) must be >= jj0 + nnode * space_ndim" (mat_10_bdb.rs:132) */
    GL_ERR_AXISYM_NDIM = 6,   /* "axisymmetric requires space_ndim = 2" (mat_10_bdb.rs:135) */
    GL_ERR_NROW_KK_NNODE = 7, /* "nrow(K) must be >= ii0 + nnode" (mat_01_nsn.rs:57) */
    GL_ERR_NCOL_KK_NNODE = 8, /* "ncol(K) must be >= jj0 + nnode" (mat_01_nsn.rs:60) */
    GL_ERR_A_LEN = 9,         /* "a.len() must be >= ii0 + nnode" (vec_01_ns.rs:91) */
    GL_ERR_B_LEN = 10,        /* "b.len() must be >= ii0 + nnode * space_ndim" (vec_02_nv.rs:93) */
    GL_ERR_C_LEN = 11,        /* "c.len() must be >= ii0 + nnode" (vec_03_bv.rs:92) */
    GL_ERR_D_LEN = 12,        /* "d.len() must be >= ii0 + nnode * space_ndim" (vec_04_bt.rs:100) */
    GL_ERR_GAUSS_LIN = 13,    /* "requested number of integration points is not available for Lin class" (gauss.rs:197) */
    GL_ERR_GAUSS_TRI = 14,
    GL_ERR_GAUSS_QUA = 15,
    GL_ERR_GAUSS_TET = 16,
    GL_ERR_GAUSS_HEX = 17,
    GL_ERR_KIND_UNSUPPORTED = 18, /* unknown GeoKind ordinal, or shape functions at an arbitrary point requested for a table-only kind */
    GL_ERR_SPACE_NDIM = 20,       /* "space_ndim must be 2 or 3" (scratchpad.rs:153) */
    GL_ERR_NROW_KK_NNODE_B = 21,  /* "nrow(K) must be >= ii0 + pad_b.nnode" (mat_04_nsb.rs:78) */
    GL_ERR_NCOL_KK_NDOF_PAD = 22, /* "ncol(K) must be >= jj0 + pad.nnode * space_ndim" (mat_04_nsb.rs:81) */
    GL_ERR_NROW_KK_NDOF_PAD = 23, /* "nrow(K) must be >= ii0 + pad.nnode * space_ndim" (mat_06_nvn.rs:81) */
    GL_ERR_NCOL_KK_NNODE_B = 24,  /* "ncol(K) must be >= jj0 + pad_b.nnode" (mat_06_nvn.rs:84) */
    GL_ERR_BRY_2D = 25,           /* "in 2D, geometry ndim must be equal to 1 (a line)" (mat_01_nsn_bry.rs:56) */
    GL_ERR_BRY_3D = 26,           /* "in 3D, geometry ndim must be equal to 2 (a surface)" (mat_01_nsn_bry.rs:59) */
    GL_ERR_NORMAL_2D = 27,        /* "calc_normal_vector requires geo_ndim = 1 in 2D (CABLE in 2D)" (scratchpad_calc_normal_vector.rs:115) */
    GL_ERR_NORMAL_3D = 28,        /* "calc_normal_vector requires geo_ndim = 2 in 3D (SHELL in 3D)" (scratchpad_calc_normal_vector.rs:118) */
    /* errors of the batched boundary itself (no reference counterpart) */
    GL_ERR_INVALID_ARG = 100,
    GL_ERR_CELL_RANGE = 101,
    GL_ERR_OUT_CAPACITY = 102,
    GL_ERR_COEF = 103,
    GL_ERR_MESH = 104,
    GL_ERR_CUDA = 200,
    GL_ERR_NO_DEVICE = 201
} gl_status;

/* integration cases; numbering follows the reference's file names (src/integ/mod.rs:304-353) */
typedef enum gl_op {
    GL_OP_MAT_10_BDB = 0, /* integ::mat_10_bdb  K = int B.D.B       (src/integ/mat_10_bdb.rs:121)  ndof x ndof */
    GL_OP_MAT_01_NSN = 1, /* integ::mat_01_nsn  K = int N s N       (src/integ/mat_01_nsn.rs:48)   nnode x nnode */
    GL_OP_MAT_03_BTB = 2, /* integ::mat_03_btb  K = int B.T.B       (src/integ/mat_03_btb.rs:50)   nnode x nnode */
    GL_OP_VEC_01_NS = 3,  /* integ::vec_01_ns   a = int N s         (src/integ/vec_01_ns.rs:83)    nnode */
    GL_OP_VEC_03_BV = 4,  /* integ::vec_03_bv   c = int B.w         (src/integ/vec_03_bv.rs:84)    nnode */
    GL_OP_VEC_02_NV = 5,  /* integ::vec_02_nv   b = int N v         (src/integ/vec_02_nv.rs:85)    ndof */
    GL_OP_VEC_04_BT = 6,  /* integ::vec_04_bt   d = int B.sigma     (src/integ/vec_04_bt.rs:93)    ndof */
    GL_OP_JACOBIAN = 8,   /* Scratchpad::calc_jacobian at one ksi (scratchpad_calc_jacobian.rs:101; Mesh::check_jacobian, mesh/check.rs:43-60)  1 */
    GL_OP_MAT_02_BVN = 9, /* integ::mat_02_bvn  K = int (B.v) N     (src/integ/mat_02_bvn.rs:48)   nnode x nnode; coefficient v (space_ndim) */
    GL_OP_MAT_08_NTN = 10, /* integ::mat_08_ntn K = int N T N       (src/integ/mat_08_ntn.rs:56)   ndof x ndof;   coefficient Mandel vector (ns) */
    GL_OP_MAT_09_NVB = 11, /* integ::mat_09_nvb K = int N v B       (src/integ/mat_09_nvb.rs:54)   ndof x ndof;   coefficient v (space_ndim) */
    /* two-pad coupling matrices: the second Scratchpad (pad_b) is gl_args.kind_b on the FIRST nnode_b nodes of every cell
     * (Qua8 + Qua4, Tri6 + Tri3, Hex20 + Hex8, Tet10 + Tet4, or the cell's own kind) */
    GL_OP_MAT_04_NSB = 12, /* integ::mat_04_nsb K = int Nb s B      (src/integ/mat_04_nsb.rs:66)   nnode_b x ndof; coefficient s */
    GL_OP_MAT_05_BTN = 13, /* integ::mat_05_btn K = int Bb.T N      (src/integ/mat_05_btn.rs:68)   nnode_b x ndof; coefficient Mandel vector (ns) */
    GL_OP_MAT_06_NVN = 14, /* integ::mat_06_nvn K = int N v Nb      (src/integ/mat_06_nvn.rs:69)   ndof x nnode_b; coefficient v (space_ndim) */
    GL_OP_MAT_07_BSN = 15, /* integ::mat_07_bsn K = int B s Nb      (src/integ/mat_07_bsn.rs:69)   ndof x nnode_b; coefficient s */
    /* boundary integrals over cells with geo_ndim = space_ndim - 1 (lines in 2D, surfaces in 3D).  The reference's closures
     * receive the unit normal un; as data, a record is (s0, w[space_ndim]) with s = s0 + w.un for the two scalar cases and
     * (v[space_ndim], qn) with v = v + qn un for vec_02_nv_bry (a traction plus a pressure-like normal load) */
    GL_OP_MAT_01_NSN_BRY = 16, /* integ::mat_01_nsn_bry (src/integ/mat_01_nsn_bry.rs:48)  nnode x nnode */
    GL_OP_VEC_01_NS_BRY = 17,  /* integ::vec_01_ns_bry  (src/integ/vec_01_ns_bry.rs:54)   nnode */
    GL_OP_VEC_02_NV_BRY = 18,  /* integ::vec_02_nv_bry  (src/integ/vec_02_nv_bry.rs:61)   ndof */
    GL_OP_NORMAL = 19          /* Scratchpad::calc_normal_vector at every Gauss point (scratchpad_calc_normal_vector.rs:111):
                                  ngauss x (un[space_ndim], mag_n) per cell */
} gl_op;

typedef enum gl_location { GL_HOST = 0, GL_DEVICE = 1 } gl_location;

/* Mirrors integ::CommonArgs (src/integ/common_args.rs:5-43); pad and gauss are implied by the mesh
 * (one Scratchpad per cell, gathered by Mesh::set_pad, src/mesh/mesh.rs:448-456) and by `ngauss`. */
typedef struct gl_args {
    int32_t ngauss;       /* 0 = Gauss::new(kind) default rule; else Gauss::new_sized(kind.class(), ngauss) (gauss.rs:87,189) */
    int32_t axisymmetric; /* CommonArgs.axisymmetric: integrate over 1 radian, r = sum N.x0 */
    int32_t clear;        /* CommonArgs.clear: 1 = zero-fill the element block first, 0 = accumulate into `out` */
    int32_t packed;       /* batched boundary only (no reference counterpart): 1 = every element MATRIX leaves as its upper triangle,
                             packed column-major -- K(i,j), i <= j, at j(j+1)/2 + i; n(n+1)/2 doubles instead of n^2 -- for callers
                             that assemble into symmetric storage (K is symmetric whenever D / T is major-symmetric).  Halves the
                             device->host bytes of the host-output path.  Needs tight blocks (ii0 = jj0 = nrow = ncol = 0),
                             clear = 1 and a single-kind mesh; the lower triangle is NOT checked against the upper one. */
    double alpha;         /* CommonArgs.alpha (e.g. plane-stress thickness) */
    uint32_t ii0, jj0;    /* CommonArgs.ii0 / jj0: offsets inside each element's block */
    uint32_t nrow, ncol;  /* per-element block dims; 0 = tight (nrow = ii0 + size, ncol = jj0 + size; vectors use nrow only) */
    double ksi[3];        /* GL_OP_JACOBIAN only: the reference point */
    int32_t kind_b;       /* two-pad cases (GL_OP_MAT_04_NSB .. GL_OP_MAT_07_BSN): GeoKind ordinal of the second pad */
    int32_t reserved;
} gl_args;

/* The reference takes the coefficient as a closure evaluated at every Gauss point
 * (FnMut(&mut Tensor4, p, &N, &B), mat_10_bdb.rs:121-124, and siblings).  A device cannot call a host closure, so the
 * batched boundary takes the coefficient as DATA.  One record is:
 *   mat_10_bdb : D as the ns x ns Mandel matrix `dd.matrix()`, column-major, ns = 2*space_ndim
 *   mat_03_btb, mat_08_ntn, vec_04_bt : the Mandel vector `tt.vector()` (ns)
 *   mat_01_nsn, vec_01_ns : the scalar s
 *   mat_02_bvn, mat_09_nvb, vec_02_nv, vec_03_bv : the vector v / w (space_ndim)
 */
typedef enum gl_coef_mode {
    GL_COEF_UNIFORM = 0,    /* one record for every cell and Gauss point */
    GL_COEF_PER_MARKER = 1, /* one record per cell marker: markers[k] -> record k */
    GL_COEF_PER_CELL = 2,   /* record (e - cell_begin) for cell e */
    GL_COEF_PER_GAUSS = 3   /* record (e - cell_begin)*ngauss + p for Gauss point p of cell e */
} gl_coef_mode;

typedef struct gl_coef {
    int32_t mode;           /* gl_coef_mode */
    int32_t location;       /* gl_location of `data` (UNIFORM / PER_MARKER tables must be GL_HOST) */
    const double *data;
    const int32_t *markers; /* PER_MARKER: host array of nrecord marker values */
    uint64_t nrecord;       /* PER_MARKER: number of records */
} gl_coef;

typedef struct gl_out {
    double *ptr;      /* element e occupies ptr[off[e-cell_begin] .. +size), elements in original cell-id order */
    int32_t location; /* gl_location */
    int32_t reserved;
    uint64_t capacity; /* doubles available at ptr */
} gl_out;

/* ---- context ------------------------------------------------------------------------------------ */

/* `stream` is a cudaStream_t (or NULL for the device's default stream) on which all work of this ctx is ordered. */
GL_API int gl_ctx_create(int device, void *stream, gl_ctx **ctx);
GL_API void gl_ctx_destroy(gl_ctx *ctx);
GL_API int gl_ctx_set_stream(gl_ctx *ctx, void *stream);
/* Waits for the ctx stream; returns GL_ERR_ZERO_DET (and records the smallest failing cell id) if any kernel since the
 * last sync met |det J| <= 1e-15. */
GL_API int gl_ctx_sync(gl_ctx *ctx);
GL_API const char *gl_last_error(const gl_ctx *ctx);
GL_API uint64_t gl_last_error_cell(const gl_ctx *ctx);
GL_API const char *gl_status_string(int status);
/* number of kernels launched by this ctx so far (bench.py's gpu_launches) */
GL_API uint64_t gl_ctx_launch_count(const gl_ctx *ctx);
/* which kernel the most recent integration launch of this ctx used: "family<template parameters>", e.g. "hex8_bdb<pat=2>",
 * "syrk_dmma<nn=20,pat=2>", "mass_dmma<nn=10,sd=3,mask=5>", "generic<sd=3,gd=3>".  The dispatch tries the tuned kernels in
 * turn and ends at the generic one; tests and bench.py assert the family so that a tuned kernel that starts declining a
 * configuration cannot go unnoticed.  After gl_integ_fused: every distinct kernel of that call, joined by " + " (one per
 * GeoKind bucket and group of cases).  (No reference counterpart.) */
GL_API const char *gl_ctx_last_kernel(const gl_ctx *ctx);

/* ---- GeoKind / Gauss queries (host only; replace GeoKind::nnode, ndim, class and Gauss::new, new_sized, npoint, coords) ---- */
GL_API int gl_kind_nnode(int kind);
GL_API int gl_kind_geo_ndim(int kind);
GL_API int gl_kind_class(int kind);
GL_API int gl_kind_from_name(const char *name); /* GeoKind::from (enums.rs:217-244); -1 if unknown */
GL_API const char *gl_kind_name(int kind);
/* 1 for every GeoKind (all 20 integrate: N and L come from per-(kind, rule) tables) */
GL_API int gl_kind_supported(int kind);
/* rule lookup: *npoint points, data = npoint x 4 doubles {r,s,t,w} (static storage) */
GL_API int gl_gauss_rule(int kind, int ngauss, int *npoint, const double **data);
/* N (nnode) and L = dN/dksi (nnode x geo_ndim, row-major) at an ARBITRARY point, for all 20 GeoKinds (replaces
 * Scratchpad::calc_interp / calc_deriv, src/shapes/scratchpad.rs): closed family-wise formulas for Lin2/3, Tri3/6, Qua4/8/9,
 * Tet4/10, Hex8/20; the cubic and quartic kinds through their nodal basis (see gl_shapes.cpp) */
GL_API int gl_shape_eval(int kind, const double *ksi, double *nn, double *ll);
/* GeoKind::reference_coords (src/shapes/enums.rs:1049): natural coordinates (geo_ndim values) of node m */
GL_API int gl_kind_reference_coords(int kind, int m, double *ksi);
/* recovery::get_extrap_matrix (src/recovery/get_extrap_matrix.rs:129): the matrix E (nnode x npoint, column-major in `ee`) that
 * extrapolates values at the Gauss points of the rule (ngauss = 0: the kind's default) to the nodes; host only */
GL_API int gl_extrap_matrix(int kind, int ngauss, double *ee);

/* ---- mesh (replaces Mesh/Point/Cell + Mesh::set_pad, src/mesh/mesh.rs:27-54,110-135,448-456) ------ */

/* coords: AoS npoint x ndim (Mesh.points[i].coords flattened); kinds: ncell GeoKind ordinals; markers: ncell (may be
 * NULL = all zero); conn_off: ncell+1 offsets into conn; conn: point ids (usize in the reference).  All HOST pointers.
 * The library transposes to structure-of-arrays, narrows ids to u32 and buckets cells by GeoKind. */
GL_API int gl_mesh_upload(gl_ctx *ctx, int ndim, uint64_t npoint, const double *coords, uint64_t ncell,
                          const uint8_t *kinds, const int32_t *markers, const uint64_t *conn_off, const uint64_t *conn,
                          gl_mesh **mesh);
/* Homogeneous fast path: every cell has `kind`; conn is ncell x nnode row-major, 32-bit point ids. */
GL_API int gl_mesh_upload_homogeneous(gl_ctx *ctx, int ndim, uint64_t npoint, const double *coords, uint64_t ncell,
                                      int kind, const int32_t *markers, const uint32_t *conn, gl_mesh **mesh);
/* Replace the coordinates of an uploaded mesh (e.g. updated-Lagrangian steps); coords as in gl_mesh_upload. */
GL_API int gl_mesh_update_coords(gl_ctx *ctx, gl_mesh *mesh, const double *coords, int location);
/* Device-side lattice generator: the mesh that Block::new(box).set_ndiv(ndiv).subdivide(kind) returns for an axis-aligned
 * box [xmin, xmax] -- cell id = i + nx*(j + ny*k), first-touch point numbering (src/mesh/block.rs:749-812; fixtures
 * src/mesh/samples.rs:1407, 1901) -- built directly in HBM from a closed form of that numbering: no host arrays, no upload.
 * kind: Qua4 / Qua8 / Qua9 (ndiv[0..1], 2D) or Hex8 / Hex20 (ndiv[0..2], 3D); every cell gets `marker`. */
GL_API int gl_mesh_block(gl_ctx *ctx, int kind, const uint32_t *ndiv, const double *xmin, const double *xmax, int32_t marker,
                         gl_mesh **mesh);
/* The same for every Block target -- Qua4, Qua8, Qua9 (2D), Hex8, Hex20 (3D) -- and for a GENERAL block: `control` holds the block's
 * 4 corners or 8 Qua8 nodes (2D), 8 corners or 20 Hex20 nodes (3D), AoS ncontrol x ndim in the reference's local node order; points
 * are placed by the block's own interpolation (Block::subdivide, src/mesh/block.rs:661-812), numbered by first touch in cell order
 * (three passes over the lattice: atomicMin of the touching (cell, node) key, per-cell count + prefix, numbering).
 * gl_mesh_block accepts the quadratic targets as well (box corners). */
GL_API int gl_mesh_block_general(gl_ctx *ctx, int kind, const uint32_t *ndiv, int ncontrol, const double *control, int32_t marker,
                                 gl_mesh **mesh);
/* Copies a device mesh back to HOST arrays (either pointer may be NULL): coords npoint x ndim (AoS, as Mesh.points),
 * conn ncell x nnode 32-bit point ids (single-kind meshes only). */
GL_API int gl_mesh_download(gl_ctx *ctx, const gl_mesh *mesh, double *coords, uint32_t *conn);
GL_API void gl_mesh_free(gl_ctx *ctx, gl_mesh *mesh);
GL_API uint64_t gl_mesh_ncell(const gl_mesh *mesh);
GL_API uint64_t gl_mesh_npoint(const gl_mesh *mesh);
GL_API int gl_mesh_ndim(const gl_mesh *mesh);

/* ---- mesh ingestion: gemlab's MSH text format (Mesh::read / Mesh::from_text, src/mesh/read_text_mesh.rs:312-450,
 * README "MSH file format") and its JSON format (Mesh::read_json, src/mesh/mesh.rs:245-256: the serde image of Mesh).  Host code;
 * grammar and error strings are the reference's ("cannot parse point id", "the id and index of cells must equal each other",
 * "not all marked faces have been found", "cannot parse JSON file", ...).  On failure the functions return GL_ERR_MESH and
 * gl_mesh_text_last_error() (thread-local) holds the string. */
enum { GL_MESH_FORMAT_MSH = 0, GL_MESH_FORMAT_JSON = 1 };
typedef struct gl_host_mesh gl_host_mesh;
GL_API int gl_host_mesh_parse(const char *text, uint64_t len, int format, gl_host_mesh **mesh);
GL_API int gl_host_mesh_read(const char *path, int format, gl_host_mesh **mesh); /* "cannot open file" like Mesh::read */
GL_API void gl_host_mesh_free(gl_host_mesh *mesh);
GL_API const char *gl_mesh_text_last_error(void);
GL_API int gl_host_mesh_ndim(const gl_host_mesh *mesh);
GL_API uint64_t gl_host_mesh_npoint(const gl_host_mesh *mesh);
GL_API uint64_t gl_host_mesh_ncell(const gl_host_mesh *mesh);
GL_API uint64_t gl_host_mesh_nmarked_edge(const gl_host_mesh *mesh);
GL_API uint64_t gl_host_mesh_nmarked_face(const gl_host_mesh *mesh);
/* flat arrays in the layout gl_mesh_upload takes (owned by the handle) */
GL_API const double *gl_host_mesh_coords(const gl_host_mesh *mesh);         /* npoint x ndim */
GL_API const int32_t *gl_host_mesh_point_markers(const gl_host_mesh *mesh); /* npoint */
GL_API const uint8_t *gl_host_mesh_kinds(const gl_host_mesh *mesh);         /* ncell GeoKind ordinals */
GL_API const int32_t *gl_host_mesh_markers(const gl_host_mesh *mesh);       /* ncell */
GL_API const uint64_t *gl_host_mesh_conn_off(const gl_host_mesh *mesh);     /* ncell + 1 */
GL_API const uint64_t *gl_host_mesh_conn(const gl_host_mesh *mesh);
GL_API const int64_t *gl_host_mesh_marked_edges(const gl_host_mesh *mesh);  /* nmarked_edge x (marker, p1, p2) */
GL_API const int64_t *gl_host_mesh_marked_faces(const gl_host_mesh *mesh);  /* nmarked_face x (marker, p1, p2, p3, p4); p4 = -1 for triangles */
GL_API int gl_mesh_upload_host(gl_ctx *ctx, const gl_host_mesh *host, gl_mesh **mesh);
/* parse + upload in one call; the error string is also left in gl_last_error(ctx) */
GL_API int gl_mesh_from_text(gl_ctx *ctx, const char *text, uint64_t len, int format, gl_mesh **mesh);

/* ---- output layout ------------------------------------------------------------------------------- */

/* doubles per element block of `op` for a cell of `kind` (tight, ii0 = jj0 = 0); the two-pad cases and GL_OP_NORMAL depend on
 * args (kind_b, ngauss): use gl_out_total / gl_out_offsets for them (this function returns 0) */
GL_API uint64_t gl_op_out_size(int op, int kind, int ndim);
/* doubles per coefficient record */
GL_API uint64_t gl_op_coef_size(int op, int ndim);
/* total doubles needed for [cell_begin, cell_end) with these args */
GL_API int gl_out_total(const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                        uint64_t *total);
/* off[0 .. cell_end-cell_begin] (HOST): start of every element block, off[n] = total */
GL_API int gl_out_offsets(const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                          uint64_t *off);

/* ---- the batched integ entry points -------------------------------------------------------------- */

GL_API int gl_integ(gl_ctx *ctx, const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end,
                    const gl_args *args, const gl_coef *coef, gl_out *out);

/* replaces integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-233) */
GL_API int gl_mat_10_bdb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
/* replaces integ::mat_01_nsn (src/integ/mat_01_nsn.rs:48-102) */
GL_API int gl_mat_01_nsn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
/* replaces integ::mat_03_btb (src/integ/mat_03_btb.rs:50-131) */
GL_API int gl_mat_03_btb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
/* replace integ::mat_02_bvn (src/integ/mat_02_bvn.rs:48-120), mat_08_ntn (mat_08_ntn.rs:56-135), mat_09_nvb (mat_09_nvb.rs:54-125) */
GL_API int gl_mat_02_bvn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
GL_API int gl_mat_08_ntn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
GL_API int gl_mat_09_nvb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
/* replaces integ::vec_01_ns (src/integ/vec_01_ns.rs:83-130) */
GL_API int gl_vec_01_ns(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                        const gl_coef *coef, gl_out *out);
/* replaces integ::vec_02_nv (src/integ/vec_02_nv.rs:85-140) */
GL_API int gl_vec_02_nv(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                        const gl_coef *coef, gl_out *out);
/* replaces integ::vec_03_bv (src/integ/vec_03_bv.rs:84-141) */
GL_API int gl_vec_03_bv(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                        const gl_coef *coef, gl_out *out);
/* replaces integ::vec_04_bt (src/integ/vec_04_bt.rs:93-170) */
GL_API int gl_vec_04_bt(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                        const gl_coef *coef, gl_out *out);
/* replace integ::mat_04_nsb, mat_05_btn, mat_06_nvn, mat_07_bsn (two pads; args->kind_b names the second one) */
GL_API int gl_mat_04_nsb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
GL_API int gl_mat_05_btn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
GL_API int gl_mat_06_nvn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
GL_API int gl_mat_07_bsn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out);
/* replace integ::mat_01_nsn_bry, vec_01_ns_bry, vec_02_nv_bry (boundary cells: geo_ndim = space_ndim - 1) */
GL_API int gl_mat_01_nsn_bry(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                             const gl_coef *coef, gl_out *out);
GL_API int gl_vec_01_ns_bry(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                            const gl_coef *coef, gl_out *out);
GL_API int gl_vec_02_nv_bry(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                            const gl_coef *coef, gl_out *out);
/* batched Scratchpad::calc_normal_vector at the Gauss points of the rule (scratchpad_calc_normal_vector.rs:111-176) */
GL_API int gl_normal_vectors(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                             gl_out *out);
/* replaces the per-cell Scratchpad::calc_jacobian loop of Mesh::check_jacobian (src/mesh/check.rs:43-60) */
GL_API int gl_jacobian(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                       gl_out *out);

/* Several integration cases over ONE geometry pass (one Jacobian / gradient evaluation per Gauss point for all of them).
 * Semantically identical to calling gl_integ once per item with the same mesh, range and args; the library fuses the
 * items into one kernel when it can -- today: any subset of mat_01_nsn, mat_03_btb, vec_01_ns, vec_03_bv with uniform
 * coefficients and device outputs (e.g. the mass matrix and the body-force vector of BASELINE.json config 4), and the pair
 * mat_10_bdb + vec_04_bt of a Hex8 mesh with the stress given per Gauss point (stiffness and internal-force vector, the two
 * sides of a Newton step), and mat_10_bdb (uniform modulus) + mat_03_btb and / or vec_03_bv over a 2D mesh, whose Qua8 cells
 * take all of them from the geometric tensors of the tensor-core stiffness kernel (stiffness, conductivity matrix and flux
 * vector of BASELINE.json config 5) -- and otherwise runs them one after the other.  The reference has no counterpart: its callers invoke integ::mat_01_nsn and
 * integ::vec_01_ns separately, each recomputing the Jacobian (src/integ/mat_01_nsn.rs:75-76, vec_01_ns.rs:106-107). */
typedef struct gl_fused_item {
    int32_t op; /* gl_op */
    int32_t reserved;
    const gl_coef *coef;
    gl_out *out;
} gl_fused_item;
GL_API int gl_integ_fused(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                          const gl_fused_item *items, int nitems);

/* ---- assembly hand-off: element blocks -> global sparse matrix / vector on the device ----------
 * The reference stops at the element matrix; its callers (pmsim-style assemblers) map the local dof (m, i) -- local row
 * ii = i + m*ndim, src/integ/mat_10_bdb.rs:31-46 -- to a global equation and add every entry to a sparse matrix.
 * `eq` is the point -> equation table those assemblers hold: eq[point*npd + i], npd = gl_op_dofs_per_node(op, ndim)
 * (ndim for mat_10_bdb / vec_02_nv / vec_04_bt, 1 for the scalar-dof cases); a negative entry means "not assembled"
 * (prescribed dof).  Element blocks must be tight (ii0 = jj0 = 0, nrow = ncol = 0); args->clear is ignored. */
GL_API int gl_op_dofs_per_node(int op, int ndim);
/* rows[k], cols[k] (DEVICE, int32) = global row / column of the k-th double of the batched output of `op` over
 * [cell_begin, cell_end): (rows, cols, out) is then the COO triplet list a CooMatrix::put loop would build, without
 * moving the values.  cols may be NULL for vector cases. */
GL_API int gl_triplet_indices(gl_ctx *ctx, const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end,
                              const gl_args *args, const int32_t *eq, int eq_location, int32_t *rows, int32_t *cols);
/* Integrates `op` over the range and ADDS the element blocks into a global object on the device:
 *   matrix cases: CSR values `vals` of the pattern (rowptr int64 [nrow+1], colidx int32 ascending within a row; DEVICE);
 *                 *missing (may be NULL) receives the number of entries the pattern lacks (then GL_ERR_INVALID_ARG);
 *   vector cases: the global vector `vals` (rowptr = colidx = NULL).
 * The element blocks only pass through a bounded scratch slab (<= 256 MB); the sum is order-independent up to rounding. */
GL_API int gl_assemble(gl_ctx *ctx, const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                       const gl_coef *coef, const int32_t *eq, int eq_location, const int64_t *rowptr, const int32_t *colidx,
                       double *vals, uint64_t *missing);

/* Planned assembly (what a Newton loop does: the same mesh, equation table and sparsity pattern at every iteration).
 * gl_assemble_plan_create finds, ONCE, the slot of every element-block entry in the CSR value array (or the global vector) and
 * keeps that position map on the device (4 bytes per entry); gl_assemble_planned then integrates `op` and adds the blocks:
 *   - mat_10_bdb on every solid kind (uniform / per-marker / per-cell / per-Gauss moduli): ONE kernel per GeoKind -- the element
 *     blocks go from shared memory straight into `vals` (RED.ADD.F64); the element-matrix slab is never written to HBM;
 *   - the other cases: element blocks through a bounded scratch slab, then a position-map scatter (no binary search).
 * Asynchronous on the ctx stream (gl_ctx_sync reports zero determinants); `vals` is accumulated into, not cleared.  A plan refers
 * to its mesh: free the plan before the mesh.
 * *missing (may be NULL): entries the pattern lacks (then GL_ERR_INVALID_ARG and no plan).  nnz must be < 2^32 - 1.
 * args->packed = 1 selects SYMMETRIC STORAGE: only entries with global row <= column are assembled (the pattern holds the upper
 * triangle; strictly lower entries are skipped, not reported missing) -- half the adds, for solvers that keep one triangle. */
typedef struct gl_asm_plan gl_asm_plan;
GL_API int gl_assemble_plan_create(gl_ctx *ctx, const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                                   const int32_t *eq, int eq_location, const int64_t *rowptr, const int32_t *colidx, uint64_t *missing,
                                   gl_asm_plan **plan);
GL_API int gl_assemble_planned(gl_ctx *ctx, const gl_asm_plan *plan, const gl_coef *coef, double *vals);
GL_API void gl_assemble_plan_free(gl_ctx *ctx, gl_asm_plan *plan);

/* legacy (gemlab <= 1.x) names used by BASELINE.json's north_star: G was renamed to B in v2 */
GL_API int gl_mat_10_gdg(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_out *out); /* = gl_mat_10_bdb */
GL_API int gl_vec_10_gs(gl_ctx *ctx, const gl_mesh *mesh, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                        const gl_coef *coef, gl_out *out); /* = gl_vec_03_bv */

/* default-initialised CommonArgs (CommonArgs::new: clear = true, axisymmetric = false, alpha = 1, ii0 = jj0 = 0) */
GL_API void gl_args_default(gl_args *args);

/* russell_tensor::LinElasticity::new(E, nu, two_dim, plane_stress).get_modulus().matrix() as a coefficient record
 * (ns x ns column-major; ns = 4 if two_dim else 6) -- convenience for callers without russell_tensor */
GL_API void gl_lin_elasticity(double young, double poisson, int two_dim, int plane_stress, double *dd);

/* ---- device-resident output slabs (no reference counterpart: the reference's output is one host Matrix per cell) ----------
 * The batched output of one call that STAYS on its device: element e of [cell_begin, cell_end) at ptr[off[e - cell_begin]],
 * off from gl_out_offsets (homogeneous meshes: (e - cell_begin) * block).  This is the type the host wrappers hand out
 * (gemlab_cuda::DeviceSlab, gemlab::DeviceSlab) so that a caller can feed gl_triplet_indices / its own kernels / a sparse
 * solver without the element matrices ever crossing PCIe. */
typedef struct gl_slab {
    double *ptr;         /* DEVICE pointer; NULL on input = the library allocates (cudaMalloc) and owns it until gl_slab_free */
    uint64_t capacity;   /* doubles available at ptr (input when ptr != NULL) */
    uint64_t size;       /* out: doubles written */
    uint64_t cell_begin; /* out: the cells this slab holds */
    uint64_t cell_end;
    int32_t device;      /* out: CUDA device ordinal */
    int32_t owned;       /* out: 1 if the library allocated ptr */
} gl_slab;
/* gl_integ into a slab (asynchronous on the ctx stream like every device-output call; gl_ctx_sync reports zero determinants) */
GL_API int gl_integ_slab(gl_ctx *ctx, const gl_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end, const gl_args *args,
                         const gl_coef *coef, gl_slab *slab);
/* copies slab->size doubles to HOST memory and synchronises the ctx */
GL_API int gl_slab_download(gl_ctx *ctx, const gl_slab *slab, double *host);
GL_API void gl_slab_free(gl_slab *slab);

/* ---- single-process multi-GPU driver (no reference counterpart: the reference is single threaded) ------------
 * One gl_ctx + stream + host thread per device; the mesh is replicated, the CELL RANGE is split into contiguous pieces
 *     piece g of [b, e) = [b + floor(g n / G), b + floor((g + 1) n / G)),  n = e - b     (gl_multi_partition)
 * and there is no collective: cells are independent.  Host outputs land concatenated in cell order -- bit-identical to the
 * single-device call; device outputs stay on their devices, one gl_slab per device.  Coefficient records must be on the HOST
 * (PER_CELL / PER_GAUSS records are advanced to each piece); UNIFORM and PER_MARKER tables are replicated. */
typedef struct gl_multi gl_multi;
typedef struct gl_multi_mesh gl_multi_mesh;
GL_API int gl_multi_create(const int *devices, int ndevice, gl_multi **multi);
GL_API void gl_multi_destroy(gl_multi *multi);
GL_API int gl_multi_ndevice(const gl_multi *multi);
GL_API gl_ctx *gl_multi_ctx(gl_multi *multi, int g); /* ctx of the g-th device (gl_last_error_cell, gl_ctx_last_kernel, ...) */
GL_API const char *gl_multi_last_error(const gl_multi *multi);
GL_API void gl_multi_partition(uint64_t cell_begin, uint64_t cell_end, int ndevice, int g, uint64_t *piece_begin, uint64_t *piece_end);
/* arguments as gl_mesh_upload; the replicas are uploaded concurrently */
GL_API int gl_multi_mesh_upload(gl_multi *multi, int ndim, uint64_t npoint, const double *coords, uint64_t ncell, const uint8_t *kinds,
                                const int32_t *markers, const uint64_t *conn_off, const uint64_t *conn, gl_multi_mesh **mesh);
GL_API void gl_multi_mesh_free(gl_multi *multi, gl_multi_mesh *mesh);
GL_API const gl_mesh *gl_multi_mesh_replica(const gl_multi_mesh *mesh, int g);
/* out->location must be GL_HOST (one buffer for the whole range); returns after every device has finished */
GL_API int gl_multi_integ(gl_multi *multi, const gl_multi_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end,
                          const gl_args *args, const gl_coef *coef, gl_out *out);
/* slabs[0 .. ndevice): piece g stays on device g (ptr == NULL: allocated by the library).  Asynchronous: gl_multi_sync waits. */
GL_API int gl_multi_integ_slabs(gl_multi *multi, const gl_multi_mesh *mesh, int op, uint64_t cell_begin, uint64_t cell_end,
                                const gl_args *args, const gl_coef *coef, gl_slab *slabs);
GL_API int gl_multi_sync(gl_multi *multi);

/* ---- micro-benchmarks used to measure the FP64 roofline denominators on the device (bench.py) ---- */
/* runs `iters` dependent-chain DFMA (kind 0) or DMMA m8n8k4 (kind 1) per thread on the whole device and returns the
 * achieved TFLOP/s measured with CUDA events; kind 2 interleaves both */
GL_API int gl_bench_fp64_peak(gl_ctx *ctx, int kind, int iters, double *tflops);
/* writes `nbytes` of the DEVICE buffer `dev` once and returns the achieved write bandwidth in GB/s: kind 0 = 4,608-byte
 * cp.async.bulk shared->global copies with `warps_per_sm` resident warps (the mechanism of the stiffness kernels),
 * kind 1 = coalesced 16-byte stores from registers, kind 2 = cudaMemsetAsync */
GL_API int gl_bench_hbm_write(gl_ctx *ctx, int kind, double *dev, uint64_t nbytes, int warps_per_sm, double *gbs);

#ifdef __cplusplus
}
#endif
#endif /* GEMLAB_CUDA_H */

// gl_internal.hpp -- internal declarations of libgemlab_cuda (not part of the C ABI)
#pragma once

#include <cstddef>
#include <cstdint>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#include "../../include/gemlab_cuda.h"

namespace gl {

// GeoKind ordinals, GeoKind::VALUES order (src/shapes/enums.rs:1467-1493)
enum {
    GL_LIN2 = 0, GL_LIN3, GL_LIN4, GL_LIN5,
    GL_TRI3, GL_TRI6, GL_TRI10, GL_TRI15,
    GL_QUA4, GL_QUA8, GL_QUA9, GL_QUA12, GL_QUA16, GL_QUA17,
    GL_TET4, GL_TET10, GL_TET20,
    GL_HEX8, GL_HEX20, GL_HEX32,
    GL_NKIND
};

constexpr int MAX_NNODE = 32;
constexpr double ZERO_DETERMINANT = 1e-15; // russell_lab::mat_inverse threshold
constexpr unsigned long long NO_ERR_CELL = ~0ull;

// ---- host shapes / rules (gl_shapes.cpp) ---------------------------------------------------------
int kind_nnode(int kind);
int kind_geo_ndim(int kind);
int kind_class(int kind);
const char *kind_name(int kind);
int kind_from_name(const char *name);
bool kind_supported(int kind);
int shape_eval(int kind, const double *ksi, double *nn, double *ll);
int kind_reference_coords(int kind, int m, double *ksi);
int extrap_matrix(int kind, int ngauss, double *ee); // nnode x npoint, column-major
int gauss_rule(int kind, int ngauss, int *npoint, const double **data);
const double *shape_table(int kind, int ng); // N[ng][nnode] | L[ng][nnode][gd], exact; nullptr if not tabulated
const char *status_string(int status);
void lin_elasticity(double young, double poisson, int two_dim, int plane_stress, double *dd);

// ---- device-side parameter block of the integration kernels ---------------------------------------
// Reference-element table of one (kind, rule): doubles laid out as  w[ng] | N[ng][nnode] | L[ng][nnode][gd]
struct RuleTable {
    int kind = 0, ng = 0, nnode = 0, gd = 0;
    std::vector<double> host;
    double *dev = nullptr;
    size_t ndoubles() const { return host.size(); }
};

struct IntegParams {
    // mesh (structure of arrays)
    const double *coords;     // [sd][npoint]
    const double *coords_pm;  // point-major copy [npoint][2] (2D) / [npoint][4] (3D, padded): one 32-byte sector per point for
                              // kernels whose lanes gather whole points of unrelated cells; nullptr if not built
    unsigned long long npoint;
    const uint32_t *conn;     // [nnode][ncell_k]  (bucket of one GeoKind)
    unsigned long long ncell_k;
    const uint32_t *cell_ids; // bucket-local index -> original cell id; nullptr = identity
    unsigned long long lo, hi;      // bucket-local sub-range to integrate
    unsigned long long cell_begin;  // original id of the first cell written at out[0] (chunk start)
    unsigned long long coef_begin;  // original id that coefficient record 0 belongs to (start of the call's range)
    // reference element
    const double *rule;
    int ng, nnode;
    // second pad of the two-pad cases: Nb[ng][nnode_b] | Lb[ng][nnode_b][gd] at the same Gauss points; its nodes are the cell's first nnode_b
    const double *rule_b;
    int nnode_b;
    // CommonArgs
    int op, axisym, clear;
    double alpha;
    // coefficient records
    const double *coef;       // device pointer (PER_MARKER / PER_CELL / PER_GAUSS) or nullptr (UNIFORM, in coef_uniform)
    unsigned long long coef_cell_stride, coef_gp_stride; // in doubles
    const uint32_t *coef_index; // PER_MARKER: bucket-local cell -> record index; nullptr otherwise
    int coef_len;
    int coef_pat;             // mat_10_bdb with record data: structure common to ALL records when the host could inspect them
                              // (2 = symmetric with uncoupled normal / diagonal shear blocks, 1 = major-symmetric, 0 = unknown / general)
    const int *coef_pat_dev;  // device-resident records: their common structure, found by a probe kernel earlier on the stream;
                              // the launcher then enqueues every pattern variant and all but the matching one exit at once
    int *pg_flag;             // PER_GAUSS moduli: device flag of the in-kernel symmetry check (see gl_kernels_bdb.cu), owned by the ctx
    double coef_uniform[36];
    // output
    double *out;
    const unsigned long long *out_off; // bucket-local cell -> absolute block offset (mixed meshes); nullptr = homogeneous
    unsigned long long out_base;       // subtracted from out_off (offset of cell_begin)
    unsigned int nrow, ncol, ii0, jj0; // element block (ncol = 1 for vectors)
    unsigned int size_r, size_c;       // extent actually integrated (ndof / nnode, ...)
    // fused internal-force vector (Hex8 kernel only): stress records sigma[(e - coef_begin)*ng + p][ns] (Mandel vectors, DEVICE)
    // and the output d[(e - cell_begin)*ndof ..]; nullptr = not requested
    const double *sigma;
    double *out_d;
    // fused assembly (gl_assemble_planned): instead of writing the element blocks, add entry (row, col) of the block of the cell with
    // original id e to scatter_vals[scatter_pos[off(e) + row * ncol + col]] (0xFFFFFFFF = not assembled); off(e) as for `out`
    const uint32_t *scatter_pos;
    double *scatter_vals;
    // failure reporting: smallest original id of a cell with |det J| <= 1e-15
    unsigned long long *err_cell;
    // kernels whose warps draw cells from a queue (gl_kernels_syrk.cu): a counter owned by the ctx, zeroed by the launcher
    unsigned long long *work_counter;
    // tile
    int cells_per_cta;
};

// Tuning / testing knobs are read from the environment ONLY in a library built with -DGL_TUNING (make TUNING=1 ->
// libgemlab_cuda_tuning.so); the product library ignores them.
#ifdef GL_TUNING
const char *tuning_env(const char *name);
#else
inline const char *tuning_env(const char *) { return nullptr; }
#endif

// every launcher records the kernel it enqueued ("family<template parameters>"); gl_api.cu copies it into the ctx after a launch
// (gl_ctx_last_kernel) so that tests and bench.py can assert WHICH kernel ran, not only that one did
void note_kernel(const char *fmt, ...);
const char *noted_kernel();

// launchers (gl_kernels_generic.cu, gl_kernels_hex8.cu, ...); return the number of kernels launched or < 0 on CUDA error
int launch_generic(const IntegParams &p, int sd, int gd, void *stream);
size_t generic_smem_bytes(int sd, int gd, int nnode, int ng, int op, int cells_per_cta, int nnode_b = 0);
int generic_pick_cells_per_cta(int sd, int gd, int nnode, int ng, int op, int nnode_b = 0);

// tuned kernels; return 0 if the configuration is not covered (caller falls back to the generic kernel)
int launch_hex8_bdb(const IntegParams &p, void *stream);
int launch_hex8_gauss(const IntegParams &p, void *stream); // Hex8, one modulus per Gauss point (gl_kernels_hex8pg.cu)
int launch_bdb(const IntegParams &p, int sd, void *stream);
bool bdb_covers(int nnode, int sd, int ng, bool axisym, bool record_moduli, bool per_gauss); // would launch_bdb take this stiffness configuration?
int launch_bdb_tile(const IntegParams &p, int sd, void *stream); // register-tiled variant for Hex20 / Tet10 (gl_kernels_tile.cu)
int launch_bdb_syrk(const IntegParams &p, int sd, void *stream); // FP64 tensor-core (DMMA) symmetric rank-ng update for Hex20 / Tet10 (gl_kernels_syrk.cu)

// fused scalar-dof cases (gl_kernels_scalar.cu): any subset of mat_01_nsn (bit 0), mat_03_btb (bit 1), vec_01_ns (bit 2),
// vec_03_bv (bit 3) from one geometry pass; uniform coefficients
struct ScalarFused {
    int mask;
    double s_mat, s_vec; // mat_01_nsn / vec_01_ns scalars
    double t[6];         // mat_03_btb Mandel vector
    double w[3];         // vec_03_bv vector
    double *out_m01, *out_m03, *out_v01, *out_v03;
    const unsigned long long *moff; // mixed meshes: bucket-local cell -> absolute offset of its nnode x nnode block
    unsigned long long mbase;
    const unsigned long long *voff; // ... of its nnode vector
    unsigned long long vbase;
};
int launch_scalar_fused(const IntegParams &p, const ScalarFused &f, int sd, void *stream);
// FP64 tensor-core (DMMA) geometric tensors for Qua8 stiffness, plane / axisymmetric (gl_kernels_qua8.cu); `extra` (mask of mat_03_btb = 2,
// vec_03_bv = 8) asks for the conductivity matrix and / or the flux vector from the same pass
int launch_bdb_qua8(const IntegParams &p, int sd, void *stream, const ScalarFused *extra = nullptr);
int launch_bdb_small2d(const IntegParams &p, int sd, void *stream); // thread-per-cell variant for Tri3 / Qua4 (gl_kernels_small2d.cu)
// mat_01_nsn and/or vec_01_ns as a (cells x ng).(ng x entries) product on the FP64 tensor cores (gl_kernels_mass.cu); 0 if not covered
// mat_03_btb and/or vec_03_bv, one thread per cell (gl_kernels_scalar_tpc.cu); 0 if not covered
int launch_scalar_tpc(const IntegParams &p, const ScalarFused &f, int sd, void *stream);
// vec_02_nv / vec_04_bt, one thread per cell, every coefficient mode (gl_kernels_vec_tpc.cu); 0 if not covered
int launch_vec_tpc(const IntegParams &p, int sd, void *stream);
int launch_mass_dmma(const IntegParams &p, const ScalarFused &f, int sd, void *stream);

// micro-benchmarks
int bench_fp64(int kind, int iters, void *stream, double *tflops);
int bench_hbm_write(int kind, double *dev, unsigned long long nbytes, int warps_per_sm, void *stream, double *gbs);

} // namespace gl

// cache key of a mixed-mesh block layout: (size class, ii0, jj0)
using BlockKey = std::tuple<int, uint32_t, uint32_t>;

// ---- opaque handles ---------------------------------------------------------------------------------
struct gl_mesh_bucket {
    int kind = 0, nnode = 0;
    uint64_t ncell = 0;
    std::vector<uint32_t> h_cell_ids; // empty => identity (homogeneous mesh)
    uint32_t *d_cell_ids = nullptr;
    uint32_t *d_conn = nullptr;       // [nnode][ncell]
    uint32_t *d_marker_idx = nullptr; // [ncell] index into mesh->marker_values; nullptr if the mesh has one marker
    std::map<BlockKey, unsigned long long *> d_out_off; // size-class -> absolute block offsets (mixed meshes, lazy)
};

struct gl_mesh {
    int ndim = 0;
    uint64_t npoint = 0, ncell = 0;
    double *d_coords = nullptr; // [ndim][npoint]
    double *d_coords_pm = nullptr; // lazily built point-major copy (see IntegParams::coords_pm)
    bool homogeneous = true;
    std::vector<gl_mesh_bucket> buckets;
    std::vector<uint8_t> h_kinds;        // per cell (mixed meshes only)
    std::vector<int32_t> marker_values;  // sorted unique markers
    std::map<BlockKey, std::vector<uint64_t>> h_prefix; // size-class -> exclusive prefix of block sizes (mixed meshes, lazy)
    int device = 0;
};

struct gl_ctx {
    int device = 0;
    void *stream = nullptr;
    int last_status = 0;
    std::string last_error;
    uint64_t last_error_cell = gl::NO_ERR_CELL;
    uint64_t launches = 0;
    std::string last_kernel; // kernel family of the most recent integration launch (gl_ctx_last_kernel)
    bool collect_kernels = false; // inside gl_integ_fused: last_kernel accumulates every distinct kernel of the call
    unsigned long long *d_err_cell = nullptr;
    unsigned long long *d_work_counter = nullptr; // see IntegParams::work_counter
    int *d_coef_pat = nullptr; // result of the modulus-structure probe (device-resident PER_CELL moduli)
    unsigned long long *h_err_cell = nullptr; // pinned
    std::map<std::pair<int, int>, std::unique_ptr<gl::RuleTable>> rules; // (kind, ngauss) -> table
    std::map<std::tuple<int, int, int>, std::unique_ptr<gl::RuleTable>> rules_b; // (kind, kind_b, ngauss) -> table of the second pad
    // scratch for host<->device staging
    double *d_stage[2] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    double *d_coef = nullptr;
    size_t coef_bytes = 0;
    double *d_full = nullptr; // scratch slab of full element blocks (gl_args.packed)
    size_t full_bytes = 0;
    double *d_rule_tmp = nullptr; // for GL_OP_JACOBIAN tables
    // freed mesh buffers kept for reuse: an upload -> integrate -> free cycle (one per step of a host-driven loop) then
    // performs no cudaMalloc / cudaFree, whose implicit device synchronisations dominate small uploads
    std::vector<std::pair<void *, size_t>> pool;
    size_t pool_bytes = 0;
    void *copy_stream = nullptr;
    double *h_coef = nullptr; // pinned staging of PER_MARKER tables (keeps those calls asynchronous)
    size_t h_coef_bytes = 0;
    void *ev_coef = nullptr;  // completion of the last copy out of h_coef
    // host-resident PER_CELL / PER_GAUSS records go through two pinned bounce buffers (upload_records, gl_api.cu)
    double *h_rec[2] = {nullptr, nullptr};
    void *ev_rec[2] = {nullptr, nullptr};
    void *ev[4] = {nullptr, nullptr, nullptr, nullptr};
};

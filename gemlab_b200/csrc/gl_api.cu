// gl_api.cu -- C ABI of libgemlab_cuda (see include/gemlab_cuda.h): context, device mesh, batched integ entry points.
//
// Host-side responsibilities (everything the reference does once per CELL in Rust is done here once per CALL):
//   * Mesh -> structure of arrays on the device, bucketed by GeoKind (replaces Mesh::set_pad's AoS pointer chase,
//     src/mesh/mesh.rs:448-456);
//   * reference-element tables N, L, w per (kind, rule) (replaces the fn_interp / fn_deriv calls per Gauss point);
//   * CommonArgs validation with the reference's error strings (src/integ/mat_10_bdb.rs:126-136 and siblings);
//   * coefficient closures turned into data records (UNIFORM / PER_MARKER / PER_CELL / PER_GAUSS);
//   * output placement: element blocks in original cell-id order, column-major, tight or (nrow, ncol, ii0, jj0);
//   * host-buffer calls are pipelined in chunks: kernel(chunk k+1) overlaps the D2H copy of chunk k.
// There is deliberately NO CPU fallback: without a usable sm_100 device every entry point fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <thread>

#include "gl_internal.hpp"

using namespace gl;

namespace gl {
static thread_local char g_noted_kernel[128] = "";
void note_kernel(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    std::vsnprintf(g_noted_kernel, sizeof(g_noted_kernel), fmt, ap);
    va_end(ap);
}
const char *noted_kernel() { return g_noted_kernel; }
#ifdef GL_TUNING
const char *tuning_env(const char *name) { return std::getenv(name); }
#endif
} // namespace gl

namespace {

int fail(gl_ctx *ctx, int status, const char *detail = nullptr) {
    if (ctx) {
        ctx->last_status = status;
        ctx->last_error = status_string(status);
        if (detail && status >= GL_ERR_INVALID_ARG) {
            ctx->last_error += ": ";
            ctx->last_error += detail;
        }
    }
    return status;
}

int cuda_fail(gl_ctx *ctx, cudaError_t e, const char *what) {
    char buf[256];
    std::snprintf(buf, sizeof(buf), "%s: %s", what, cudaGetErrorString(e));
    return fail(ctx, GL_ERR_CUDA, buf);
}

#define CUDA_TRY(ctx, call)                                   \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call); \
    } while (0)

// size classes of the element blocks
enum { SZ_NDOF2 = 0, SZ_NNODE2 = 1, SZ_NNODE = 2, SZ_NDOF = 3, SZ_ONE = 4, SZ_NB_NDOF = 5, SZ_NDOF_NB = 6, SZ_NORMAL = 7 };

int op_size_class(int op) {
    switch (op) {
    case GL_OP_MAT_10_BDB: case GL_OP_MAT_08_NTN: case GL_OP_MAT_09_NVB: return SZ_NDOF2;
    case GL_OP_MAT_01_NSN: case GL_OP_MAT_03_BTB: case GL_OP_MAT_02_BVN: case GL_OP_MAT_01_NSN_BRY: return SZ_NNODE2;
    case GL_OP_VEC_01_NS: case GL_OP_VEC_03_BV: case GL_OP_VEC_01_NS_BRY: return SZ_NNODE;
    case GL_OP_VEC_02_NV: case GL_OP_VEC_04_BT: case GL_OP_VEC_02_NV_BRY: return SZ_NDOF;
    case GL_OP_JACOBIAN: return SZ_ONE;
    case GL_OP_MAT_04_NSB: case GL_OP_MAT_05_BTN: return SZ_NB_NDOF;
    case GL_OP_MAT_06_NVN: case GL_OP_MAT_07_BSN: return SZ_NDOF_NB;
    case GL_OP_NORMAL: return SZ_NORMAL;
    default: return -1;
    }
}

bool op_is_matrix(int op) {
    return op == GL_OP_MAT_10_BDB || op == GL_OP_MAT_01_NSN || op == GL_OP_MAT_03_BTB || op == GL_OP_MAT_02_BVN ||
           op == GL_OP_MAT_08_NTN || op == GL_OP_MAT_09_NVB || op == GL_OP_MAT_01_NSN_BRY ||
           (op >= GL_OP_MAT_04_NSB && op <= GL_OP_MAT_07_BSN);
}

bool op_two_pads(int op) { return op >= GL_OP_MAT_04_NSB && op <= GL_OP_MAT_07_BSN; }
bool op_boundary(int op) { return op >= GL_OP_MAT_01_NSN_BRY && op <= GL_OP_NORMAL; }

bool op_uses_gradient(int op) {
    // the reference calls Scratchpad::calc_gradient for these (mat_01_nsn too, src/integ/mat_01_nsn.rs:76)
    return op == GL_OP_MAT_10_BDB || op == GL_OP_MAT_01_NSN || op == GL_OP_MAT_03_BTB || op == GL_OP_VEC_03_BV ||
           op == GL_OP_VEC_04_BT || op == GL_OP_MAT_02_BVN || op == GL_OP_MAT_08_NTN || op == GL_OP_MAT_09_NVB || op_two_pads(op);
}

// rows / cols actually integrated for a cell of `kind`; the two-pad cases read args->kind_b, GL_OP_NORMAL the rule size
void op_extent(int op, int kind, int ndim, const gl_args *a, unsigned &size_r, unsigned &size_c) {
    const unsigned nnode = (unsigned)kind_nnode(kind), ndof = nnode * (unsigned)ndim;
    const unsigned nb = (a && op_two_pads(op) && a->kind_b >= 0 && a->kind_b < GL_NKIND) ? (unsigned)kind_nnode(a->kind_b) : 0u;
    switch (op_size_class(op)) {
    case SZ_NDOF2: size_r = ndof; size_c = ndof; break;
    case SZ_NNODE2: size_r = nnode; size_c = nnode; break;
    case SZ_NNODE: size_r = nnode; size_c = 1; break;
    case SZ_NDOF: size_r = ndof; size_c = 1; break;
    case SZ_NB_NDOF: size_r = nb; size_c = ndof; break;
    case SZ_NDOF_NB: size_r = ndof; size_c = nb; break;
    case SZ_NORMAL: {
        int ng = 0;
        const double *data = nullptr;
        if (!a || gauss_rule(kind, a->ngauss, &ng, &data) != GL_OK) ng = 0;
        size_r = (unsigned)ng * (unsigned)(ndim + 1);
        size_c = 1;
    } break;
    default: size_r = 1; size_c = 1; break;
    }
}

struct BlockDims {
    unsigned nrow, ncol, size_r, size_c;
    uint64_t block() const { return (uint64_t)nrow * ncol; }
};

// resolves the element block of (op, kind) under args; returns the reference's dimension errors
int block_dims(int op, int kind, int ndim, const gl_args *a, BlockDims &bd) {
    op_extent(op, kind, ndim, a, bd.size_r, bd.size_c);
    const bool mat = op_is_matrix(op);
    const unsigned ii0 = a->ii0, jj0 = mat ? a->jj0 : 0;
    bd.nrow = a->nrow ? a->nrow : ii0 + bd.size_r;
    bd.ncol = mat ? (a->ncol ? a->ncol : jj0 + bd.size_c) : 1;
    if (bd.nrow < ii0 + bd.size_r) {
        switch (op) {
        case GL_OP_MAT_10_BDB: case GL_OP_MAT_08_NTN: case GL_OP_MAT_09_NVB: return GL_ERR_NROW_KK_NDOF;
        case GL_OP_MAT_01_NSN: case GL_OP_MAT_03_BTB: case GL_OP_MAT_02_BVN: case GL_OP_MAT_01_NSN_BRY: return GL_ERR_NROW_KK_NNODE;
        case GL_OP_MAT_04_NSB: case GL_OP_MAT_05_BTN: return GL_ERR_NROW_KK_NNODE_B;
        case GL_OP_MAT_06_NVN: case GL_OP_MAT_07_BSN: return GL_ERR_NROW_KK_NDOF_PAD;
        case GL_OP_VEC_01_NS: case GL_OP_VEC_01_NS_BRY: return GL_ERR_A_LEN;
        case GL_OP_VEC_02_NV: case GL_OP_VEC_02_NV_BRY: return GL_ERR_B_LEN;
        case GL_OP_VEC_03_BV: return GL_ERR_C_LEN;
        case GL_OP_VEC_04_BT: return GL_ERR_D_LEN;
        default: return GL_ERR_INVALID_ARG;
        }
    }
    if (mat && bd.ncol < jj0 + bd.size_c) {
        switch (op_size_class(op)) {
        case SZ_NDOF2: return GL_ERR_NCOL_KK_NDOF;
        case SZ_NB_NDOF: return GL_ERR_NCOL_KK_NDOF_PAD;
        case SZ_NDOF_NB: return GL_ERR_NCOL_KK_NNODE_B;
        default: return GL_ERR_NCOL_KK_NNODE;
        }
    }
    return GL_OK;
}

uint64_t coef_len(int op, int ndim) {
    const uint64_t ns = 2 * (uint64_t)ndim;
    switch (op) {
    case GL_OP_MAT_10_BDB: return ns * ns;
    case GL_OP_MAT_01_NSN: case GL_OP_VEC_01_NS: return 1;
    case GL_OP_MAT_03_BTB: case GL_OP_VEC_04_BT: case GL_OP_MAT_08_NTN: return ns;
    case GL_OP_VEC_02_NV: case GL_OP_VEC_03_BV: case GL_OP_MAT_02_BVN: case GL_OP_MAT_09_NVB: return (uint64_t)ndim;
    case GL_OP_MAT_04_NSB: case GL_OP_MAT_07_BSN: return 1;
    case GL_OP_MAT_05_BTN: return ns;
    case GL_OP_MAT_06_NVN: return (uint64_t)ndim;
    case GL_OP_MAT_01_NSN_BRY: case GL_OP_VEC_01_NS_BRY: case GL_OP_VEC_02_NV_BRY: return (uint64_t)ndim + 1;
    default: return 0;
    }
}

// bucket-local sub-range [lo, hi) of original ids [b, e)
void bucket_range(const gl_mesh *mesh, const gl_mesh_bucket &bk, uint64_t b, uint64_t e, uint64_t &lo, uint64_t &hi) {
    if (bk.h_cell_ids.empty()) { // identity (homogeneous mesh)
        lo = std::min<uint64_t>(b, bk.ncell);
        hi = std::min<uint64_t>(e, bk.ncell);
        (void)mesh;
        return;
    }
    lo = (uint64_t)(std::lower_bound(bk.h_cell_ids.begin(), bk.h_cell_ids.end(), (uint32_t)b) - bk.h_cell_ids.begin());
    hi = (uint64_t)(std::lower_bound(bk.h_cell_ids.begin(), bk.h_cell_ids.end(), (uint32_t)std::min<uint64_t>(e, 0xffffffffull)) - bk.h_cell_ids.begin());
    if (e > 0xffffffffull) hi = bk.ncell;
}

// exclusive prefix of block sizes over the whole (mixed) mesh, cached by the args that shape the block
const std::vector<uint64_t> *mesh_prefix(gl_mesh *mesh, int op, const gl_args *a, BlockKey *key_out, int *status) {
    *status = GL_OK;
    const int cls = op_size_class(op);
    // key: size class + offsets (explicit nrow/ncol make every block equal; handled by the caller)
    // the full offsets (a truncated key would alias tight-block layouts with large ii0 / jj0); the block sizes of the two-pad cases
    // and of GL_OP_NORMAL also depend on kind_b / the rule
    const BlockKey key(cls | ((op_two_pads(op) ? (a->kind_b & 0xff) : 0) << 8) | ((cls == SZ_NORMAL ? (a->ngauss & 0xff) : 0) << 16), a->ii0, a->jj0);
    *key_out = key;
    auto it = mesh->h_prefix.find(key);
    if (it != mesh->h_prefix.end()) return &it->second;
    uint64_t per_kind[GL_NKIND];
    for (int k = 0; k < GL_NKIND; k++) {
        BlockDims bd;
        gl_args tight = *a;
        tight.nrow = 0;
        tight.ncol = 0;
        block_dims(op, k, mesh->ndim, &tight, bd);
        per_kind[k] = bd.block();
    }
    std::vector<uint64_t> pre(mesh->ncell + 1);
    uint64_t acc = 0;
    for (uint64_t e = 0; e < mesh->ncell; e++) {
        pre[e] = acc;
        acc += per_kind[mesh->h_kinds[e]];
    }
    pre[mesh->ncell] = acc;
    auto ins = mesh->h_prefix.emplace(key, std::move(pre));
    return &ins.first->second;
}

RuleTable *get_rule(gl_ctx *ctx, int kind, int ngauss, int *status) {
    auto key = std::make_pair(kind, ngauss);
    auto it = ctx->rules.find(key);
    if (it != ctx->rules.end()) {
        *status = GL_OK;
        return it->second.get();
    }
    int ng = 0;
    const double *data = nullptr;
    *status = gauss_rule(kind, ngauss, &ng, &data);
    if (*status) return nullptr;
    if (!kind_supported(kind)) {
        *status = GL_ERR_KIND_UNSUPPORTED;
        return nullptr;
    }
    auto rt = std::make_unique<RuleTable>();
    rt->kind = kind;
    rt->ng = ng;
    rt->nnode = kind_nnode(kind);
    rt->gd = kind_geo_ndim(kind);
    const int nn = rt->nnode, gd = rt->gd;
    rt->host.assign((size_t)ng * (1 + nn + nn * gd), 0.0);
    double *w = rt->host.data(), *N = w + ng, *L = N + (size_t)ng * nn;
    const double *exact = shape_table(kind, ng); // bit-identical to the reference's own expressions (shape_tables.inc)
    for (int p = 0; p < ng; p++) {
        w[p] = data[p * 4 + 3];
        if (!exact) shape_eval(kind, data + p * 4, N + (size_t)p * nn, L + (size_t)p * nn * gd);
    }
    if (exact) std::memcpy(N, exact, sizeof(double) * (size_t)ng * nn * (1 + gd));
    if (cudaMalloc(&rt->dev, rt->host.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(rt->dev, rt->host.data(), rt->host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        *status = GL_ERR_CUDA;
        return nullptr;
    }
    RuleTable *raw = rt.get();
    ctx->rules[key] = std::move(rt);
    return raw;
}

// table of the SECOND pad of a two-pad case: Nb[ng][nnode_b] | Lb[ng][nnode_b][gd] of kind_b at the Gauss points of the rule of
// the cell's own kind (pad_b.fn_interp / pad_b.calc_gradient at args.gauss, src/integ/mat_04_nsb.rs:98, mat_05_btn.rs:106)
RuleTable *get_rule_b(gl_ctx *ctx, int kind, int kind_b, int ngauss, int *status) {
    auto key = std::make_tuple(kind, kind_b, ngauss);
    auto it = ctx->rules_b.find(key);
    if (it != ctx->rules_b.end()) {
        *status = GL_OK;
        return it->second.get();
    }
    int ng = 0;
    const double *data = nullptr;
    *status = gauss_rule(kind, ngauss, &ng, &data);
    if (*status) return nullptr;
    auto rt = std::make_unique<RuleTable>();
    rt->kind = kind_b;
    rt->ng = ng;
    rt->nnode = kind_nnode(kind_b);
    rt->gd = kind_geo_ndim(kind_b);
    const int nn = rt->nnode, gd = rt->gd;
    rt->host.assign((size_t)ng * (nn + nn * gd), 0.0);
    double *N = rt->host.data(), *L = N + (size_t)ng * nn;
    for (int p = 0; p < ng; p++) {
        *status = shape_eval(kind_b, data + p * 4, N + (size_t)p * nn, L + (size_t)p * nn * gd);
        if (*status) return nullptr;
    }
    if (cudaMalloc(&rt->dev, rt->host.size() * sizeof(double)) != cudaSuccess ||
        cudaMemcpy(rt->dev, rt->host.data(), rt->host.size() * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
        *status = GL_ERR_CUDA;
        return nullptr;
    }
    RuleTable *raw = rt.get();
    ctx->rules_b[key] = std::move(rt);
    return raw;
}

int ensure_buffer(gl_ctx *ctx, double **buf, size_t *have, size_t need) {
    if (*have >= need) return GL_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr;
    *have = 0;
    CUDA_TRY(ctx, cudaMalloc(buf, need));
    *have = need;
    return GL_OK;
}

// ---- caching allocator of the mesh buffers (see gl_ctx::pool) ----------------------------------------------------------
constexpr size_t POOL_LIMIT = (size_t)6 << 30;

cudaError_t pool_alloc(gl_ctx *ctx, void **ptr, size_t bytes) {
    if (bytes == 0) bytes = 16;
    size_t best = ctx->pool.size();
    for (size_t i = 0; i < ctx->pool.size(); i++)
        if (ctx->pool[i].second >= bytes && ctx->pool[i].second <= bytes + bytes / 4 + 4096 &&
            (best == ctx->pool.size() || ctx->pool[i].second < ctx->pool[best].second))
            best = i;
    if (best != ctx->pool.size()) {
        *ptr = ctx->pool[best].first;
        ctx->pool_bytes -= ctx->pool[best].second;
        ctx->pool.erase(ctx->pool.begin() + (long)best);
        return cudaSuccess;
    }
    return cudaMalloc(ptr, bytes);
}

// `bytes` must be the size passed to pool_alloc; the caller guarantees that all work using the buffer has completed or is
// ordered on the ctx stream before any reuse (every consumer runs on that stream)
void pool_free(gl_ctx *ctx, void *ptr, size_t bytes) {
    if (!ptr) return;
    if (bytes == 0) bytes = 16;
    if (ctx && ctx->pool_bytes + bytes <= POOL_LIMIT && ctx->pool.size() < 64) {
        ctx->pool.emplace_back(ptr, bytes);
        ctx->pool_bytes += bytes;
    } else {
        cudaFree(ptr);
    }
}

template <typename T>
cudaError_t pool_alloc_t(gl_ctx *ctx, T **ptr, size_t bytes) { return pool_alloc(ctx, reinterpret_cast<void **>(ptr), bytes); }

struct CoefResolved {
    int pat = 0; // see IntegParams::coef_pat
    const int *pat_dev = nullptr; // see IntegParams::coef_pat_dev
    const double *dev = nullptr; // nullptr => uniform record in `uniform`
    uint64_t cell_stride = 0, gp_stride = 0;
    bool per_marker = false;
    double uniform[36] = {0};
    int len = 0;
};

} // namespace

// =====================================================================================================
// context
// =====================================================================================================

extern "C" {

int gl_ctx_create(int device, void *stream, gl_ctx **out) {
    if (!out) return GL_ERR_INVALID_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return GL_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GL_ERR_NO_DEVICE;
    if (prop.major != 10) return GL_ERR_NO_DEVICE; // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return GL_ERR_CUDA;
    gl_ctx *ctx = new gl_ctx();
    ctx->device = device;
    ctx->stream = stream;
    if (cudaMalloc(&ctx->d_err_cell, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_work_counter, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_coef_pat, sizeof(int)) != cudaSuccess ||
        cudaMallocHost(&ctx->h_err_cell, sizeof(unsigned long long)) != cudaSuccess ||
        cudaMalloc(&ctx->d_rule_tmp, sizeof(double) * GL_NKIND * (1 + MAX_NNODE + 3 * MAX_NNODE)) != cudaSuccess) {
        delete ctx;
        return GL_ERR_CUDA;
    }
    *ctx->h_err_cell = NO_ERR_CELL;
    cudaMemcpy(ctx->d_err_cell, ctx->h_err_cell, sizeof(unsigned long long), cudaMemcpyHostToDevice);
    cudaStream_t cs;
    if (cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) {
        delete ctx;
        return GL_ERR_CUDA;
    }
    ctx->copy_stream = cs;
    for (int i = 0; i < 4; i++) {
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        ctx->ev[i] = ev;
    }
    {
        cudaEvent_t ev;
        cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
        ctx->ev_coef = ev;
    }
    *out = ctx;
    return GL_OK;
}

void gl_ctx_destroy(gl_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize((cudaStream_t)ctx->stream);
    for (auto &kv : ctx->rules)
        if (kv.second->dev) cudaFree(kv.second->dev);
    for (auto &kv : ctx->rules_b)
        if (kv.second->dev) cudaFree(kv.second->dev);
    if (ctx->d_err_cell) cudaFree(ctx->d_err_cell);
    if (ctx->d_work_counter) cudaFree(ctx->d_work_counter);
    if (ctx->d_coef_pat) cudaFree(ctx->d_coef_pat);
    if (ctx->h_err_cell) cudaFreeHost(ctx->h_err_cell);
    if (ctx->d_rule_tmp) cudaFree(ctx->d_rule_tmp);
    for (int i = 0; i < 2; i++)
        if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
    if (ctx->d_coef) cudaFree(ctx->d_coef);
    if (ctx->d_full) cudaFree(ctx->d_full);
    for (auto &pb : ctx->pool) cudaFree(pb.first);
    if (ctx->copy_stream) cudaStreamDestroy((cudaStream_t)ctx->copy_stream);
    for (int i = 0; i < 4; i++)
        if (ctx->ev[i]) cudaEventDestroy((cudaEvent_t)ctx->ev[i]);
    if (ctx->ev_coef) cudaEventDestroy((cudaEvent_t)ctx->ev_coef);
    if (ctx->h_coef) cudaFreeHost(ctx->h_coef);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_rec[i]) cudaEventDestroy((cudaEvent_t)ctx->ev_rec[i]);
        if (ctx->h_rec[i]) cudaFreeHost(ctx->h_rec[i]);
    }
    delete ctx;
}

int gl_ctx_set_stream(gl_ctx *ctx, void *stream) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    ctx->stream = stream;
    return GL_OK;
}

int gl_ctx_sync(gl_ctx *ctx) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_err_cell, ctx->d_err_cell, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                  (cudaStream_t)ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize((cudaStream_t)ctx->stream));
    if (*ctx->h_err_cell != NO_ERR_CELL) {
        ctx->last_error_cell = *ctx->h_err_cell;
        *ctx->h_err_cell = NO_ERR_CELL;
        CUDA_TRY(ctx, cudaMemcpy(ctx->d_err_cell, ctx->h_err_cell, sizeof(unsigned long long), cudaMemcpyHostToDevice));
        return fail(ctx, GL_ERR_ZERO_DET);
    }
    return GL_OK;
}

const char *gl_last_error(const gl_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }
uint64_t gl_last_error_cell(const gl_ctx *ctx) { return ctx ? ctx->last_error_cell : NO_ERR_CELL; }
const char *gl_status_string(int status) { return status_string(status); }
uint64_t gl_ctx_launch_count(const gl_ctx *ctx) { return ctx ? ctx->launches : 0; }
const char *gl_ctx_last_kernel(const gl_ctx *ctx) { return ctx ? ctx->last_kernel.c_str() : ""; }

int gl_kind_nnode(int kind) { return kind_nnode(kind); }
int gl_kind_geo_ndim(int kind) { return kind_geo_ndim(kind); }
int gl_kind_class(int kind) { return kind_class(kind); }
int gl_kind_from_name(const char *name) { return name ? kind_from_name(name) : -1; }
const char *gl_kind_name(int kind) { return kind_name(kind); }
int gl_kind_supported(int kind) { return kind_supported(kind) ? 1 : 0; }
int gl_gauss_rule(int kind, int ngauss, int *npoint, const double **data) {
    if (!npoint || !data) return GL_ERR_INVALID_ARG;
    return gauss_rule(kind, ngauss, npoint, data);
}
int gl_shape_eval(int kind, const double *ksi, double *nn, double *ll) {
    if (!ksi || !nn || !ll) return GL_ERR_INVALID_ARG;
    return shape_eval(kind, ksi, nn, ll);
}

int gl_kind_reference_coords(int kind, int m, double *ksi) {
    if (!ksi) return GL_ERR_INVALID_ARG;
    return kind_reference_coords(kind, m, ksi);
}
int gl_extrap_matrix(int kind, int ngauss, double *ee) {
    if (!ee) return GL_ERR_INVALID_ARG;
    return extrap_matrix(kind, ngauss, ee);
}

void gl_args_default(gl_args *a) {
    if (!a) return;
    std::memset(a, 0, sizeof(*a));
    a->clear = 1;
    a->alpha = 1.0;
    a->kind_b = -1;
}

void gl_lin_elasticity(double young, double poisson, int two_dim, int plane_stress, double *dd) {
    lin_elasticity(young, poisson, two_dim, plane_stress, dd);
}

uint64_t gl_op_out_size(int op, int kind, int ndim) {
    if (op_size_class(op) < 0 || kind < 0 || kind >= GL_NKIND) return 0;
    unsigned r, c;
    op_extent(op, kind, ndim, nullptr, r, c);
    return (uint64_t)r * c;
}

uint64_t gl_op_coef_size(int op, int ndim) { return coef_len(op, ndim); }

} // extern "C" (reopened below)

// host -> device on the ctx stream: straight from pinned memory (asynchronous), through the pinned bounce buffers from pageable memory
// (defined next to upload_records)
int gl_h2d_any(gl_ctx *ctx, void *dst, const void *src, size_t bytes);

// ---- device-side structure-of-arrays transposition (upload path of homogeneous meshes) -----------------------------------
namespace {

template <typename IdT>
__global__ void conn_to_soa_kernel(const IdT *__restrict__ raw, uint32_t *__restrict__ soa, unsigned long long ncell, int nnode,
                                   unsigned long long npoint, int *__restrict__ bad) {
    // raw: [ncell][nnode] (Cell.points flattened); soa: [nnode][ncell].  Reads are coalesced over the flattened array,
    // writes are strided by nnode across a warp but stay within nnode cache lines per warp.
    const unsigned long long total = ncell * (unsigned long long)nnode;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long c = i / nnode;
        const int m = (int)(i - c * nnode);
        const unsigned long long pid = (unsigned long long)raw[i];
        if (pid >= npoint) *bad = 1;
        soa[(unsigned long long)m * ncell + c] = (uint32_t)pid;
    }
}

__global__ void coords_to_soa_kernel(const double *__restrict__ aos, double *__restrict__ soa, unsigned long long npoint, int ndim) {
    const unsigned long long total = npoint * (unsigned long long)ndim;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long pt = i / ndim;
        const int j = (int)(i - pt * ndim);
        soa[(unsigned long long)j * npoint + pt] = aos[i];
    }
}

__global__ void coords_to_point_major_kernel(const double *__restrict__ soa, double *__restrict__ pm, unsigned long long npoint, int ndim) {
    const int w = ndim == 2 ? 2 : 4;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < npoint;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        for (int j = 0; j < w; j++) pm[i * w + j] = j < ndim ? soa[(unsigned long long)j * npoint + i] : 0.0;
    }
}

// uploads a homogeneous mesh: raw host arrays go to the device as they are and are transposed there
// Homogeneous upload in two halves so that the caller can validate its host arrays while the copies are in flight:
// begin() enqueues the H2D copies and the SoA / narrowing kernels on the ctx stream, finish() waits and reports out-of-range ids.
struct PendingUpload {
    void *d_raw = nullptr;
    int *d_bad = nullptr;
    size_t raw_bytes = 0;
    int *h_bad = nullptr; // pinned
};

template <typename IdT>
int upload_homogeneous_begin(gl_ctx *ctx, gl_mesh *mesh, gl_mesh_bucket &bk, const double *coords, const IdT *conn, PendingUpload &pu) {
    cudaStream_t st = (cudaStream_t)ctx->stream;
    const size_t conn_bytes = (size_t)bk.ncell * bk.nnode * sizeof(IdT);
    const size_t coord_bytes = (size_t)mesh->npoint * mesh->ndim * sizeof(double);
    const size_t conn_at = (coord_bytes + 255) & ~(size_t)255;
    pu.raw_bytes = conn_at + conn_bytes; // two staging areas: the second copy does not wait for the first kernel
    CUDA_TRY(ctx, pool_alloc(ctx, &pu.d_raw, pu.raw_bytes));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &pu.d_bad, sizeof(int)));
    CUDA_TRY(ctx, cudaMemsetAsync(pu.d_bad, 0, sizeof(int), st));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &mesh->d_coords, coord_bytes));
    {
        int rc = gl_h2d_any(ctx, pu.d_raw, coords, coord_bytes);
        if (rc) return rc;
    }
    coords_to_soa_kernel<<<1184, 256, 0, st>>>((const double *)pu.d_raw, mesh->d_coords, mesh->npoint, mesh->ndim);
    if (bk.ncell) {
        char *d_conn_raw = (char *)pu.d_raw + conn_at;
        CUDA_TRY(ctx, pool_alloc_t(ctx, &bk.d_conn, (size_t)bk.ncell * bk.nnode * sizeof(uint32_t)));
        int rc = gl_h2d_any(ctx, d_conn_raw, conn, conn_bytes);
        if (rc) return rc;
        conn_to_soa_kernel<IdT><<<1184, 256, 0, st>>>((const IdT *)d_conn_raw, bk.d_conn, bk.ncell, bk.nnode, mesh->npoint, pu.d_bad);
    }
    ctx->launches += bk.ncell ? 2 : 1;
    return GL_OK;
}

int upload_homogeneous_finish(gl_ctx *ctx, PendingUpload &pu) {
    cudaStream_t st = (cudaStream_t)ctx->stream;
    int bad = 0;
    cudaError_t ce = cudaMemcpyAsync(&bad, pu.d_bad, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    pool_free(ctx, pu.d_raw, pu.raw_bytes);
    pool_free(ctx, pu.d_bad, sizeof(int));
    pu = PendingUpload();
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "mesh upload");
    if (bad) return fail(ctx, GL_ERR_MESH, "point id out of range");
    return GL_OK;
}

template <typename IdT>
int upload_homogeneous_device(gl_ctx *ctx, gl_mesh *mesh, gl_mesh_bucket &bk, const double *coords, const IdT *conn) {
    PendingUpload pu;
    int st = upload_homogeneous_begin<IdT>(ctx, mesh, bk, coords, conn, pu);
    if (st) return st;
    return upload_homogeneous_finish(ctx, pu);
}

} // namespace

extern "C" {

// =====================================================================================================
// mesh
// =====================================================================================================

static int finish_mesh_upload(gl_ctx *ctx, gl_mesh *mesh, const std::vector<double> &soa) {
    CUDA_TRY(ctx, pool_alloc_t(ctx, &mesh->d_coords, soa.size() * sizeof(double)));
    CUDA_TRY(ctx, cudaMemcpy(mesh->d_coords, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
    return GL_OK;
}

static void coords_to_soa(int ndim, uint64_t npoint, const double *coords, std::vector<double> &soa) {
    soa.resize((size_t)ndim * npoint);
    for (uint64_t i = 0; i < npoint; i++)
        for (int j = 0; j < ndim; j++) soa[(size_t)j * npoint + i] = coords[i * ndim + j];
}

static int build_markers(gl_ctx *ctx, gl_mesh *mesh, const int32_t *markers, std::vector<uint32_t> &marker_idx) {
    (void)ctx;
    marker_idx.clear();
    if (!markers || mesh->ncell == 0) {
        mesh->marker_values = {0};
        return GL_OK;
    }
    // distinct markers without sorting the whole array (meshes carry a handful of markers)
    std::vector<int32_t> uniq;
    {
        int32_t last = markers[0];
        uniq.push_back(last);
        for (uint64_t e = 1; e < mesh->ncell; e++) {
            const int32_t mk = markers[e];
            if (mk == last) continue;
            last = mk;
            if (std::find(uniq.begin(), uniq.end(), mk) == uniq.end()) uniq.push_back(mk);
        }
        std::sort(uniq.begin(), uniq.end());
    }
    mesh->marker_values = uniq;
    if (uniq.size() > 1) {
        marker_idx.resize(mesh->ncell);
        for (uint64_t e = 0; e < mesh->ncell; e++)
            marker_idx[e] = (uint32_t)(std::lower_bound(uniq.begin(), uniq.end(), markers[e]) - uniq.begin());
    }
    return GL_OK;
}

int gl_mesh_upload(gl_ctx *ctx, int ndim, uint64_t npoint, const double *coords, uint64_t ncell, const uint8_t *kinds,
                   const int32_t *markers, const uint64_t *conn_off, const uint64_t *conn, gl_mesh **out) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    if (!out || !coords || !kinds || !conn_off || !conn) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (ndim < 2 || ndim > 3) return fail(ctx, GL_ERR_SPACE_NDIM);
    if (npoint == 0 || npoint > 0xffffffffull || ncell > 0xffffffffull) return fail(ctx, GL_ERR_MESH, "npoint/ncell must fit in 32 bits");
    cudaSetDevice(ctx->device);
    // Single-kind meshes (the common case, and every step of a host-driven loop): the copies and the device-side transposition are
    // enqueued FIRST, the host-side validation of kinds / offsets / markers runs while they are in flight.
    bool one_kind = ncell > 0;
    for (uint64_t e = 1; e < ncell && one_kind; e++) one_kind = kinds[e] == kinds[0];
    if (one_kind) {
        const int k = kinds[0];
        if (k >= GL_NKIND || !kind_supported(k)) return fail(ctx, GL_ERR_KIND_UNSUPPORTED);
        if (kind_geo_ndim(k) > ndim) return fail(ctx, GL_ERR_MESH, "geo_ndim cannot be greater than space_ndim");
        auto mesh = std::make_unique<gl_mesh>();
        mesh->ndim = ndim;
        mesh->npoint = npoint;
        mesh->ncell = ncell;
        mesh->device = ctx->device;
        mesh->homogeneous = true;
        gl_mesh_bucket bk;
        bk.kind = k;
        bk.nnode = kind_nnode(k);
        bk.ncell = ncell;
        PendingUpload pu;
        int st = upload_homogeneous_begin<uint64_t>(ctx, mesh.get(), bk, coords, conn + conn_off[0], pu);
        bool offsets_ok = true;
        if (!st) {
            const uint64_t nn = (uint64_t)bk.nnode;
            uint64_t prev = conn_off[0];
            for (uint64_t e = 0; e < ncell; e++) {
                const uint64_t next = conn_off[e + 1];
                offsets_ok = offsets_ok && (next - prev == nn);
                prev = next;
            }
        }
        std::vector<uint32_t> marker_idx;
        if (!st) build_markers(ctx, mesh.get(), markers, marker_idx);
        if (pu.d_raw) {
            const int st2 = upload_homogeneous_finish(ctx, pu);
            if (!st) st = st2;
        }
        if (!st && !offsets_ok) st = fail(ctx, GL_ERR_MESH, "cell connectivity length != nnode");
        if (!st && !marker_idx.empty()) {
            if (cudaMalloc(&bk.d_marker_idx, marker_idx.size() * sizeof(uint32_t)) != cudaSuccess ||
                cudaMemcpy(bk.d_marker_idx, marker_idx.data(), marker_idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice) != cudaSuccess)
                st = cuda_fail(ctx, cudaGetLastError(), "marker table");
        }
        mesh->buckets.push_back(std::move(bk));
        if (st) {
            gl_mesh_free(ctx, mesh.release());
            return st;
        }
        *out = mesh.release();
        return GL_OK;
    }
    // validate and count per kind
    uint64_t count[GL_NKIND] = {0};
    {
        uint64_t nn_of[256];
        unsigned char ok_of[256];
        for (int k = 0; k < 256; k++) {
            const bool known = k < GL_NKIND && kind_supported(k);
            nn_of[k] = known ? (uint64_t)kind_nnode(k) : 0;
            ok_of[k] = !known ? 0 : (kind_geo_ndim(k) > ndim ? 1 : 2);
        }
        uint64_t prev = conn_off[0];
        for (uint64_t e = 0; e < ncell; e++) {
            const unsigned k = kinds[e];
            const uint64_t next = conn_off[e + 1];
            if (ok_of[k] != 2 || next - prev != nn_of[k]) {
                if (ok_of[k] == 0) return fail(ctx, GL_ERR_KIND_UNSUPPORTED);
                if (ok_of[k] == 1) return fail(ctx, GL_ERR_MESH, "geo_ndim cannot be greater than space_ndim");
                return fail(ctx, GL_ERR_MESH, "cell connectivity length != nnode");
            }
            prev = next;
            count[k]++;
        }
    }
    auto mesh = std::make_unique<gl_mesh>();
    mesh->ndim = ndim;
    mesh->npoint = npoint;
    mesh->ncell = ncell;
    mesh->device = ctx->device;
    int nkinds = 0;
    for (int k = 0; k < GL_NKIND; k++) nkinds += count[k] ? 1 : 0;
    mesh->homogeneous = nkinds <= 1;
    std::vector<uint32_t> marker_idx;
    build_markers(ctx, mesh.get(), markers, marker_idx);
    if (!mesh->homogeneous) mesh->h_kinds.assign(kinds, kinds + ncell);
    for (int k = 0; k < GL_NKIND; k++) {
        if (!count[k]) continue;
        gl_mesh_bucket bk;
        bk.kind = k;
        bk.nnode = kind_nnode(k);
        bk.ncell = count[k];
        std::vector<uint32_t> ids;
        ids.reserve(count[k]);
        for (uint64_t e = 0; e < ncell; e++)
            if (kinds[e] == k) ids.push_back((uint32_t)e);
        std::vector<uint32_t> soa((size_t)bk.nnode * bk.ncell);
        for (uint64_t c = 0; c < bk.ncell; c++) {
            const uint64_t *pts = conn + conn_off[ids[c]];
            for (int m = 0; m < bk.nnode; m++) {
                if (pts[m] >= npoint) return fail(ctx, GL_ERR_MESH, "point id out of range");
                soa[(size_t)m * bk.ncell + c] = (uint32_t)pts[m];
            }
        }
        CUDA_TRY(ctx, pool_alloc_t(ctx, &bk.d_conn, soa.size() * sizeof(uint32_t)));
        CUDA_TRY(ctx, cudaMemcpy(bk.d_conn, soa.data(), soa.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        if (!marker_idx.empty()) {
            std::vector<uint32_t> mi(bk.ncell);
            for (uint64_t c = 0; c < bk.ncell; c++) mi[c] = marker_idx[ids[c]];
            CUDA_TRY(ctx, cudaMalloc(&bk.d_marker_idx, mi.size() * sizeof(uint32_t)));
            CUDA_TRY(ctx, cudaMemcpy(bk.d_marker_idx, mi.data(), mi.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
        }
        if (!mesh->homogeneous) {
            CUDA_TRY(ctx, cudaMalloc(&bk.d_cell_ids, ids.size() * sizeof(uint32_t)));
            CUDA_TRY(ctx, cudaMemcpy(bk.d_cell_ids, ids.data(), ids.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
            bk.h_cell_ids = std::move(ids);
        }
        mesh->buckets.push_back(std::move(bk));
    }
    std::vector<double> soa;
    coords_to_soa(ndim, npoint, coords, soa);
    int st = finish_mesh_upload(ctx, mesh.get(), soa);
    if (st) return st;
    *out = mesh.release();
    return GL_OK;
}

int gl_mesh_upload_homogeneous(gl_ctx *ctx, int ndim, uint64_t npoint, const double *coords, uint64_t ncell, int kind,
                               const int32_t *markers, const uint32_t *conn, gl_mesh **out) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    if (!out || !coords || !conn) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (ndim < 2 || ndim > 3) return fail(ctx, GL_ERR_SPACE_NDIM);
    if (!kind_supported(kind)) return fail(ctx, GL_ERR_KIND_UNSUPPORTED);
    if (kind_geo_ndim(kind) > ndim) return fail(ctx, GL_ERR_MESH, "geo_ndim cannot be greater than space_ndim");
    if (npoint == 0 || npoint > 0xffffffffull || ncell > 0xffffffffull) return fail(ctx, GL_ERR_MESH, "npoint/ncell must fit in 32 bits");
    cudaSetDevice(ctx->device);
    auto mesh = std::make_unique<gl_mesh>();
    mesh->ndim = ndim;
    mesh->npoint = npoint;
    mesh->ncell = ncell;
    mesh->device = ctx->device;
    mesh->homogeneous = true;
    std::vector<uint32_t> marker_idx;
    build_markers(ctx, mesh.get(), markers, marker_idx);
    gl_mesh_bucket bk;
    bk.kind = kind;
    bk.nnode = kind_nnode(kind);
    bk.ncell = ncell;
    {
        int st = upload_homogeneous_device<uint32_t>(ctx, mesh.get(), bk, coords, conn);
        if (st) return st;
    }
    if (ncell && !marker_idx.empty()) {
        CUDA_TRY(ctx, cudaMalloc(&bk.d_marker_idx, marker_idx.size() * sizeof(uint32_t)));
        CUDA_TRY(ctx, cudaMemcpy(bk.d_marker_idx, marker_idx.data(), marker_idx.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    }
    mesh->buckets.push_back(std::move(bk));
    *out = mesh.release();
    return GL_OK;
}

int gl_mesh_update_coords(gl_ctx *ctx, gl_mesh *mesh, const double *coords, int location) {
    if (!ctx || !mesh || !coords) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (location != GL_HOST) return fail(ctx, GL_ERR_INVALID_ARG, "gl_mesh_update_coords takes host AoS coordinates");
    cudaSetDevice(ctx->device);
    std::vector<double> soa;
    coords_to_soa(mesh->ndim, mesh->npoint, coords, soa);
    CUDA_TRY(ctx, cudaStreamSynchronize((cudaStream_t)ctx->stream));
    CUDA_TRY(ctx, cudaMemcpy(mesh->d_coords, soa.data(), soa.size() * sizeof(double), cudaMemcpyHostToDevice));
    if (mesh->d_coords_pm) {
        coords_to_point_major_kernel<<<1184, 256, 0, (cudaStream_t)ctx->stream>>>(mesh->d_coords, mesh->d_coords_pm, mesh->npoint, mesh->ndim);
        ctx->launches++;
    }
    return GL_OK;
}

void gl_mesh_free(gl_ctx *ctx, gl_mesh *mesh) {
    if (!mesh) return;
    if (ctx) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize((cudaStream_t)ctx->stream);
    }
    // the two large buffers go back to the ctx pool (all work on them is complete: the stream was just synchronised)
    pool_free(ctx, mesh->d_coords, (size_t)mesh->npoint * mesh->ndim * sizeof(double));
    pool_free(ctx, mesh->d_coords_pm, (size_t)mesh->npoint * (mesh->ndim == 2 ? 2 : 4) * sizeof(double));
    for (auto &bk : mesh->buckets) {
        pool_free(ctx, bk.d_conn, (size_t)bk.ncell * bk.nnode * sizeof(uint32_t));
        if (bk.d_cell_ids) cudaFree(bk.d_cell_ids);
        if (bk.d_marker_idx) cudaFree(bk.d_marker_idx);
        for (auto &kv : bk.d_out_off)
            if (kv.second) cudaFree(kv.second);
    }
    delete mesh;
}

uint64_t gl_mesh_ncell(const gl_mesh *mesh) { return mesh ? mesh->ncell : 0; }
uint64_t gl_mesh_npoint(const gl_mesh *mesh) { return mesh ? mesh->npoint : 0; }
int gl_mesh_ndim(const gl_mesh *mesh) { return mesh ? mesh->ndim : 0; }

// =====================================================================================================
// output layout
// =====================================================================================================

static int range_layout(const gl_mesh *cmesh, int op, uint64_t b, uint64_t e, const gl_args *a, uint64_t *total,
                        uint64_t *off /* may be null */) {
    gl_mesh *mesh = const_cast<gl_mesh *>(cmesh);
    if (!mesh || !a || op_size_class(op) < 0) return GL_ERR_INVALID_ARG;
    if (b > e || e > mesh->ncell) return GL_ERR_CELL_RANGE;
    const bool explicit_block = a->nrow != 0 || (op_is_matrix(op) && a->ncol != 0);
    const bool packed = a->packed && op_is_matrix(op); // vectors have no triangle: the flag is ignored for them
    if (packed) {
        // upper triangles only: square tight blocks of one kind, overwritten
        if (explicit_block || a->ii0 || a->jj0 || !a->clear || !mesh->homogeneous) return GL_ERR_INVALID_ARG;
    }
    if (mesh->homogeneous || explicit_block) {
        // all blocks equal (for explicit dims every kind in range must fit; checked at launch)
        uint64_t blk = 0;
        if (!mesh->buckets.empty()) {
            BlockDims bd;
            int st = block_dims(op, mesh->buckets[0].kind, mesh->ndim, a, bd);
            if (st) return st;
            blk = packed ? (uint64_t)bd.size_r * (bd.size_r + 1) / 2 : bd.block();
        }
        if (total) *total = blk * (e - b);
        if (off)
            for (uint64_t i = 0; i <= e - b; i++) off[i] = i * blk;
        return GL_OK;
    }
    BlockKey key;
    int st;
    const std::vector<uint64_t> *pre = mesh_prefix(mesh, op, a, &key, &st);
    if (st) return st;
    if (total) *total = (*pre)[e] - (*pre)[b];
    if (off)
        for (uint64_t i = 0; i <= e - b; i++) off[i] = (*pre)[b + i] - (*pre)[b];
    return GL_OK;
}

int gl_out_total(const gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a, uint64_t *total) {
    if (!total) return GL_ERR_INVALID_ARG;
    return range_layout(mesh, op, b, e, a, total, nullptr);
}

int gl_out_offsets(const gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a, uint64_t *off) {
    if (!off) return GL_ERR_INVALID_ARG;
    return range_layout(mesh, op, b, e, a, nullptr, off);
}

// =====================================================================================================
// integration
// =====================================================================================================

// builds the point-major coordinate copy of a mesh on first use (thread-per-cell kernels gather whole points)
static int ensure_point_major(gl_ctx *ctx, gl_mesh *mesh) {
    if (mesh->d_coords_pm) return GL_OK;
    CUDA_TRY(ctx, pool_alloc_t(ctx, &mesh->d_coords_pm, (size_t)mesh->npoint * (mesh->ndim == 2 ? 2 : 4) * sizeof(double)));
    coords_to_point_major_kernel<<<1184, 256, 0, (cudaStream_t)ctx->stream>>>(mesh->d_coords, mesh->d_coords_pm, mesh->npoint, mesh->ndim);
    ctx->launches++;
    return GL_OK;
}

// gl_ctx_last_kernel: the kernel of the most recent launch; inside gl_integ_fused, every distinct kernel of the call joined by " + "
static void set_last_kernel(gl_ctx *ctx) {
    const std::string k = noted_kernel();
    if (!ctx->collect_kernels || ctx->last_kernel.empty()) {
        ctx->last_kernel = k;
        return;
    }
    const std::string hay = " + " + ctx->last_kernel + " + ";
    if (hay.find(" + " + k + " + ") == std::string::npos) ctx->last_kernel += " + " + k;
}
struct KernelCollector { // scope of one gl_integ_fused call
    gl_ctx *ctx;
    explicit KernelCollector(gl_ctx *c) : ctx(c) {
        ctx->collect_kernels = true;
        ctx->last_kernel.clear();
    }
    ~KernelCollector() { ctx->collect_kernels = false; }
};

// block offsets of a bucket's cells on the device (mixed meshes), cached per block layout
static int bucket_device_offsets(gl_ctx *ctx, gl_mesh_bucket &bk, const std::vector<uint64_t> *pre, const BlockKey &key,
                                 const unsigned long long **out) {
    auto it = bk.d_out_off.find(key);
    if (it == bk.d_out_off.end()) {
        std::vector<unsigned long long> offs(bk.ncell);
        for (uint64_t c = 0; c < bk.ncell; c++) offs[c] = (*pre)[bk.h_cell_ids[c]];
        unsigned long long *d = nullptr;
        CUDA_TRY(ctx, cudaMalloc(&d, offs.size() * sizeof(unsigned long long)));
        CUDA_TRY(ctx, cudaMemcpy(d, offs.data(), offs.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
        it = bk.d_out_off.emplace(key, d).first;
    }
    *out = it->second;
    return GL_OK;
}

// gl_integ_fused, stiffness + conductivity matrix + flux vector: the scalar cases the Qua8 buckets take from the stiffness kernel's
// geometric tensors (gl_kernels_qua8.cu); `done` marks the buckets served that way
struct Qua8Fusion {
    const ScalarFused *f;
    const std::vector<uint64_t> *mpre, *vpre;
    BlockKey mkey, vkey;
    std::vector<char> *done;
};

// launches the kernels that integrate original cells [b, e) into the DEVICE buffer d_out (block of cell b at d_out[0])
static int launch_range(gl_ctx *ctx, gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a,
                        const CoefResolved &cr, uint64_t coef_begin, double *d_out, const double *d_sigma = nullptr,
                        double *d_force = nullptr, bool *force_done = nullptr, const uint32_t *sc_pos = nullptr, double *sc_vals = nullptr,
                        int only_bucket = -1, Qua8Fusion *qf = nullptr) {
    if (b == e) return GL_OK;
    const bool explicit_block = a->nrow != 0 || (op_is_matrix(op) && a->ncol != 0);
    const std::vector<uint64_t> *pre = nullptr;
    BlockKey key;
    if (!mesh->homogeneous && !explicit_block) {
        int st;
        pre = mesh_prefix(mesh, op, a, &key, &st);
        if (st) return fail(ctx, st);
    }
    for (size_t bi = 0; bi < mesh->buckets.size(); bi++) {
        auto &bk = mesh->buckets[bi];
        if (only_bucket >= 0 && (int)bi != only_bucket) continue;
        uint64_t lo, hi;
        bucket_range(mesh, bk, b, e, lo, hi);
        if (lo >= hi) continue;
        const int gd = kind_geo_ndim(bk.kind);
        BlockDims bd;
        int st = block_dims(op, bk.kind, mesh->ndim, a, bd);
        if (st) return fail(ctx, st);
        if (op == GL_OP_MAT_10_BDB || op == GL_OP_VEC_04_BT)
            if (a->axisymmetric && mesh->ndim != 2) return fail(ctx, GL_ERR_AXISYM_NDIM);
        if (op_boundary(op)) {
            if (mesh->ndim == 2 && gd != 1) return fail(ctx, op == GL_OP_NORMAL ? GL_ERR_NORMAL_2D : GL_ERR_BRY_2D);
            if (mesh->ndim == 3 && gd != 2) return fail(ctx, op == GL_OP_NORMAL ? GL_ERR_NORMAL_3D : GL_ERR_BRY_3D);
        }
        if (op_uses_gradient(op) && gd != mesh->ndim) return fail(ctx, GL_ERR_GRADIENT_NDIM);
        if (op_two_pads(op)) {
            if (a->kind_b < 0 || a->kind_b >= GL_NKIND) return fail(ctx, GL_ERR_KIND_UNSUPPORTED, "args.kind_b is not a GeoKind");
            if (kind_geo_ndim(a->kind_b) != gd || kind_nnode(a->kind_b) > bk.nnode)
                return fail(ctx, GL_ERR_INVALID_ARG, "the second pad must have the cell's geo_ndim and at most its nodes");
        }

        IntegParams p;
        std::memset(&p, 0, sizeof(p));
        p.coords = mesh->d_coords;
        p.npoint = mesh->npoint;
        if (op == GL_OP_MAT_01_NSN || op == GL_OP_VEC_01_NS || op == GL_OP_MAT_03_BTB || op == GL_OP_VEC_03_BV ||
            ((op == GL_OP_VEC_02_NV || op == GL_OP_VEC_04_BT) && gd == mesh->ndim) ||
            (op == GL_OP_MAT_10_BDB && mesh->ndim == 2 && (bk.nnode == 3 || bk.nnode == 4))) { // thread-per-cell kernels
            st = ensure_point_major(ctx, mesh);
            if (st) return st;
            p.coords_pm = mesh->d_coords_pm;
        }
        p.conn = bk.d_conn;
        p.ncell_k = bk.ncell;
        p.cell_ids = bk.d_cell_ids;
        p.lo = lo;
        p.hi = hi;
        p.cell_begin = b;
        p.coef_begin = coef_begin;
        p.op = op;
        p.axisym = a->axisymmetric ? 1 : 0;
        p.clear = a->clear ? 1 : 0;
        p.alpha = a->alpha;
        p.nnode = bk.nnode;
        if (op == GL_OP_JACOBIAN) {
            // one-point "rule" at args->ksi: w = 1 | N | L
            double tmp[1 + MAX_NNODE + 3 * MAX_NNODE];
            tmp[0] = 1.0;
            st = shape_eval(bk.kind, a->ksi, tmp + 1, tmp + 1 + bk.nnode);
            if (st) return fail(ctx, st);
            double *dst = ctx->d_rule_tmp + (size_t)bk.kind * (1 + MAX_NNODE + 3 * MAX_NNODE);
            CUDA_TRY(ctx, cudaMemcpyAsync(dst, tmp, sizeof(double) * (1 + bk.nnode + bk.nnode * gd), cudaMemcpyHostToDevice,
                                          (cudaStream_t)ctx->stream));
            p.rule = dst;
            p.ng = 1;
        } else {
            RuleTable *rt = get_rule(ctx, bk.kind, a->ngauss, &st);
            if (!rt) return fail(ctx, st);
            p.rule = rt->dev;
            p.ng = rt->ng;
        }
        if (op_two_pads(op)) {
            RuleTable *rb = get_rule_b(ctx, bk.kind, a->kind_b, a->ngauss, &st);
            if (!rb) return fail(ctx, st);
            p.rule_b = rb->dev;
            p.nnode_b = rb->nnode;
        }
        // coefficient
        p.coef_len = cr.len;
        p.coef_pat = cr.pat;
        p.coef_pat_dev = cr.pat_dev;
        p.pg_flag = ctx->d_coef_pat;
        std::memcpy(p.coef_uniform, cr.uniform, sizeof(p.coef_uniform));
        p.coef = cr.dev;
        p.coef_cell_stride = cr.cell_stride;
        p.coef_gp_stride = cr.gp_stride;
        p.coef_index = cr.per_marker ? bk.d_marker_idx : nullptr;
        if (cr.dev && cr.gp_stride && cr.cell_stride == 0) p.coef_cell_stride = (uint64_t)p.ng * cr.gp_stride; // PER_GAUSS
        // output
        p.out = d_out;
        p.nrow = bd.nrow;
        p.ncol = bd.ncol;
        p.ii0 = a->ii0;
        p.jj0 = op_is_matrix(op) ? a->jj0 : 0;
        p.size_r = bd.size_r;
        p.size_c = bd.size_c;
        p.err_cell = ctx->d_err_cell;
        p.work_counter = ctx->d_work_counter;
        if (pre) {
            auto it = bk.d_out_off.find(key);
            if (it == bk.d_out_off.end()) {
                std::vector<unsigned long long> offs(bk.ncell);
                for (uint64_t c = 0; c < bk.ncell; c++) offs[c] = (*pre)[bk.h_cell_ids[c]];
                unsigned long long *d = nullptr;
                CUDA_TRY(ctx, cudaMalloc(&d, offs.size() * sizeof(unsigned long long)));
                CUDA_TRY(ctx, cudaMemcpy(d, offs.data(), offs.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
                it = bk.d_out_off.emplace(key, d).first;
            }
            p.out_off = it->second;
            p.out_base = (*pre)[b];
        }
        int nl = 0;
        p.sigma = d_sigma; // stiffness + internal force from one pass: only the Hex8 kernel takes them
        p.out_d = d_force;
        if (sc_pos) {
            // fused assembly (gl_assemble_planned): only the kernels with a scatter epilogue; the caller checked stiffness_fusable()
            p.scatter_pos = sc_pos;
            p.scatter_vals = sc_vals;
            nl = launch_hex8_bdb(p, ctx->stream);
            if (nl == 0) nl = launch_bdb_qua8(p, mesh->ndim, ctx->stream);
            if (nl == 0) nl = launch_bdb(p, mesh->ndim, ctx->stream);
            if (nl == 0) return fail(ctx, GL_ERR_INVALID_ARG, "internal: no kernel with a scatter epilogue took the configuration");
            if (nl < 0) return cuda_fail(ctx, cudaGetLastError(), "kernel launch");
            ctx->launches += (uint64_t)nl;
            set_last_kernel(ctx);
            continue;
        }
        if (qf && op == GL_OP_MAT_10_BDB && mesh->ndim == 2 && (bk.kind == GL_QUA8 || bk.kind == GL_TRI6)) {
            ScalarFused fb = *qf->f;
            if (qf->mpre) {
                st = bucket_device_offsets(ctx, bk, qf->mpre, qf->mkey, &fb.moff);
                if (st) return st;
                fb.mbase = (*qf->mpre)[b];
            }
            if (qf->vpre) {
                st = bucket_device_offsets(ctx, bk, qf->vpre, qf->vkey, &fb.voff);
                if (st) return st;
                fb.vbase = (*qf->vpre)[b];
            }
            nl = launch_bdb_qua8(p, mesh->ndim, ctx->stream, &fb);
            if (nl > 0) (*qf->done)[bi] = 1;
        }
        if (nl == 0 && op == GL_OP_MAT_10_BDB) nl = launch_hex8_bdb(p, ctx->stream);
        if (nl == 0 && op == GL_OP_MAT_10_BDB && d_sigma == nullptr) nl = launch_hex8_gauss(p, ctx->stream);
        if (force_done) *force_done = nl > 0 && d_sigma != nullptr;
        if (nl == 0 && op == GL_OP_MAT_10_BDB && gd == mesh->ndim) nl = launch_bdb_small2d(p, mesh->ndim, ctx->stream);
        if (nl == 0 && op == GL_OP_MAT_10_BDB && gd == mesh->ndim) nl = launch_bdb_syrk(p, mesh->ndim, ctx->stream);
        if (nl == 0 && op == GL_OP_MAT_10_BDB && gd == mesh->ndim) nl = launch_bdb_qua8(p, mesh->ndim, ctx->stream);
        if (nl == 0 && op == GL_OP_MAT_10_BDB && gd == mesh->ndim) nl = launch_bdb_tile(p, mesh->ndim, ctx->stream);
        if (nl == 0 && op == GL_OP_MAT_10_BDB && gd == mesh->ndim) nl = launch_bdb(p, mesh->ndim, ctx->stream);
        if (nl == 0 && gd == mesh->ndim && cr.dev == nullptr && a->clear && a->ii0 == 0 && (a->jj0 == 0 || !op_is_matrix(op)) &&
            a->nrow == 0 && a->ncol == 0 &&
            (op == GL_OP_MAT_01_NSN || op == GL_OP_MAT_03_BTB || op == GL_OP_VEC_01_NS || op == GL_OP_VEC_03_BV)) {
            ScalarFused f;
            std::memset(&f, 0, sizeof(f));
            if (op == GL_OP_MAT_01_NSN) { f.mask = 1; f.s_mat = cr.uniform[0]; f.out_m01 = d_out; }
            if (op == GL_OP_MAT_03_BTB) { f.mask = 2; std::memcpy(f.t, cr.uniform, sizeof(double) * 2 * mesh->ndim); f.out_m03 = d_out; }
            if (op == GL_OP_VEC_01_NS) { f.mask = 4; f.s_vec = cr.uniform[0]; f.out_v01 = d_out; }
            if (op == GL_OP_VEC_03_BV) { f.mask = 8; std::memcpy(f.w, cr.uniform, sizeof(double) * mesh->ndim); f.out_v03 = d_out; }
            if (op_is_matrix(op)) { f.moff = p.out_off; f.mbase = p.out_base; }
            else { f.voff = p.out_off; f.vbase = p.out_base; }
            nl = launch_mass_dmma(p, f, mesh->ndim, ctx->stream);
            if (nl == 0) nl = launch_scalar_tpc(p, f, mesh->ndim, ctx->stream);
            if (nl == 0) nl = launch_scalar_fused(p, f, mesh->ndim, ctx->stream);
        }
        if (nl == 0 && gd == mesh->ndim) nl = launch_vec_tpc(p, mesh->ndim, ctx->stream);
        if (nl == 0) {
            p.cells_per_cta = generic_pick_cells_per_cta(mesh->ndim, gd, p.nnode, p.ng, op, p.nnode_b);
            nl = launch_generic(p, mesh->ndim, gd, ctx->stream);
        }
        if (nl < 0) return cuda_fail(ctx, cudaGetLastError(), "kernel launch");
        ctx->launches += (uint64_t)nl;
        set_last_kernel(ctx);
    }
    return GL_OK;
}

// structure shared by all ns x ns Mandel moduli of a host table (column-major records): 2 = major-symmetric with
// uncoupled normal / diagonal shear blocks, 1 = major-symmetric, 0 = general
static int modulus_pattern(const double *records, uint64_t nrec, int ns) {
    int pat = 2;
    for (uint64_t r = 0; r < nrec && pat > 0; r++) {
        const double *d = records + r * (uint64_t)(ns * ns);
        for (int a = 0; a < ns && pat > 0; a++)
            for (int b = 0; b < ns; b++) {
                if (d[a + b * ns] != d[b + a * ns]) { pat = 0; break; }
                const bool nz = (a < 3 && b < 3) || a == b;
                if (!nz && d[a + b * ns] != 0.0) pat = 1;
            }
    }
    return pat;
}

// two pinned 16 MB bounce buffers + their events: pageable host memory <-> device (upload_records, drain_to_pageable)
static int ensure_bounce(gl_ctx *ctx) {
    for (int i = 0; i < 2; i++) {
        if (!ctx->h_rec[i]) {
            CUDA_TRY(ctx, cudaMallocHost(&ctx->h_rec[i], 16u << 20));
            cudaEvent_t ev;
            CUDA_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
            ctx->ev_rec[i] = ev;
        }
    }
    return GL_OK;
}

// work(0) .. work(nthr - 1), on up to nthr host threads; a thread that cannot be started (resource limits) has its share done inline
// -- nothing may throw across the C ABI
static void run_parallel(int nthr, const std::function<void(int)> &work) {
    std::thread th[7];
    bool started[7] = {false, false, false, false, false, false, false};
    for (int t = 1; t < nthr && t <= 7; t++) {
        try {
            th[t - 1] = std::thread(work, t);
            started[t - 1] = true;
        } catch (...) {
            started[t - 1] = false;
        }
    }
    work(0);
    for (int t = 1; t < nthr && t <= 7; t++) {
        if (started[t - 1]) th[t - 1].join();
        else work(t);
    }
}

static void parallel_memcpy(void *dst, const void *src, size_t bytes) {
    const unsigned hw = std::thread::hardware_concurrency();
    const int nthr = bytes < ((size_t)4 << 20) ? 1 : (int)std::max(1u, std::min(8u, hw ? hw : 1u));
    if (nthr == 1) {
        std::memcpy(dst, src, bytes);
        return;
    }
    auto work = [&](int t) {
        const size_t a = (bytes * (size_t)t / nthr) & ~(size_t)63, b = t == nthr - 1 ? bytes : (bytes * (size_t)(t + 1) / nthr) & ~(size_t)63;
        std::memcpy((char *)dst + a, (const char *)src + a, b - a);
    };
    run_parallel(nthr, work);
}

static bool host_pointer_is_pinned(const void *p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
}

// Device -> PAGEABLE host memory (a Rust Vec, a numpy array): a cudaMemcpyAsync straight into it runs at 15 GB/s against 52 GB/s for
// pinned memory.  Here 16 MB pieces go through the two pinned bounce buffers on `s_copy`; host threads copy piece i out while the
// DMA of piece i + 1 runs.  Returns when the data is in dst.
static int drain_to_pageable(gl_ctx *ctx, double *dst, const double *src_dev, uint64_t n, cudaStream_t s_copy) {
    int stb = ensure_bounce(ctx);
    if (stb) return stb;
    const uint64_t piece = (16u << 20) / sizeof(double);
    const uint64_t np = (n + piece - 1) / piece;
    for (uint64_t q = 0; q < np; q++) {
        const int j = (int)(q & 1);
        const uint64_t len = std::min(piece, n - q * piece);
        CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)ctx->ev_rec[j])); // an earlier user of the buffer (an upload) has finished
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_rec[j], src_dev + q * piece, len * sizeof(double), cudaMemcpyDeviceToHost, s_copy));
        CUDA_TRY(ctx, cudaEventRecord((cudaEvent_t)ctx->ev_rec[j], s_copy));
        if (q >= 1) {
            CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)ctx->ev_rec[j ^ 1]));
            parallel_memcpy(dst + (q - 1) * piece, ctx->h_rec[j ^ 1], piece * sizeof(double));
        }
    }
    if (np > 0) {
        const int j = (int)((np - 1) & 1);
        CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)ctx->ev_rec[j]));
        parallel_memcpy(dst + (np - 1) * piece, ctx->h_rec[j], (n - (np - 1) * piece) * sizeof(double));
    }
    return GL_OK;
}

// Host-resident records -> device.  A plain cudaMemcpyAsync from pageable memory crawls (3.5 GB/s with the structure scan in front
// of it: 36 ms per 10^6 2D records); here the records pass through two pinned bounce buffers, a few host threads
// copy (and, for stiffness records, scan: ns > 0) chunk i + 1 while the DMA of chunk i runs.  *pat as modulus_pattern().
static int upload_records(gl_ctx *ctx, const double *src, uint64_t ndoubles, double *dst, int ns, int *pat) {
    const uint64_t nn = ns > 0 ? (uint64_t)ns * ns : 1;
    if (pat) *pat = ns > 0 ? 2 : 0;
    const uint64_t small = (1u << 20) / sizeof(double);
    if (ndoubles <= small) { // not worth the threads
        if (ns > 0 && pat) *pat = modulus_pattern(src, ndoubles / nn, ns);
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, ndoubles * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)ctx->stream));
        return GL_OK;
    }
    const uint64_t chunk = ((16u << 20) / sizeof(double) / nn) * nn; // whole records per chunk
    int stb = ensure_bounce(ctx);
    if (stb) return stb;
    const unsigned hw = std::thread::hardware_concurrency();
    const int nthr = (int)std::max(1u, std::min(8u, hw ? hw : 1u));
    int pats[8] = {2, 2, 2, 2, 2, 2, 2, 2};
    uint64_t k = 0;
    for (uint64_t off = 0; off < ndoubles; off += chunk, k++) {
        const uint64_t n = std::min(chunk, ndoubles - off);
        double *buf = ctx->h_rec[k & 1];
        CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)ctx->ev_rec[k & 1])); // the last DMA out of this buffer (this call's or an earlier one's) has finished
        const uint64_t nrec = n / nn;
        auto work = [&](int t) {
            const uint64_t r0 = nrec * (uint64_t)t / nthr, r1 = nrec * (uint64_t)(t + 1) / nthr;
            const uint64_t d0 = r0 * nn, d1 = t == nthr - 1 ? n : r1 * nn;
            std::memcpy(buf + d0, src + off + d0, (d1 - d0) * sizeof(double));
            if (ns > 0 && pats[t] > 0) pats[t] = std::min(pats[t], modulus_pattern(buf + d0, r1 - r0, ns));
        };
        run_parallel(nthr, work);
        CUDA_TRY(ctx, cudaMemcpyAsync(dst + off, buf, n * sizeof(double), cudaMemcpyHostToDevice, (cudaStream_t)ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord((cudaEvent_t)ctx->ev_rec[k & 1], (cudaStream_t)ctx->stream));
    }
    if (ns > 0 && pat)
        for (int t = 0; t < nthr; t++) *pat = std::min(*pat, pats[t]);
    return GL_OK;
}

} // extern "C"

int gl_h2d_any(gl_ctx *ctx, void *dst, const void *src, size_t bytes) {
    cudaStream_t st = (cudaStream_t)ctx->stream;
    if (bytes < ((size_t)1 << 20) || host_pointer_is_pinned(src)) {
        CUDA_TRY(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
        return GL_OK;
    }
    const size_t n8 = bytes / sizeof(double);
    int rc = upload_records(ctx, (const double *)src, n8, (double *)dst, 0, nullptr);
    if (rc) return rc;
    if (bytes > n8 * sizeof(double))
        CUDA_TRY(ctx, cudaMemcpyAsync((char *)dst + n8 * sizeof(double), (const char *)src + n8 * sizeof(double), bytes - n8 * sizeof(double),
                                      cudaMemcpyHostToDevice, st));
    return GL_OK;
}

// modulus_pattern() for records that live on the device: *pat starts at 2 and is lowered by the records that break a pattern.
// The verdict stays on the device -- the stiffness launchers enqueue all three pattern variants and the two that do not
// match return at once -- so the call remains asynchronous.
__global__ void modulus_pattern_reset_kernel(int *pat) { *pat = 2; }

__global__ void modulus_pattern_kernel(const double *__restrict__ records, unsigned long long nrec, int ns, int *pat) {
    const int nn = ns * ns;
    int mine = 2;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nrec * (unsigned long long)nn;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long r = i / nn;
        const int e = (int)(i - r * nn), a = e % ns, b = e / ns;
        const double v = records[i];
        if (v != records[r * nn + b + a * ns]) mine = 0;
        else if (!((a < 3 && b < 3) || a == b) && v != 0.0 && mine > 1) mine = 1;
    }
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 16));
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 8));
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 4));
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 2));
    mine = min(mine, __shfl_xor_sync(0xffffffffu, mine, 1));
    if ((threadIdx.x & 31) == 0 && mine < 2) atomicMin(pat, mine);
}

extern "C" {

static int resolve_coef(gl_ctx *ctx, gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef,
                        CoefResolved &cr) {
    const uint64_t len = coef_len(op, mesh->ndim);
    cr.len = (int)len;
    if (len == 0) return GL_OK; // GL_OP_JACOBIAN
    if (!coef || !coef->data) return fail(ctx, GL_ERR_COEF, "coefficient data is required");
    switch (coef->mode) {
    case GL_COEF_UNIFORM:
        if (coef->location != GL_HOST) return fail(ctx, GL_ERR_COEF, "UNIFORM coefficient must be a host record");
        std::memcpy(cr.uniform, coef->data, len * sizeof(double));
        return GL_OK;
    case GL_COEF_PER_MARKER: {
        if (coef->location != GL_HOST || !coef->markers || coef->nrecord == 0)
            return fail(ctx, GL_ERR_COEF, "PER_MARKER needs host records and a host marker list");
        const size_t nm = mesh->marker_values.size();
        std::vector<double> table(nm * len, 0.0);
        for (size_t k = 0; k < nm; k++) {
            bool found = false;
            for (uint64_t r = 0; r < coef->nrecord && !found; r++)
                if (coef->markers[r] == mesh->marker_values[k]) {
                    std::memcpy(&table[k * len], coef->data + r * len, len * sizeof(double));
                    found = true;
                }
            if (!found) return fail(ctx, GL_ERR_COEF, "a cell marker has no coefficient record");
        }
        if (nm == 1) {
            std::memcpy(cr.uniform, table.data(), len * sizeof(double));
            return GL_OK;
        }
        int st = ensure_buffer(ctx, &ctx->d_coef, &ctx->coef_bytes, table.size() * sizeof(double));
        if (st) return st;
        if (op == GL_OP_MAT_10_BDB) cr.pat = modulus_pattern(table.data(), nm, 2 * mesh->ndim);
        // the table goes through a pinned buffer owned by the ctx, so the call stays asynchronous (no stream synchronisation per
        // call); the buffer is reused only after the previous copy out of it has completed
        const size_t tbytes = table.size() * sizeof(double);
        if (ctx->h_coef_bytes < tbytes) {
            if (ctx->h_coef) {
                CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)ctx->ev_coef));
                cudaFreeHost(ctx->h_coef);
                ctx->h_coef = nullptr;
                ctx->h_coef_bytes = 0;
            }
            CUDA_TRY(ctx, cudaMallocHost(&ctx->h_coef, std::max<size_t>(tbytes, 4096)));
            ctx->h_coef_bytes = std::max<size_t>(tbytes, 4096);
        } else {
            CUDA_TRY(ctx, cudaEventSynchronize((cudaEvent_t)ctx->ev_coef));
        }
        std::memcpy(ctx->h_coef, table.data(), tbytes);
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_coef, ctx->h_coef, tbytes, cudaMemcpyHostToDevice, (cudaStream_t)ctx->stream));
        CUDA_TRY(ctx, cudaEventRecord((cudaEvent_t)ctx->ev_coef, (cudaStream_t)ctx->stream));
        cr.dev = ctx->d_coef;
        cr.cell_stride = len;
        cr.gp_stride = 0;
        cr.per_marker = true;
        return GL_OK;
    }
    case GL_COEF_PER_CELL:
    case GL_COEF_PER_GAUSS: {
        const bool per_gauss = coef->mode == GL_COEF_PER_GAUSS;
        uint64_t nrec = e - b;
        if (per_gauss) {
            // every kind in the range must use the same number of Gauss points
            int ng_common = -1;
            for (auto &bk : mesh->buckets) {
                uint64_t lo, hi;
                bucket_range(mesh, bk, b, e, lo, hi);
                if (lo >= hi) continue;
                int ng = 0, st;
                const double *data;
                st = gauss_rule(bk.kind, a->ngauss, &ng, &data);
                if (st) return fail(ctx, st);
                if (ng_common >= 0 && ng != ng_common)
                    return fail(ctx, GL_ERR_COEF, "PER_GAUSS needs one rule size for all kinds in the range");
                ng_common = ng;
            }
            nrec *= (uint64_t)std::max(ng_common, 1);
            cr.gp_stride = len;
            cr.cell_stride = 0; // filled per bucket as ng * len
        } else {
            cr.gp_stride = 0;
            cr.cell_stride = len;
        }
        if (coef->location == GL_DEVICE) {
            cr.dev = coef->data;
            // 3D only: there the pattern removes two thirds of the contraction and a record is 6 % of the element block; in 2D
            // the probe's extra pass over the records costs more than the general variant does (Qua4: 4.6 vs 6.5 G el/s)
            if (op == GL_OP_MAT_10_BDB && !per_gauss && nrec > 0 && mesh->ndim == 3) {
                const int ns = 2 * mesh->ndim;
                unsigned long long blocks = (nrec * (uint64_t)(ns * ns) + 255) / 256;
                if (blocks > 148ull * 16) blocks = 148ull * 16;
                modulus_pattern_reset_kernel<<<1, 1, 0, (cudaStream_t)ctx->stream>>>(ctx->d_coef_pat);
                modulus_pattern_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)ctx->stream>>>(coef->data, nrec, ns, ctx->d_coef_pat);
                CUDA_TRY(ctx, cudaGetLastError());
                cr.pat_dev = ctx->d_coef_pat;
            }
        } else {
            int st = ensure_buffer(ctx, &ctx->d_coef, &ctx->coef_bytes, nrec * len * sizeof(double));
            if (st) return st;
            const bool scan = op == GL_OP_MAT_10_BDB && !per_gauss; // structure common to all records -> kernel variant
            st = upload_records(ctx, coef->data, nrec * len, ctx->d_coef, scan ? 2 * mesh->ndim : 0, scan ? &cr.pat : nullptr);
            if (st) return st;
            cr.dev = ctx->d_coef;
        }
        return GL_OK;
    }
    default:
        return fail(ctx, GL_ERR_COEF, "unknown coefficient mode");
    }
}

} // extern "C"

// upper triangles of square column-major blocks, packed column-major: dst[cell * tri + j (j + 1) / 2 + i] = src[cell * n * n + i + j n],
// i <= j.  One warp per (cell, column): coalesced reads of the column head, contiguous writes.
__global__ void pack_upper_kernel(const double *__restrict__ src, double *__restrict__ dst, unsigned long long ncell, unsigned int n) {
    const unsigned long long tri = (unsigned long long)n * (n + 1) / 2;
    const unsigned int lane = threadIdx.x & 31;
    const unsigned long long warp = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const unsigned long long nwarp = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
    for (unsigned long long t = warp; t < ncell * n; t += nwarp) {
        const unsigned long long cell = t / n;
        const unsigned int j = (unsigned int)(t - cell * n);
        const double *s = src + cell * n * n + (unsigned long long)j * n;
        double *d = dst + cell * tri + (unsigned long long)j * (j + 1) / 2;
        for (unsigned int i = lane; i <= j; i += 32) d[i] = s[i];
    }
}

extern "C" {

// gl_args.packed: the kernels write full blocks into a bounded scratch slab, a pack pass keeps the upper triangles
static int integ_packed(gl_ctx *ctx, gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a, const CoefResolved &cr,
                        gl_out *out) {
    BlockDims bd;
    gl_args full = *a;
    full.packed = 0;
    int st = block_dims(op, mesh->buckets[0].kind, mesh->ndim, &full, bd);
    if (st) return fail(ctx, st);
    const uint64_t n = bd.size_r, sq = n * n, tri = n * (n + 1) / 2;
    const uint64_t per_chunk = std::max<uint64_t>(((uint64_t)192 << 20) / (sq * sizeof(double)), 1);
    const uint64_t chunk_cells = std::min<uint64_t>(per_chunk, e - b);
    st = ensure_buffer(ctx, &ctx->d_full, &ctx->full_bytes, chunk_cells * sq * sizeof(double));
    if (st) return st;
    cudaStream_t s_compute = (cudaStream_t)ctx->stream, s_copy = (cudaStream_t)ctx->copy_stream;
    auto pack = [&](uint64_t ncell, double *dst) {
        const unsigned long long warps = ncell * n;
        const unsigned grid = (unsigned)std::min<unsigned long long>((warps * 32 + 255) / 256, 148ull * 64);
        pack_upper_kernel<<<grid, 256, 0, s_compute>>>(ctx->d_full, dst, ncell, (unsigned)n);
        ctx->launches++;
    };
    if (out->location == GL_DEVICE) {
        for (uint64_t cb = b; cb < e; cb += per_chunk) {
            const uint64_t ce = std::min(e, cb + per_chunk);
            st = launch_range(ctx, mesh, op, cb, ce, &full, cr, b, ctx->d_full);
            if (st) return st;
            pack(ce - cb, out->ptr + (cb - b) * tri);
        }
        if (cudaGetLastError() != cudaSuccess) return cuda_fail(ctx, cudaGetLastError(), "pack kernel");
        return GL_OK;
    }
    const size_t stage = chunk_cells * tri * sizeof(double);
    if (ctx->stage_bytes < stage) {
        for (int i = 0; i < 2; i++) {
            if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
            ctx->d_stage[i] = nullptr;
        }
        ctx->stage_bytes = 0;
        for (int i = 0; i < 2; i++) CUDA_TRY(ctx, cudaMalloc(&ctx->d_stage[i], stage));
        ctx->stage_bytes = stage;
    }
    cudaEvent_t ev_kernel[2] = {(cudaEvent_t)ctx->ev[0], (cudaEvent_t)ctx->ev[1]};
    cudaEvent_t ev_copy[2] = {(cudaEvent_t)ctx->ev[2], (cudaEvent_t)ctx->ev[3]};
    // pageable destination: drained by the host through pinned bounce buffers, one chunk behind the kernels (as in gl_integ)
    const bool pageable = !host_pointer_is_pinned(out->ptr);
    auto drain = [&](size_t kk, uint64_t cb0, uint64_t ce0) -> int {
        const int buf = (int)(kk & 1);
        CUDA_TRY(ctx, cudaStreamWaitEvent(s_copy, ev_kernel[buf], 0));
        return drain_to_pageable(ctx, out->ptr + (cb0 - b) * tri, ctx->d_stage[buf], (ce0 - cb0) * tri, s_copy);
    };
    size_t k = 0;
    uint64_t prev_cb = 0, prev_ce = 0;
    for (uint64_t cb = b; cb < e; cb += per_chunk, k++) {
        const uint64_t ce = std::min(e, cb + per_chunk);
        const int buf = (int)(k & 1);
        if (k >= 2 && !pageable) CUDA_TRY(ctx, cudaStreamWaitEvent(s_compute, ev_copy[buf], 0));
        st = launch_range(ctx, mesh, op, cb, ce, &full, cr, b, ctx->d_full);
        if (st) {
            cudaStreamSynchronize(s_copy);
            cudaStreamSynchronize(s_compute);
            return st;
        }
        pack(ce - cb, ctx->d_stage[buf]);
        CUDA_TRY(ctx, cudaEventRecord(ev_kernel[buf], s_compute));
        if (!pageable) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(s_copy, ev_kernel[buf], 0));
            CUDA_TRY(ctx, cudaMemcpyAsync(out->ptr + (cb - b) * tri, ctx->d_stage[buf], (ce - cb) * tri * sizeof(double), cudaMemcpyDeviceToHost, s_copy));
            CUDA_TRY(ctx, cudaEventRecord(ev_copy[buf], s_copy));
        } else {
            if (k >= 1) {
                st = drain(k - 1, prev_cb, prev_ce);
                if (st) return st;
            }
            prev_cb = cb;
            prev_ce = ce;
        }
    }
    if (pageable && k >= 1) {
        st = drain(k - 1, prev_cb, prev_ce);
        if (st) return st;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(s_copy));
    return gl_ctx_sync(ctx);
}

int gl_integ(gl_ctx *ctx, const gl_mesh *cmesh, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef,
             gl_out *out) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    gl_mesh *mesh = const_cast<gl_mesh *>(cmesh);
    if (!mesh || !a || !out) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (op_size_class(op) < 0) return fail(ctx, GL_ERR_INVALID_ARG, "unknown op");
    if (mesh->device != ctx->device) return fail(ctx, GL_ERR_INVALID_ARG, "mesh belongs to another device");
    if (b > e || e > mesh->ncell) return fail(ctx, GL_ERR_CELL_RANGE);
    cudaSetDevice(ctx->device);
    uint64_t total = 0;
    int st = range_layout(mesh, op, b, e, a, &total, nullptr);
    if (st) return fail(ctx, st);
    if (e > b && (!out->ptr || out->capacity < total)) return fail(ctx, GL_ERR_OUT_CAPACITY);
    CoefResolved cr;
    st = resolve_coef(ctx, mesh, op, b, e, a, coef, cr);
    if (st) return st;
    if (b == e) return GL_OK;

    if (a->packed && op_is_matrix(op)) return integ_packed(ctx, mesh, op, b, e, a, cr, out);
    if (out->location == GL_DEVICE) return launch_range(ctx, mesh, op, b, e, a, cr, b, out->ptr);

    // ---- host output: pipeline chunks through two device staging buffers --------------------------------
    const size_t chunk_target = (size_t)128 << 20;
    const size_t avg_block = (size_t)std::max<uint64_t>(total / (e - b), 1) * sizeof(double);
    uint64_t cells_per_chunk = std::max<uint64_t>(chunk_target / avg_block, 1);
    // staging must hold the largest chunk; blocks may be unequal on mixed meshes, so size by the real layout
    std::vector<uint64_t> cut{b};
    while (cut.back() < e) cut.push_back(std::min(e, cut.back() + cells_per_chunk));
    size_t max_bytes = 0;
    std::vector<uint64_t> chunk_total(cut.size() - 1), chunk_off(cut.size() - 1);
    uint64_t run = 0;
    for (size_t k = 0; k + 1 < cut.size(); k++) {
        uint64_t t = 0;
        st = range_layout(mesh, op, cut[k], cut[k + 1], a, &t, nullptr);
        if (st) return fail(ctx, st);
        chunk_total[k] = t;
        chunk_off[k] = run;
        run += t;
        max_bytes = std::max(max_bytes, (size_t)t * sizeof(double));
    }
    if (ctx->stage_bytes < max_bytes) {
        for (int i = 0; i < 2; i++) {
            if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
            ctx->d_stage[i] = nullptr;
        }
        ctx->stage_bytes = 0;
        for (int i = 0; i < 2; i++) CUDA_TRY(ctx, cudaMalloc(&ctx->d_stage[i], max_bytes));
        ctx->stage_bytes = max_bytes;
    }
    cudaStream_t s_compute = (cudaStream_t)ctx->stream, s_copy = (cudaStream_t)ctx->copy_stream;
    cudaEvent_t ev_kernel[2] = {(cudaEvent_t)ctx->ev[0], (cudaEvent_t)ctx->ev[1]};
    cudaEvent_t ev_copy[2] = {(cudaEvent_t)ctx->ev[2], (cudaEvent_t)ctx->ev[3]};
    // pageable destination: chunks are drained by the host through pinned bounce buffers (drain_to_pageable), one chunk behind the kernels
    const bool pageable = !host_pointer_is_pinned(out->ptr);
    auto drain = [&](size_t k) -> int {
        const int buf = (int)(k & 1);
        CUDA_TRY(ctx, cudaStreamWaitEvent(s_copy, ev_kernel[buf], 0));
        return drain_to_pageable(ctx, out->ptr + chunk_off[k], ctx->d_stage[buf], chunk_total[k], s_copy);
    };
    for (size_t k = 0; k + 1 < cut.size(); k++) {
        const int buf = (int)(k & 1);
        if (k >= 2 && !pageable) CUDA_TRY(ctx, cudaStreamWaitEvent(s_compute, ev_copy[buf], 0)); // (pageable: chunk k - 2 was drained before this point)
        if (!a->clear)
            CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[buf], out->ptr + chunk_off[k], chunk_total[k] * sizeof(double),
                                          cudaMemcpyHostToDevice, s_compute));
        st = launch_range(ctx, mesh, op, cut[k], cut[k + 1], a, cr, b, ctx->d_stage[buf]);
        if (st) {
            cudaStreamSynchronize(s_copy);
            cudaStreamSynchronize(s_compute);
            return st;
        }
        CUDA_TRY(ctx, cudaEventRecord(ev_kernel[buf], s_compute));
        if (!pageable) {
            CUDA_TRY(ctx, cudaStreamWaitEvent(s_copy, ev_kernel[buf], 0));
            CUDA_TRY(ctx, cudaMemcpyAsync(out->ptr + chunk_off[k], ctx->d_stage[buf], chunk_total[k] * sizeof(double),
                                          cudaMemcpyDeviceToHost, s_copy));
            CUDA_TRY(ctx, cudaEventRecord(ev_copy[buf], s_copy));
        } else {
            // the kernels of chunk k are enqueued: drain chunk k - 1 on the host while they run
            if (k >= 1) {
                st = drain(k - 1);
                if (st) return st;
            }
        }
    }
    if (pageable && cut.size() >= 2) {
        st = drain(cut.size() - 2);
        if (st) return st;
    }
    CUDA_TRY(ctx, cudaStreamSynchronize(s_copy));
    return gl_ctx_sync(ctx);
}

// ---- fused entry point ---------------------------------------------------------------------------------------------------
static bool scalar_family(int op) {
    return op == GL_OP_MAT_01_NSN || op == GL_OP_MAT_03_BTB || op == GL_OP_VEC_01_NS || op == GL_OP_VEC_03_BV;
}

int gl_integ_fused(gl_ctx *ctx, const gl_mesh *cmesh, uint64_t b, uint64_t e, const gl_args *a, const gl_fused_item *items, int nitems) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    gl_mesh *mesh = const_cast<gl_mesh *>(cmesh);
    if (!mesh || !a || !items || nitems <= 0) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (b > e || e > mesh->ncell) return fail(ctx, GL_ERR_CELL_RANGE);
    KernelCollector collector(ctx);
    if (a->packed) { // packed matrices go through the pack pass of gl_integ
        for (int i = 0; i < nitems; i++) {
            int st = gl_integ(ctx, cmesh, items[i].op, b, e, a, items[i].coef, items[i].out);
            if (st) return st;
        }
        return GL_OK;
    }
    // stiffness + internal-force vector (the two sides of a Newton step) from ONE geometry pass: Hex8 meshes, sigma per Gauss
    // point; any other combination of these two cases runs them one after the other
    if (nitems == 2 && e > b) {
        const gl_fused_item *ik = nullptr, *id = nullptr;
        for (int i = 0; i < 2; i++) {
            if (items[i].op == GL_OP_MAT_10_BDB) ik = &items[i];
            if (items[i].op == GL_OP_VEC_04_BT) id = &items[i];
        }
        int ng_rule = 0;
        const double *rule_data = nullptr;
        // the fused kernel is the 8-point one: with any other rule the stress buffer has another size (ng records per cell)
        const bool rule8 = gauss_rule(GL_HEX8, a->ngauss, &ng_rule, &rule_data) == GL_OK && ng_rule == 8;
        if (ik && id && rule8 && ik->coef && id->coef && ik->out && id->out && mesh->homogeneous && mesh->buckets.size() == 1 &&
            mesh->buckets[0].kind == GL_HEX8 && a->clear && !a->axisymmetric && a->ii0 == 0 && a->jj0 == 0 && a->nrow == 0 &&
            a->ncol == 0 && ik->coef->mode != GL_COEF_PER_GAUSS && id->coef->mode == GL_COEF_PER_GAUSS && id->coef->data &&
            ik->out->location == GL_DEVICE && id->out->location == GL_DEVICE && ik->out->ptr && id->out->ptr) {
            if (ik->out->capacity < (e - b) * 576 || id->out->capacity < (e - b) * 24) return fail(ctx, GL_ERR_OUT_CAPACITY);
            cudaSetDevice(ctx->device);
            CoefResolved crk;
            int st = resolve_coef(ctx, mesh, GL_OP_MAT_10_BDB, b, e, a, ik->coef, crk);
            if (st) return st;
            const double *d_sigma = id->coef->data;
            double *tmp = nullptr;
            const size_t sig_bytes = (size_t)(e - b) * 8 * 6 * sizeof(double);
            if (id->coef->location != GL_DEVICE) {
                CUDA_TRY(ctx, pool_alloc_t(ctx, &tmp, sig_bytes));
                int stu = upload_records(ctx, id->coef->data, sig_bytes / sizeof(double), tmp, 0, nullptr);
                if (stu) return stu;
                d_sigma = tmp;
            }
            bool done = false;
            st = launch_range(ctx, mesh, GL_OP_MAT_10_BDB, b, e, a, crk, b, ik->out->ptr, d_sigma, id->out->ptr, &done);
            if (tmp) {
                cudaStreamSynchronize((cudaStream_t)ctx->stream);
                pool_free(ctx, tmp, sig_bytes);
            }
            if (st) return st;
            if (!done) return gl_integ(ctx, cmesh, GL_OP_VEC_04_BT, b, e, a, id->coef, id->out); // e.g. another rule than 2x2x2
            return GL_OK;
        }
    }
    // stiffness + scalar-dof cases in one call (BASELINE.json configs[4]: mat_10_bdb + mat_03_btb + vec_03_bv): the stiffness runs on
    // its own kernels and the scalar cases share the fused scalar kernel -- except on Qua8 buckets, which take mat_03_btb / vec_03_bv
    // from the geometric tensors of the tensor-core stiffness kernel (gl_kernels_qua8.cu)
    const gl_fused_item *const all_items = items;
    const int all_nitems = nitems;
    const gl_fused_item *ik = nullptr;
    gl_fused_item sitems[4];
    {
        int nk = 0, ns = 0;
        bool ok = nitems >= 2 && nitems <= 5 && e > b;
        for (int i = 0; i < nitems && ok; i++) {
            if (items[i].op == GL_OP_MAT_10_BDB) {
                nk++;
                ik = &items[i];
            } else if (scalar_family(items[i].op) && ns < 4) {
                sitems[ns++] = items[i];
            } else {
                ok = false;
            }
        }
        ok = ok && nk == 1 && ns >= 1 && ik->coef && ik->out && ik->coef->mode == GL_COEF_UNIFORM && ik->coef->data &&
             ik->out->location == GL_DEVICE && ik->out->ptr;
        if (ok) {
            items = sitems;
            nitems = ns;
        } else {
            ik = nullptr;
        }
    }
    // can the items share one kernel?
    bool fusable = nitems <= 4 && a->clear && a->ii0 == 0 && a->jj0 == 0 && a->nrow == 0 && a->ncol == 0 && e > b;
    int mask = 0;
    for (int i = 0; i < nitems && fusable; i++) {
        const gl_fused_item &it = items[i];
        if (!scalar_family(it.op) || !it.coef || !it.out || it.coef->mode != GL_COEF_UNIFORM || it.coef->location != GL_HOST ||
            !it.coef->data || it.out->location != GL_DEVICE || !it.out->ptr)
            fusable = false;
        const int bit = it.op == GL_OP_MAT_01_NSN ? 1 : (it.op == GL_OP_MAT_03_BTB ? 2 : (it.op == GL_OP_VEC_01_NS ? 4 : 8));
        if (mask & bit) fusable = false; // the same case twice
        mask |= bit;
    }
    for (auto &bk : mesh->buckets)
        if (kind_geo_ndim(bk.kind) != mesh->ndim) fusable = false;
    if (!fusable) {
        for (int i = 0; i < all_nitems; i++) {
            int st = gl_integ(ctx, cmesh, all_items[i].op, b, e, a, all_items[i].coef, all_items[i].out);
            if (st) return st;
        }
        return GL_OK;
    }
    cudaSetDevice(ctx->device);
    ScalarFused f;
    std::memset(&f, 0, sizeof(f));
    f.mask = mask;
    int mat_op = -1, vec_op = -1;
    for (int i = 0; i < nitems; i++) {
        const gl_fused_item &it = items[i];
        uint64_t total = 0;
        int st = range_layout(mesh, it.op, b, e, a, &total, nullptr);
        if (st) return fail(ctx, st);
        if (it.out->capacity < total) return fail(ctx, GL_ERR_OUT_CAPACITY);
        switch (it.op) {
        case GL_OP_MAT_01_NSN: f.s_mat = it.coef->data[0]; f.out_m01 = it.out->ptr; mat_op = it.op; break;
        case GL_OP_MAT_03_BTB: std::memcpy(f.t, it.coef->data, sizeof(double) * 2 * mesh->ndim); f.out_m03 = it.out->ptr; mat_op = it.op; break;
        case GL_OP_VEC_01_NS: f.s_vec = it.coef->data[0]; f.out_v01 = it.out->ptr; vec_op = it.op; break;
        default: std::memcpy(f.w, it.coef->data, sizeof(double) * mesh->ndim); f.out_v03 = it.out->ptr; vec_op = it.op; break;
        }
    }
    const std::vector<uint64_t> *mpre = nullptr, *vpre = nullptr;
    BlockKey mkey, vkey;
    int st = GL_OK;
    if (!mesh->homogeneous) {
        if (mat_op >= 0) mpre = mesh_prefix(mesh, mat_op, a, &mkey, &st);
        if (vec_op >= 0) vpre = mesh_prefix(mesh, vec_op, a, &vkey, &st);
    }
    std::vector<char> done(mesh->buckets.size(), 0);
    if (ik) {
        uint64_t total = 0;
        st = range_layout(mesh, GL_OP_MAT_10_BDB, b, e, a, &total, nullptr);
        if (st) return fail(ctx, st);
        if (ik->out->capacity < total) return fail(ctx, GL_ERR_OUT_CAPACITY);
        CoefResolved crk;
        st = resolve_coef(ctx, mesh, GL_OP_MAT_10_BDB, b, e, a, ik->coef, crk);
        if (st) return st;
        Qua8Fusion qf{&f, mpre, vpre, mkey, vkey, &done};
        st = launch_range(ctx, mesh, GL_OP_MAT_10_BDB, b, e, a, crk, b, ik->out->ptr, nullptr, nullptr, nullptr, nullptr, nullptr, -1, &qf);
        if (st) return st;
    }
    for (size_t bi = 0; bi < mesh->buckets.size(); bi++) {
        if (done[bi]) continue; // served by the stiffness kernel
        auto &bk = mesh->buckets[bi];
        uint64_t lo, hi;
        bucket_range(mesh, bk, b, e, lo, hi);
        if (lo >= hi) continue;
        IntegParams p;
        std::memset(&p, 0, sizeof(p));
        p.coords = mesh->d_coords;
        p.npoint = mesh->npoint;
        if ((mask & ~5) == 0 || (mask & ~10) == 0) { // thread-per-cell kernels gather whole points
            int stp = ensure_point_major(ctx, mesh);
            if (stp) return stp;
            p.coords_pm = mesh->d_coords_pm;
        }
        p.conn = bk.d_conn;
        p.ncell_k = bk.ncell;
        p.cell_ids = bk.d_cell_ids;
        p.lo = lo;
        p.hi = hi;
        p.cell_begin = b;
        p.coef_begin = b;
        p.axisym = a->axisymmetric ? 1 : 0;
        p.clear = 1;
        p.alpha = a->alpha;
        p.nnode = bk.nnode;
        RuleTable *rt = get_rule(ctx, bk.kind, a->ngauss, &st);
        if (!rt) return fail(ctx, st);
        p.rule = rt->dev;
        p.ng = rt->ng;
        p.err_cell = ctx->d_err_cell;
        ScalarFused fb = f;
        if (mpre) {
            st = bucket_device_offsets(ctx, bk, mpre, mkey, &fb.moff);
            if (st) return st;
            fb.mbase = (*mpre)[b];
        }
        if (vpre) {
            st = bucket_device_offsets(ctx, bk, vpre, vkey, &fb.voff);
            if (st) return st;
            fb.vbase = (*vpre)[b];
        }
        int nl = launch_mass_dmma(p, fb, mesh->ndim, ctx->stream);
        if (nl == 0) nl = launch_scalar_tpc(p, fb, mesh->ndim, ctx->stream);
        if (nl == 0) nl = launch_scalar_fused(p, fb, mesh->ndim, ctx->stream);
        if (nl < 0) return cuda_fail(ctx, cudaGetLastError(), "kernel launch");
        if (nl > 0) set_last_kernel(ctx);
        if (nl == 0) { // kind not covered by the fused kernel: run the items one after the other for the whole range
            for (int i = 0; i < nitems; i++) {
                int st2 = gl_integ(ctx, cmesh, items[i].op, b, e, a, items[i].coef, items[i].out);
                if (st2) return st2;
            }
            return GL_OK;
        }
        ctx->launches += (uint64_t)nl;
    }
    return GL_OK;
}

int gl_mat_10_bdb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_10_BDB, b, e, a, c, o);
}
int gl_mat_01_nsn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_01_NSN, b, e, a, c, o);
}
int gl_mat_03_btb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_03_BTB, b, e, a, c, o);
}
int gl_mat_02_bvn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_02_BVN, b, e, a, c, o);
}
int gl_mat_08_ntn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_08_NTN, b, e, a, c, o);
}
int gl_mat_09_nvb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_09_NVB, b, e, a, c, o);
}
int gl_vec_01_ns(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_01_NS, b, e, a, c, o);
}
int gl_vec_02_nv(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_02_NV, b, e, a, c, o);
}
int gl_vec_03_bv(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_03_BV, b, e, a, c, o);
}
int gl_vec_04_bt(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_04_BT, b, e, a, c, o);
}
int gl_mat_04_nsb(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_04_NSB, b, e, a, c, o);
}
int gl_mat_05_btn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_05_BTN, b, e, a, c, o);
}
int gl_mat_06_nvn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_06_NVN, b, e, a, c, o);
}
int gl_mat_07_bsn(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_07_BSN, b, e, a, c, o);
}
int gl_mat_01_nsn_bry(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_01_NSN_BRY, b, e, a, c, o);
}
int gl_vec_01_ns_bry(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_01_NS_BRY, b, e, a, c, o);
}
int gl_vec_02_nv_bry(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_02_NV_BRY, b, e, a, c, o);
}
int gl_normal_vectors(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_NORMAL, b, e, a, nullptr, o);
}
int gl_jacobian(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_JACOBIAN, b, e, a, nullptr, o);
}
int gl_mat_10_gdg(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_MAT_10_BDB, b, e, a, c, o);
}
int gl_vec_10_gs(gl_ctx *ctx, const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *c, gl_out *o) {
    return gl_integ(ctx, mesh, GL_OP_VEC_03_BV, b, e, a, c, o);
}

int gl_bench_fp64_peak(gl_ctx *ctx, int kind, int iters, double *tflops) {
    if (!ctx || !tflops) return GL_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    int st = bench_fp64(kind, iters, ctx->stream, tflops);
    if (st < 0) return cuda_fail(ctx, cudaGetLastError(), "fp64 micro-benchmark");
    ctx->launches += (uint64_t)st;
    return GL_OK;
}

int gl_bench_hbm_write(gl_ctx *ctx, int kind, double *dev, uint64_t nbytes, int warps_per_sm, double *gbs) {
    if (!ctx || !dev || !gbs || nbytes < 4608) return GL_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    int st = bench_hbm_write(kind, dev, nbytes, warps_per_sm, ctx->stream, gbs);
    if (st < 0) return cuda_fail(ctx, cudaGetLastError(), "hbm write micro-benchmark");
    ctx->launches += (uint64_t)st;
    return GL_OK;
}

} // extern "C"

// =====================================================================================================
// assembly hand-off: element blocks -> global sparse matrix / vector, on the device
// =====================================================================================================
// The reference stops at the element matrix; its callers (pmsim-style assemblers) map local dofs to global equations with
// the rule of src/integ/mat_10_bdb.rs:31-46 (local row ii = i + m*ndim belongs to dof i of node m) and add every entry to
// a sparse matrix.  With a point -> equation table that step needs no host work at all.
namespace {

struct ScatterParams {
    const uint32_t *conn;             // [nnode][ncell_k]
    unsigned long long ncell_k;
    const uint32_t *cell_ids;         // bucket-local -> original id, or nullptr
    unsigned long long lo, hi;        // bucket-local range
    unsigned long long cell_begin;    // original id of the cell whose block starts at offset 0
    const unsigned long long *out_off; // mixed meshes: bucket-local cell -> absolute block offset
    unsigned long long out_base;
    int nnode, npd;                   // dofs per node of this op (ndim or 1)
    int nr, nc;                       // block dims (nc = 1 for vectors)
    const int32_t *eq;                // [npoint][npd] equation numbers, negative = not assembled
};

__device__ __forceinline__ void scatter_entry(const ScatterParams &s, unsigned long long idx, unsigned long long &o, int &r, int &c) {
    const unsigned int bs = (unsigned int)s.nr * (unsigned int)s.nc;
    const unsigned long long cl = s.lo + idx / bs;
    const unsigned int q = (unsigned int)(idx % bs);
    const unsigned int jj = q / (unsigned int)s.nr, ii = q - jj * (unsigned int)s.nr;
    const unsigned long long orig = s.cell_ids ? (unsigned long long)s.cell_ids[cl] : cl;
    o = (s.out_off ? (s.out_off[cl] - s.out_base) : (orig - s.cell_begin) * bs) + q;
    const unsigned int pr = s.conn[(unsigned long long)(ii / s.npd) * s.ncell_k + cl];
    r = s.eq[(unsigned long long)pr * s.npd + ii % s.npd];
    c = 0;
    if (s.nc > 1) {
        const unsigned int pc = s.conn[(unsigned long long)(jj / s.npd) * s.ncell_k + cl];
        c = s.eq[(unsigned long long)pc * s.npd + jj % s.npd];
    }
}

__global__ void triplet_kernel(const ScatterParams s, int32_t *rows, int32_t *cols) {
    const unsigned long long total = (s.hi - s.lo) * (unsigned long long)(s.nr * s.nc);
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long o;
        int r, c;
        scatter_entry(s, idx, o, r, c);
        rows[o] = r;
        if (cols) cols[o] = c;
    }
}

__global__ void csr_scatter_kernel(const ScatterParams s, const double *blocks, const long long *rowptr, const int32_t *colidx,
                                   double *vals, unsigned long long *missing) {
    const unsigned long long total = (s.hi - s.lo) * (unsigned long long)(s.nr * s.nc);
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long o;
        int r, c;
        scatter_entry(s, idx, o, r, c);
        if (r < 0 || c < 0) continue;
        long long a = rowptr[r], b = rowptr[r + 1];
        while (a < b) { // lower bound of c in the (ascending) column indices of row r
            const long long mid = (a + b) >> 1;
            if (colidx[mid] < c) a = mid + 1;
            else b = mid;
        }
        if (a < rowptr[r + 1] && colidx[a] == c) atomicAdd(vals + a, blocks[o]);
        else atomicAdd(missing, 1ull);
    }
}

__global__ void vec_scatter_kernel(const ScatterParams s, const double *blocks, double *rhs) {
    const unsigned long long total = (s.hi - s.lo) * (unsigned long long)s.nr;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long o;
        int r, c;
        scatter_entry(s, idx, o, r, c);
        if (r >= 0) atomicAdd(rhs + r, blocks[o]);
    }
}

// fills the scatter description of one bucket for original cells [b, e) whose blocks start at offset 0 with cell b
int scatter_params(gl_ctx *ctx, gl_mesh *mesh, gl_mesh_bucket &bk, int op, const gl_args *a, uint64_t b, uint64_t e,
                   const int32_t *d_eq, ScatterParams &s, bool &empty) {
    uint64_t lo, hi;
    bucket_range(mesh, bk, b, e, lo, hi);
    empty = lo >= hi;
    if (empty) return GL_OK;
    unsigned r, c;
    op_extent(op, bk.kind, mesh->ndim, a, r, c);
    s.conn = bk.d_conn;
    s.ncell_k = bk.ncell;
    s.cell_ids = bk.d_cell_ids;
    s.lo = lo;
    s.hi = hi;
    s.cell_begin = b;
    s.out_off = nullptr;
    s.out_base = 0;
    s.nnode = bk.nnode;
    s.npd = (int)r / bk.nnode;
    s.nr = (int)r;
    s.nc = op_is_matrix(op) ? (int)c : 1;
    s.eq = d_eq;
    if (!mesh->homogeneous) {
        BlockKey key;
        int st = 0;
        const std::vector<uint64_t> *pre = mesh_prefix(mesh, op, a, &key, &st);
        if (st) return fail(ctx, st);
        auto it = bk.d_out_off.find(key);
        if (it == bk.d_out_off.end()) {
            std::vector<unsigned long long> offs(bk.ncell);
            for (uint64_t k = 0; k < bk.ncell; k++) offs[k] = (*pre)[bk.h_cell_ids[k]];
            unsigned long long *d = nullptr;
            CUDA_TRY(ctx, cudaMalloc(&d, offs.size() * sizeof(unsigned long long)));
            CUDA_TRY(ctx, cudaMemcpy(d, offs.data(), offs.size() * sizeof(unsigned long long), cudaMemcpyHostToDevice));
            it = bk.d_out_off.emplace(key, d).first;
        }
        s.out_off = it->second;
        s.out_base = (*pre)[b];
    }
    return GL_OK;
}

unsigned scatter_grid(const ScatterParams &s) {
    const unsigned long long total = (s.hi - s.lo) * (unsigned long long)(s.nr * s.nc);
    return (unsigned)std::min<unsigned long long>((total + 255) / 256, 148ull * 32);
}

// the equation table on the device (copied if the caller passed a host pointer); *owned tells the caller to free it
int eq_on_device(gl_ctx *ctx, const gl_mesh *mesh, int npd, const int32_t *eq, int location, const int32_t **d_eq, bool *owned) {
    *owned = false;
    if (location == GL_DEVICE) {
        *d_eq = eq;
        return GL_OK;
    }
    int32_t *d = nullptr;
    const size_t bytes = (size_t)mesh->npoint * npd * sizeof(int32_t);
    CUDA_TRY(ctx, cudaMalloc(&d, bytes));
    CUDA_TRY(ctx, cudaMemcpyAsync(d, eq, bytes, cudaMemcpyHostToDevice, (cudaStream_t)ctx->stream));
    *d_eq = d;
    *owned = true;
    return GL_OK;
}

int assemble_checks(gl_ctx *ctx, gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a) {
    if (!mesh || !a) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (op_size_class(op) < 0 || op == GL_OP_JACOBIAN || op_two_pads(op) || op == GL_OP_NORMAL)
        return fail(ctx, GL_ERR_INVALID_ARG, "op cannot be assembled");
    if (mesh->device != ctx->device) return fail(ctx, GL_ERR_INVALID_ARG, "mesh belongs to another device");
    if (b > e || e > mesh->ncell) return fail(ctx, GL_ERR_CELL_RANGE);
    if (a->ii0 || a->jj0 || a->nrow || a->ncol || a->packed) return fail(ctx, GL_ERR_INVALID_ARG, "assembly needs tight, full element blocks (ii0 = jj0 = 0, packed = 0)");
    return GL_OK;
}

int op_dofs_per_node(int op, int ndim) {
    return (op == GL_OP_MAT_10_BDB || op == GL_OP_VEC_02_NV || op == GL_OP_VEC_04_BT || op == GL_OP_MAT_08_NTN || op == GL_OP_MAT_09_NVB ||
            op == GL_OP_VEC_02_NV_BRY)
               ? ndim
               : 1;
}

} // namespace

extern "C" {

int gl_op_dofs_per_node(int op, int ndim) { return op_dofs_per_node(op, ndim); }

int gl_triplet_indices(gl_ctx *ctx, const gl_mesh *cmesh, int op, uint64_t b, uint64_t e, const gl_args *a, const int32_t *eq,
                       int eq_location, int32_t *rows, int32_t *cols) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    gl_mesh *mesh = const_cast<gl_mesh *>(cmesh);
    int st = assemble_checks(ctx, mesh, op, b, e, a);
    if (st) return st;
    if (!eq || !rows || (op_is_matrix(op) && !cols)) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (b == e) return GL_OK;
    cudaSetDevice(ctx->device);
    const int32_t *d_eq = nullptr;
    bool owned = false;
    st = eq_on_device(ctx, mesh, op_dofs_per_node(op, mesh->ndim), eq, eq_location, &d_eq, &owned);
    if (st) return st;
    for (auto &bk : mesh->buckets) {
        ScatterParams s;
        bool empty;
        st = scatter_params(ctx, mesh, bk, op, a, b, e, d_eq, s, empty);
        if (st) break;
        if (empty) continue;
        triplet_kernel<<<scatter_grid(s), 256, 0, (cudaStream_t)ctx->stream>>>(s, rows, op_is_matrix(op) ? cols : nullptr);
        ctx->launches++;
    }
    if (owned) {
        cudaStreamSynchronize((cudaStream_t)ctx->stream);
        cudaFree(const_cast<int32_t *>(d_eq));
    }
    if (st) return st;
    if (cudaGetLastError() != cudaSuccess) return cuda_fail(ctx, cudaGetLastError(), "triplet kernel");
    return GL_OK;
}

int gl_assemble(gl_ctx *ctx, const gl_mesh *cmesh, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef,
                const int32_t *eq, int eq_location, const int64_t *rowptr, const int32_t *colidx, double *vals, uint64_t *missing) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    gl_mesh *mesh = const_cast<gl_mesh *>(cmesh);
    int st = assemble_checks(ctx, mesh, op, b, e, a);
    if (st) return st;
    const bool is_mat = op_is_matrix(op);
    if (!eq || !vals || (is_mat && (!rowptr || !colidx))) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (missing) *missing = 0;
    if (b == e) return GL_OK;
    cudaSetDevice(ctx->device);
    gl_args args = *a;
    args.clear = 1; // the element blocks are scratch; accumulation happens in the global matrix / vector
    CoefResolved cr;
    st = resolve_coef(ctx, mesh, op, b, e, &args, coef, cr);
    if (st) return st;
    const int32_t *d_eq = nullptr;
    bool owned = false;
    st = eq_on_device(ctx, mesh, op_dofs_per_node(op, mesh->ndim), eq, eq_location, &d_eq, &owned);
    if (st) return st;
    unsigned long long *d_missing = nullptr;
    if (cudaMalloc(&d_missing, sizeof(unsigned long long)) != cudaSuccess) return cuda_fail(ctx, cudaGetLastError(), "cudaMalloc");
    cudaMemsetAsync(d_missing, 0, sizeof(unsigned long long), (cudaStream_t)ctx->stream);

    // chunks of cells through one scratch slab (<= 256 MB): integrate, then scatter
    uint64_t total = 0;
    st = range_layout(mesh, op, b, e, &args, &total, nullptr);
    const size_t avg = (size_t)std::max<uint64_t>(total / (e - b), 1) * sizeof(double);
    const uint64_t per_chunk = std::max<uint64_t>(((size_t)256 << 20) / avg, 1);
    for (uint64_t cb = b; cb < e && !st; cb += per_chunk) {
        const uint64_t ce = std::min(e, cb + per_chunk);
        uint64_t t = 0;
        st = range_layout(mesh, op, cb, ce, &args, &t, nullptr);
        if (st) { st = fail(ctx, st); break; }
        if (ctx->stage_bytes < t * sizeof(double)) {
            for (int i = 0; i < 2; i++) {
                if (ctx->d_stage[i]) cudaFree(ctx->d_stage[i]);
                ctx->d_stage[i] = nullptr;
            }
            ctx->stage_bytes = 0;
            bool ok = true;
            for (int i = 0; i < 2; i++) ok = ok && cudaMalloc(&ctx->d_stage[i], t * sizeof(double)) == cudaSuccess;
            if (!ok) { st = cuda_fail(ctx, cudaGetLastError(), "cudaMalloc"); break; }
            ctx->stage_bytes = t * sizeof(double);
        }
        st = launch_range(ctx, mesh, op, cb, ce, &args, cr, b, ctx->d_stage[0]);
        if (st) break;
        for (auto &bk : mesh->buckets) {
            ScatterParams s;
            bool empty;
            st = scatter_params(ctx, mesh, bk, op, &args, cb, ce, d_eq, s, empty);
            if (st) break;
            if (empty) continue;
            if (is_mat)
                csr_scatter_kernel<<<scatter_grid(s), 256, 0, (cudaStream_t)ctx->stream>>>(s, ctx->d_stage[0], (const long long *)rowptr,
                                                                                           colidx, vals, d_missing);
            else
                vec_scatter_kernel<<<scatter_grid(s), 256, 0, (cudaStream_t)ctx->stream>>>(s, ctx->d_stage[0], vals);
            ctx->launches++;
        }
    }
    unsigned long long h_missing = 0;
    cudaMemcpyAsync(&h_missing, d_missing, sizeof(h_missing), cudaMemcpyDeviceToHost, (cudaStream_t)ctx->stream);
    const int sync = gl_ctx_sync(ctx);
    cudaFree(d_missing);
    if (owned) cudaFree(const_cast<int32_t *>(d_eq));
    if (missing) *missing = h_missing;
    if (st) return st;
    if (sync) return sync;
    if (h_missing) return fail(ctx, GL_ERR_INVALID_ARG, "sparsity pattern lacks entries of the element blocks");
    return GL_OK;
}

} // extern "C"

// =====================================================================================================
// planned assembly: the position of every element-block entry in the global matrix is computed ONCE; after that
// integrate-and-assemble is one kernel per GeoKind for the stiffness matrix (the blocks go from shared memory straight into the
// CSR values, RED.ADD.F64 -- the element-matrix slab is never materialised) and slab + position-map scatter for the other cases
// =====================================================================================================
struct gl_asm_plan {
    gl_mesh *mesh = nullptr;
    int op = 0;
    uint64_t b = 0, e = 0, total = 0;
    gl_args args{};
    bool is_mat = true;
    uint32_t *d_pos = nullptr; // [total] position in vals of entry (row, col) of every block, ROW-major inside the block; 0xFFFFFFFF = skip
};

namespace {

constexpr uint32_t NOPOS = 0xFFFFFFFFu;

__global__ void plan_positions_kernel(const ScatterParams s, const long long *rowptr, const int32_t *colidx, uint32_t *pos,
                                      unsigned long long *missing, const int upper_only) {
    const unsigned int bs = (unsigned int)s.nr * (unsigned int)s.nc;
    const unsigned long long total = (s.hi - s.lo) * (unsigned long long)bs;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long cl = s.lo + idx / bs;
        const unsigned int q = (unsigned int)(idx % bs);
        const unsigned int ii = q / (unsigned int)s.nc, jj = q - ii * (unsigned int)s.nc; // row-major inside the block
        const unsigned long long orig = s.cell_ids ? (unsigned long long)s.cell_ids[cl] : cl;
        const unsigned long long o = (s.out_off ? (s.out_off[cl] - s.out_base) : (orig - s.cell_begin) * bs) + q;
        const unsigned int pr = s.conn[(unsigned long long)(ii / s.npd) * s.ncell_k + cl];
        const int r = s.eq[(unsigned long long)pr * s.npd + ii % s.npd];
        uint32_t u = NOPOS;
        if (s.nc == 1) {
            if (r >= 0) u = (uint32_t)r;
        } else {
            const unsigned int pc = s.conn[(unsigned long long)(jj / s.npd) * s.ncell_k + cl];
            const int c = s.eq[(unsigned long long)pc * s.npd + jj % s.npd];
            if (r >= 0 && c >= 0 && !(upper_only && r > c)) { // symmetric storage: the strictly lower entries are not assembled
                long long a = rowptr[r], b = rowptr[r + 1];
                const long long end = b;
                while (a < b) {
                    const long long mid = (a + b) >> 1;
                    if (colidx[mid] < c) a = mid + 1;
                    else b = mid;
                }
                if (a < end && colidx[a] == c && a < (long long)NOPOS) u = (uint32_t)a;
                else atomicAdd(missing, 1ull);
            }
        }
        pos[o] = u;
    }
}

// slab path: blocks are column-major in `blocks`, the position map is row-major inside the block
__global__ void planned_scatter_kernel(const ScatterParams s, const double *blocks, const uint32_t *pos, unsigned long long pos_base, double *vals) {
    const unsigned int bs = (unsigned int)s.nr * (unsigned int)s.nc;
    const unsigned long long total = (s.hi - s.lo) * (unsigned long long)bs;
    for (unsigned long long idx = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; idx < total;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long cl = s.lo + idx / bs;
        const unsigned int q = (unsigned int)(idx % bs);
        const unsigned int ii = q / (unsigned int)s.nc, jj = q - ii * (unsigned int)s.nc;
        const unsigned long long orig = s.cell_ids ? (unsigned long long)s.cell_ids[cl] : cl;
        const unsigned long long base = s.out_off ? (s.out_off[cl] - s.out_base) : (orig - s.cell_begin) * bs;
        const uint32_t u = pos[pos_base + base + q];
        if (u != NOPOS) atomicAdd(vals + u, blocks[base + ii + (unsigned long long)jj * s.nr]);
    }
}

// can the stiffness of this bucket go through a kernel with a scatter epilogue (gl_kernels_hex8.cu / gl_kernels_bdb.cu)?
bool stiffness_fusable(const gl_mesh *mesh, const gl_mesh_bucket &bk, int op, const gl_args *a, const gl_coef *coef) {
    if (op != GL_OP_MAT_10_BDB || kind_geo_ndim(bk.kind) != mesh->ndim) return false;
    if (a->axisymmetric && mesh->ndim != 2) return false;
    int ng = 0;
    const double *data = nullptr;
    if (gauss_rule(bk.kind, a->ngauss, &ng, &data)) return false;
    const bool pg = coef->mode == GL_COEF_PER_GAUSS;
    const bool drec = !pg && coef->mode != GL_COEF_UNIFORM;
    return bdb_covers(bk.nnode, mesh->ndim, ng, a->axisymmetric != 0, drec, pg);
}

} // namespace

extern "C" {

int gl_assemble_plan_create(gl_ctx *ctx, const gl_mesh *cmesh, int op, uint64_t b, uint64_t e, const gl_args *a, const int32_t *eq,
                            int eq_location, const int64_t *rowptr, const int32_t *colidx, uint64_t *missing, gl_asm_plan **out) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    gl_mesh *mesh = const_cast<gl_mesh *>(cmesh);
    if (!a) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    gl_args args = *a;
    const int upper_only = args.packed ? 1 : 0; // for a plan, `packed` selects symmetric storage: only entries with row <= col are assembled
    args.packed = 0;
    args.clear = 1;
    int st = assemble_checks(ctx, mesh, op, b, e, &args);
    if (st) return st;
    const bool is_mat = op_is_matrix(op);
    if (!eq || !out || (is_mat && (!rowptr || !colidx))) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    if (missing) *missing = 0;
    cudaSetDevice(ctx->device);
    uint64_t total = 0;
    st = range_layout(mesh, op, b, e, &args, &total, nullptr);
    if (st) return fail(ctx, st);
    auto *plan = new gl_asm_plan();
    plan->mesh = mesh;
    plan->op = op;
    plan->b = b;
    plan->e = e;
    plan->total = total;
    plan->args = args;
    plan->is_mat = is_mat;
    if (total == 0) {
        *out = plan;
        return GL_OK;
    }
    if (cudaMalloc(&plan->d_pos, total * sizeof(uint32_t)) != cudaSuccess) {
        delete plan;
        return cuda_fail(ctx, cudaGetLastError(), "cudaMalloc (position map)");
    }
    const int32_t *d_eq = nullptr;
    bool owned = false;
    st = eq_on_device(ctx, mesh, op_dofs_per_node(op, mesh->ndim), eq, eq_location, &d_eq, &owned);
    unsigned long long *d_missing = nullptr;
    if (!st && cudaMalloc(&d_missing, sizeof(unsigned long long)) != cudaSuccess) st = cuda_fail(ctx, cudaGetLastError(), "cudaMalloc");
    if (!st) {
        cudaMemsetAsync(d_missing, 0, sizeof(unsigned long long), (cudaStream_t)ctx->stream);
        for (auto &bk : mesh->buckets) {
            ScatterParams s;
            bool empty;
            st = scatter_params(ctx, mesh, bk, op, &args, b, e, d_eq, s, empty);
            if (st) break;
            if (empty) continue;
            plan_positions_kernel<<<scatter_grid(s), 256, 0, (cudaStream_t)ctx->stream>>>(s, (const long long *)rowptr, colidx, plan->d_pos, d_missing,
                                                                                          is_mat ? upper_only : 0);
            ctx->launches++;
        }
    }
    unsigned long long h_missing = 0;
    if (d_missing) cudaMemcpyAsync(&h_missing, d_missing, sizeof(h_missing), cudaMemcpyDeviceToHost, (cudaStream_t)ctx->stream);
    const cudaError_t ce = cudaStreamSynchronize((cudaStream_t)ctx->stream);
    if (d_missing) cudaFree(d_missing);
    if (owned) cudaFree(const_cast<int32_t *>(d_eq));
    if (missing) *missing = h_missing;
    if (!st && ce != cudaSuccess) st = cuda_fail(ctx, ce, "position map");
    if (!st && h_missing) st = fail(ctx, GL_ERR_INVALID_ARG, "sparsity pattern lacks entries of the element blocks");
    if (st) {
        cudaFree(plan->d_pos);
        delete plan;
        return st;
    }
    *out = plan;
    return GL_OK;
}

void gl_assemble_plan_free(gl_ctx *ctx, gl_asm_plan *plan) {
    if (!plan) return;
    if (ctx) cudaSetDevice(ctx->device);
    if (plan->d_pos) cudaFree(plan->d_pos);
    delete plan;
}

int gl_assemble_planned(gl_ctx *ctx, const gl_asm_plan *plan, const gl_coef *coef, double *vals) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    if (!plan || !coef || !vals) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    gl_mesh *mesh = plan->mesh;
    if (mesh->device != ctx->device) return fail(ctx, GL_ERR_INVALID_ARG, "plan belongs to another device");
    if (plan->total == 0) return GL_OK;
    cudaSetDevice(ctx->device);
    const int op = plan->op;
    const uint64_t b = plan->b, e = plan->e;
    const gl_args *a = &plan->args;
    CoefResolved cr;
    int st = resolve_coef(ctx, mesh, op, b, e, a, coef, cr);
    if (st) return st;
    for (size_t bi = 0; bi < mesh->buckets.size(); bi++) {
        auto &bk = mesh->buckets[bi];
        uint64_t lo, hi;
        bucket_range(mesh, bk, b, e, lo, hi);
        if (lo >= hi) continue;
        if (stiffness_fusable(mesh, bk, op, a, coef)) {
            st = launch_range(ctx, mesh, op, b, e, a, cr, b, nullptr, nullptr, nullptr, nullptr, plan->d_pos, vals, (int)bi);
            if (st) return st;
            continue;
        }
        // everything else: chunks of cells through the scratch slab, then the position-map scatter
        const size_t avg = (size_t)std::max<uint64_t>(plan->total / (e - b), 1) * sizeof(double);
        const uint64_t per_chunk = std::max<uint64_t>(((size_t)256 << 20) / avg, 1);
        for (uint64_t cb = b; cb < e; cb += per_chunk) {
            const uint64_t ce = std::min(e, cb + per_chunk);
            uint64_t t = 0, before = 0;
            st = range_layout(mesh, op, cb, ce, a, &t, nullptr);
            if (!st) st = range_layout(mesh, op, b, cb, a, &before, nullptr);
            if (st) return fail(ctx, st);
            st = ensure_buffer(ctx, &ctx->d_full, &ctx->full_bytes, t * sizeof(double));
            if (st) return st;
            st = launch_range(ctx, mesh, op, cb, ce, a, cr, b, ctx->d_full, nullptr, nullptr, nullptr, nullptr, nullptr, (int)bi);
            if (st) return st;
            ScatterParams s;
            bool empty;
            st = scatter_params(ctx, mesh, bk, op, a, cb, ce, nullptr, s, empty);
            if (st) return st;
            if (empty) continue;
            planned_scatter_kernel<<<scatter_grid(s), 256, 0, (cudaStream_t)ctx->stream>>>(s, ctx->d_full, plan->d_pos, before, vals);
            ctx->launches++;
        }
    }
    if (cudaGetLastError() != cudaSuccess) return cuda_fail(ctx, cudaGetLastError(), "planned assembly");
    return GL_OK;
}

} // extern "C"

// =====================================================================================================
// device-side lattice generator: Block::subdivide for axis-aligned boxes, built in HBM
// =====================================================================================================
// Ordering contract of the reference (src/mesh/block.rs:749-761, 770-812; fixtures src/mesh/samples.rs:1407, 1901): cells
// are numbered i + nx*(j + ny*k); a point receives the next free id the first time a cell touches it, scanning cells in id
// order and nodes in local order.  For the linear targets that numbering has a closed form: lattice point (a,b,c) is first
// touched by its "owner" cell (max(a-1,0), max(b-1,0), max(c-1,0)); its id is the number of points owned by cells with a
// smaller id plus its rank among the owner's new nodes in local order.  tests/test_gpu_block.py checks it against the
// host Block (which is pinned by the reference's sample meshes).
namespace {

// natural coordinates of the Qua9 / Hex20 nodes (src/shapes/qua9.rs, hex20.rs): lattice offsets of the targets, control points of a block
const signed char QUA_REF[9][2] = {{-1, -1}, {1, -1}, {1, 1}, {-1, 1}, {0, -1}, {1, 0}, {0, 1}, {-1, 0}, {0, 0}};
const signed char HEX_REF[20][3] = {{-1, -1, -1}, {1, -1, -1}, {1, 1, -1}, {-1, 1, -1}, {-1, -1, 1}, {1, -1, 1}, {1, 1, 1}, {-1, 1, 1},
                                    {0, -1, -1}, {1, 0, -1}, {0, 1, -1}, {-1, 0, -1}, {0, -1, 1}, {1, 0, 1}, {0, 1, 1}, {-1, 0, 1},
                                    {-1, -1, 0}, {1, -1, 0}, {1, 1, 0}, {-1, 1, 0}};

__device__ __constant__ int LOC_HEX8[8][3] = {{0, 0, 0}, {1, 0, 0}, {1, 1, 0}, {0, 1, 0}, {0, 0, 1}, {1, 0, 1}, {1, 1, 1}, {0, 1, 1}};
__device__ __constant__ int LOC_QUA4[4][2] = {{0, 0}, {1, 0}, {1, 1}, {0, 1}};

__device__ __forceinline__ unsigned long long owned_before(unsigned long long n) { return n + (n > 0 ? 1ull : 0ull); }

// first-touch id of lattice point (a, b, c); ndim = 2 ignores c / nz
__device__ unsigned long long lattice_point_id(int ndim, unsigned a, unsigned b, unsigned c, unsigned nx, unsigned ny) {
    const unsigned i = a > 0 ? a - 1 : 0, j = b > 0 ? b - 1 : 0, k = (ndim == 3 && c > 0) ? c - 1 : 0;
    const unsigned long long ax = (unsigned long long)nx + 1, ay = (unsigned long long)ny + 1;
    unsigned long long base = owned_before(j) * ax + (j == 0 ? 2ull : 1ull) * owned_before(i);
    if (ndim == 3) base = owned_before(k) * ay * ax + (k == 0 ? 2ull : 1ull) * base;
    const int dx = (int)(a - i), dy = (int)(b - j), dz = ndim == 3 ? (int)(c - k) : 0;
    int rank = 0;
    const int nloc = ndim == 3 ? 8 : 4;
    for (int q = 0; q < nloc; q++) {
        const int qx = ndim == 3 ? LOC_HEX8[q][0] : LOC_QUA4[q][0];
        const int qy = ndim == 3 ? LOC_HEX8[q][1] : LOC_QUA4[q][1];
        const int qz = ndim == 3 ? LOC_HEX8[q][2] : 0;
        if (qx == dx && qy == dy && qz == dz) break;
        const bool is_new = (qx == 1 || i == 0) && (qy == 1 || j == 0) && (ndim == 2 || qz == 1 || k == 0);
        rank += is_new ? 1 : 0;
    }
    return base + (unsigned long long)rank;
}

struct BlockGen {
    int ndim;
    unsigned nx, ny, nz;
    double x0[3], h[3];
    unsigned long long ncell, npoint;
};

__global__ void block_conn_kernel(const BlockGen g, uint32_t *conn) {
    const int nn = g.ndim == 3 ? 8 : 4;
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < g.ncell;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned i = (unsigned)(e % g.nx), j = (unsigned)((e / g.nx) % g.ny), k = (unsigned)(e / ((unsigned long long)g.nx * g.ny));
        for (int m = 0; m < nn; m++) {
            const unsigned a = i + (g.ndim == 3 ? LOC_HEX8[m][0] : LOC_QUA4[m][0]);
            const unsigned b = j + (g.ndim == 3 ? LOC_HEX8[m][1] : LOC_QUA4[m][1]);
            const unsigned c = k + (g.ndim == 3 ? LOC_HEX8[m][2] : 0);
            conn[(unsigned long long)m * g.ncell + e] = (uint32_t)lattice_point_id(g.ndim, a, b, c, g.nx, g.ny);
        }
    }
}

__global__ void block_coords_kernel(const BlockGen g, double *coords) {
    const unsigned long long ax = (unsigned long long)g.nx + 1, ay = (unsigned long long)g.ny + 1;
    for (unsigned long long q = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; q < g.npoint;
         q += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned a = (unsigned)(q % ax), b = (unsigned)((q / ax) % ay), c = (unsigned)(q / (ax * ay));
        const unsigned long long id = lattice_point_id(g.ndim, a, b, c, g.nx, g.ny);
        coords[id] = g.x0[0] + g.h[0] * (double)a;
        coords[g.npoint + id] = g.x0[1] + g.h[1] * (double)b;
        if (g.ndim == 3) coords[2 * g.npoint + id] = g.x0[2] + g.h[2] * (double)c;
    }
}

__global__ void soa_to_aos_kernel(const double *soa, double *aos, unsigned long long npoint, int ndim) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < npoint * ndim;
         i += (unsigned long long)gridDim.x * blockDim.x)
        aos[i] = soa[(i % ndim) * npoint + i / ndim];
}

__global__ void conn_to_cell_major_kernel(const uint32_t *nm, uint32_t *cm, unsigned long long ncell, int nnode) {
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < ncell * nnode;
         i += (unsigned long long)gridDim.x * blockDim.x)
        cm[i] = nm[(i % nnode) * ncell + i / nnode];
}

} // namespace

extern "C" {

int gl_mesh_block(gl_ctx *ctx, int kind, const uint32_t *ndiv, const double *xmin, const double *xmax, int32_t marker, gl_mesh **out) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    if (!out || !ndiv || !xmin || !xmax) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (kind == GL_QUA8 || kind == GL_QUA9 || kind == GL_HEX20) {
        // quadratic targets: the general generator on the box's corners (the closed form below covers the linear targets)
        const int nd = kind == GL_HEX20 ? 3 : 2;
        double corners[8 * 3];
        for (int q = 0; q < (nd == 2 ? 4 : 8); q++)
            for (int d = 0; d < nd; d++) corners[q * nd + d] = (nd == 2 ? QUA_REF[q][d] : HEX_REF[q][d]) < 0 ? xmin[d] : xmax[d];
        return gl_mesh_block_general(ctx, kind, ndiv, nd == 2 ? 4 : 8, corners, marker, out);
    }
    if (kind != GL_QUA4 && kind != GL_HEX8) return fail(ctx, GL_ERR_KIND_UNSUPPORTED, "the device lattice generator builds Qua4, Qua8, Qua9, Hex8 and Hex20 meshes");
    const int ndim = kind == GL_HEX8 ? 3 : 2;
    BlockGen g;
    g.ndim = ndim;
    g.nx = ndiv[0];
    g.ny = ndiv[1];
    g.nz = ndim == 3 ? ndiv[2] : 1;
    if (g.nx == 0 || g.ny == 0 || g.nz == 0) return fail(ctx, GL_ERR_INVALID_ARG, "ndiv must be >= 1");
    g.ncell = (unsigned long long)g.nx * g.ny * g.nz;
    g.npoint = ((unsigned long long)g.nx + 1) * ((unsigned long long)g.ny + 1) * (ndim == 3 ? (unsigned long long)g.nz + 1 : 1ull);
    if (g.npoint > 0xffffffffull) return fail(ctx, GL_ERR_MESH, "npoint/ncell must fit in 32 bits");
    for (int d = 0; d < 3; d++) {
        g.x0[d] = d < ndim ? xmin[d] : 0.0;
        const unsigned n = d == 0 ? g.nx : (d == 1 ? g.ny : g.nz);
        g.h[d] = d < ndim ? (xmax[d] - xmin[d]) / (double)n : 0.0;
    }
    cudaSetDevice(ctx->device);
    auto mesh = std::make_unique<gl_mesh>();
    mesh->ndim = ndim;
    mesh->npoint = g.npoint;
    mesh->ncell = g.ncell;
    mesh->device = ctx->device;
    mesh->homogeneous = true;
    mesh->marker_values.assign(1, marker);
    gl_mesh_bucket bk;
    bk.kind = kind;
    bk.nnode = kind_nnode(kind);
    bk.ncell = g.ncell;
    CUDA_TRY(ctx, pool_alloc_t(ctx, &mesh->d_coords, (size_t)g.npoint * ndim * sizeof(double)));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &bk.d_conn, (size_t)g.ncell * bk.nnode * sizeof(uint32_t)));
    cudaStream_t st = (cudaStream_t)ctx->stream;
    block_conn_kernel<<<1184, 256, 0, st>>>(g, bk.d_conn);
    block_coords_kernel<<<1184, 256, 0, st>>>(g, mesh->d_coords);
    ctx->launches += 2;
    if (cudaGetLastError() != cudaSuccess) return cuda_fail(ctx, cudaGetLastError(), "lattice generator");
    mesh->buckets.push_back(std::move(bk));
    *out = mesh.release();
    return GL_OK;
}

} // extern "C"

// ---- general lattice generator: every Block target (Qua4 / Qua8 / Qua9 / Hex8 / Hex20) on a straight-edged or curved block ------------
// The reference walks the cells in id order and gives a point the next free id the first time a cell touches it (GridSearch,
// src/mesh/block.rs:749-812).  On the device the same numbering comes from three passes over a lattice with f = 1 (linear targets)
// or 2 (quadratic) steps per cell: (1) atomicMin of the key cell * nnode + m at every lattice point a cell touches; (2) per cell, the
// number of nodes whose key won = the points the cell introduces, turned into an exclusive prefix over the cells; (3) the cell's new
// points get prefix + rank in local order, their coordinates come from the block's own interpolation (bi/tri-linear map of the
// corners, or the Qua8 / Hex20 serendipity map of a curved block: block.rs:794).
namespace {

struct LatGen {
    int ndim, nnode, f, ncontrol;
    unsigned nx, ny, nz;
    unsigned long long lat0, lat1, lat2, ncell;
    signed char off[20][3]; // lattice offsets of the local nodes, in 0 .. f
    signed char ref[20][3]; // natural coordinates of the block's control points (-1, 0, 1)
    double ctrl[20][3];
};

__device__ __forceinline__ unsigned long long lat_index(const LatGen &g, unsigned long long e, int m) {
    const unsigned long long i = e % g.nx, j = (e / g.nx) % g.ny, k = e / ((unsigned long long)g.nx * g.ny);
    const unsigned long long a = g.f * i + g.off[m][0], b = g.f * j + g.off[m][1], c = g.ndim == 3 ? g.f * k + g.off[m][2] : 0ull;
    return a + g.lat0 * (b + g.lat1 * c);
}

__global__ void lat_first_kernel(const LatGen g, unsigned long long *first) {
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < g.ncell * g.nnode;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = t / g.nnode;
        const int m = (int)(t - e * g.nnode);
        atomicMin(first + lat_index(g, e, m), t);
    }
}

__global__ void lat_count_kernel(const LatGen g, const unsigned long long *first, unsigned int *count) {
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < g.ncell;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned int n = 0;
        for (int m = 0; m < g.nnode; m++) n += first[lat_index(g, e, m)] == e * g.nnode + m ? 1u : 0u;
        count[e] = n;
    }
}

// single-CTA exclusive scan in chunks (the generator is not on the hot path; 10^7 cells take well under a millisecond)
__global__ void lat_scan_kernel(const unsigned int *count, unsigned int *prefix, unsigned long long n) {
    __shared__ unsigned int warp_sum[32];
    __shared__ unsigned int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for (unsigned long long base = 0; base < n; base += blockDim.x) {
        const unsigned long long i = base + threadIdx.x;
        const unsigned int v = i < n ? count[i] : 0u;
        unsigned int x = v;
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned int y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= d) x += y;
        }
        if (lane == 31) warp_sum[wid] = x;
        __syncthreads();
        if (wid == 0) {
            unsigned int w = lane < (int)(blockDim.x >> 5) ? warp_sum[lane] : 0u;
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned int y = __shfl_up_sync(0xffffffffu, w, d);
                if (lane >= d) w += y;
            }
            warp_sum[lane] = w; // inclusive sums of the warps
        }
        __syncthreads();
        const unsigned int before = carry + (wid > 0 ? warp_sum[wid - 1] : 0u);
        if (i < n) prefix[i] = before + x - v;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = before + x;
        __syncthreads();
    }
}

__device__ void block_interp(const LatGen &g, const double *ksi, double *x) {
    for (int d = 0; d < g.ndim; d++) x[d] = 0.0;
    const int nd = g.ndim;
    const int ncorner = nd == 2 ? 4 : 8;
    for (int q = 0; q < g.ncontrol; q++) {
        double w;
        if (g.ncontrol == ncorner) { // straight-edged block: bi / tri-linear map of the corners
            w = nd == 2 ? 0.25 : 0.125;
            for (int d = 0; d < nd; d++) w *= 1.0 + ksi[d] * (double)g.ref[q][d];
        } else if (q < ncorner) { // serendipity corner
            double prod = 1.0, gsum = -(double)(nd - 1);
            for (int d = 0; d < nd; d++) {
                prod *= 1.0 + ksi[d] * (double)g.ref[q][d];
                gsum += ksi[d] * (double)g.ref[q][d];
            }
            w = prod * gsum / (nd == 2 ? 4.0 : 8.0);
        } else { // serendipity mid-edge node
            w = 1.0;
            for (int d = 0; d < nd; d++) w *= g.ref[q][d] == 0 ? (1.0 - ksi[d] * ksi[d]) : (1.0 + ksi[d] * (double)g.ref[q][d]);
            w /= nd == 2 ? 2.0 : 4.0;
        }
        for (int d = 0; d < nd; d++) x[d] += w * g.ctrl[q][d];
    }
}

__global__ void lat_number_kernel(const LatGen g, const unsigned long long *first, const unsigned int *prefix, uint32_t *point_of_lattice,
                                  double *coords, unsigned long long npoint) {
    for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < g.ncell;
         e += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned int id = prefix[e];
        for (int m = 0; m < g.nnode; m++) {
            const unsigned long long li = lat_index(g, e, m);
            if (first[li] != e * g.nnode + m) continue;
            point_of_lattice[li] = id;
            const unsigned long long a = li % g.lat0, b = (li / g.lat0) % g.lat1, c = li / (g.lat0 * g.lat1);
            double ksi[3], x[3];
            ksi[0] = -1.0 + 2.0 * (double)a / (double)(g.f * g.nx);
            ksi[1] = -1.0 + 2.0 * (double)b / (double)(g.f * g.ny);
            ksi[2] = g.ndim == 3 ? -1.0 + 2.0 * (double)c / (double)(g.f * g.nz) : 0.0;
            block_interp(g, ksi, x);
            for (int d = 0; d < g.ndim; d++) coords[(unsigned long long)d * npoint + id] = x[d];
            id++;
        }
    }
}

__global__ void lat_conn_kernel(const LatGen g, const uint32_t *point_of_lattice, uint32_t *conn) {
    for (unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; t < g.ncell * g.nnode;
         t += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned long long e = t / g.nnode;
        const int m = (int)(t - e * g.nnode);
        conn[(unsigned long long)m * g.ncell + e] = point_of_lattice[lat_index(g, e, m)];
    }
}

} // namespace

extern "C" {

int gl_mesh_block_general(gl_ctx *ctx, int kind, const uint32_t *ndiv, int ncontrol, const double *control, int32_t marker, gl_mesh **out) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    if (!out || !ndiv || !control) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    *out = nullptr;
    if (kind != GL_QUA4 && kind != GL_QUA8 && kind != GL_QUA9 && kind != GL_HEX8 && kind != GL_HEX20)
        return fail(ctx, GL_ERR_KIND_UNSUPPORTED, "the lattice generator builds Qua4, Qua8, Qua9, Hex8 and Hex20 meshes");
    const int ndim = kind_geo_ndim(kind);
    if (!((ndim == 2 && (ncontrol == 4 || ncontrol == 8)) || (ndim == 3 && (ncontrol == 8 || ncontrol == 20))))
        return fail(ctx, GL_ERR_INVALID_ARG, "a block has 4 corners or 8 Qua8 nodes in 2D, 8 corners or 20 Hex20 nodes in 3D");
    LatGen g;
    std::memset(&g, 0, sizeof(g));
    g.ndim = ndim;
    g.nnode = kind_nnode(kind);
    g.f = (kind == GL_QUA4 || kind == GL_HEX8) ? 1 : 2;
    g.ncontrol = ncontrol;
    g.nx = ndiv[0];
    g.ny = ndiv[1];
    g.nz = ndim == 3 ? ndiv[2] : 1;
    if (g.nx == 0 || g.ny == 0 || g.nz == 0) return fail(ctx, GL_ERR_INVALID_ARG, "ndiv must be >= 1");
    g.ncell = (unsigned long long)g.nx * g.ny * g.nz;
    g.lat0 = (unsigned long long)g.f * g.nx + 1;
    g.lat1 = (unsigned long long)g.f * g.ny + 1;
    g.lat2 = ndim == 3 ? (unsigned long long)g.f * g.nz + 1 : 1;
    for (int m = 0; m < g.nnode; m++)
        for (int d = 0; d < ndim; d++) g.off[m][d] = (signed char)(((ndim == 2 ? QUA_REF[m][d] : HEX_REF[m][d]) + 1) * g.f / 2);
    for (int q = 0; q < ncontrol; q++)
        for (int d = 0; d < ndim; d++) {
            g.ref[q][d] = ndim == 2 ? QUA_REF[q][d] : HEX_REF[q][d];
            g.ctrl[q][d] = control[q * ndim + d];
        }
    const unsigned long long nlat = g.lat0 * g.lat1 * g.lat2;
    const unsigned long long nx = g.nx, ny = g.ny, nz = g.nz;
    unsigned long long npoint = nlat; // Qua4 / Hex8 / Qua9: every lattice point
    if (kind == GL_QUA8) npoint = nlat - nx * ny; // no cell centres
    if (kind == GL_HEX20) npoint = (nx + 1) * (ny + 1) * (nz + 1) + nx * (ny + 1) * (nz + 1) + (nx + 1) * ny * (nz + 1) + (nx + 1) * (ny + 1) * nz;
    if (npoint > 0xffffffffull || g.ncell > 0xffffffffull) return fail(ctx, GL_ERR_MESH, "npoint/ncell must fit in 32 bits");
    cudaSetDevice(ctx->device);
    auto mesh = std::make_unique<gl_mesh>();
    mesh->ndim = ndim;
    mesh->npoint = npoint;
    mesh->ncell = g.ncell;
    mesh->device = ctx->device;
    mesh->homogeneous = true;
    mesh->marker_values.assign(1, marker);
    gl_mesh_bucket bk;
    bk.kind = kind;
    bk.nnode = g.nnode;
    bk.ncell = g.ncell;
    CUDA_TRY(ctx, pool_alloc_t(ctx, &mesh->d_coords, (size_t)npoint * ndim * sizeof(double)));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &bk.d_conn, (size_t)g.ncell * bk.nnode * sizeof(uint32_t)));
    unsigned long long *first = nullptr;
    uint32_t *pol = nullptr;
    unsigned int *count = nullptr, *prefix = nullptr;
    const size_t b_first = nlat * sizeof(unsigned long long), b_pol = nlat * sizeof(uint32_t), b_cnt = g.ncell * sizeof(unsigned int);
    CUDA_TRY(ctx, pool_alloc_t(ctx, &first, b_first));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &pol, b_pol));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &count, b_cnt));
    CUDA_TRY(ctx, pool_alloc_t(ctx, &prefix, b_cnt));
    cudaStream_t st = (cudaStream_t)ctx->stream;
    CUDA_TRY(ctx, cudaMemsetAsync(first, 0xff, b_first, st));
    lat_first_kernel<<<1184, 256, 0, st>>>(g, first);
    lat_count_kernel<<<1184, 256, 0, st>>>(g, first, count);
    lat_scan_kernel<<<1, 1024, 0, st>>>(count, prefix, g.ncell);
    lat_number_kernel<<<1184, 256, 0, st>>>(g, first, prefix, pol, mesh->d_coords, npoint);
    lat_conn_kernel<<<1184, 256, 0, st>>>(g, pol, bk.d_conn);
    ctx->launches += 5;
    const cudaError_t ce = cudaGetLastError();
    // the scratch arrays go back to the pool; everything that touches them is ordered on the ctx stream
    pool_free(ctx, first, b_first);
    pool_free(ctx, pol, b_pol);
    pool_free(ctx, count, b_cnt);
    pool_free(ctx, prefix, b_cnt);
    if (ce != cudaSuccess) return cuda_fail(ctx, ce, "lattice generator");
    mesh->buckets.push_back(std::move(bk));
    *out = mesh.release();
    return GL_OK;
}

int gl_mesh_download(gl_ctx *ctx, const gl_mesh *mesh, double *coords, uint32_t *conn) {
    if (!ctx) return GL_ERR_INVALID_ARG;
    if (!mesh) return fail(ctx, GL_ERR_INVALID_ARG, "null pointer");
    cudaSetDevice(ctx->device);
    cudaStream_t st = (cudaStream_t)ctx->stream;
    if (coords) {
        double *tmp = nullptr;
        const size_t bytes = (size_t)mesh->npoint * mesh->ndim * sizeof(double);
        CUDA_TRY(ctx, pool_alloc_t(ctx, &tmp, bytes));
        soa_to_aos_kernel<<<1184, 256, 0, st>>>(mesh->d_coords, tmp, mesh->npoint, mesh->ndim);
        ctx->launches++;
        CUDA_TRY(ctx, cudaMemcpyAsync(coords, tmp, bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        pool_free(ctx, tmp, bytes);
    }
    if (conn) {
        if (!mesh->homogeneous || mesh->buckets.size() != 1) return fail(ctx, GL_ERR_MESH, "connectivity download needs a single-kind mesh");
        const gl_mesh_bucket &bk = mesh->buckets[0];
        uint32_t *tmp = nullptr;
        const size_t bytes = (size_t)bk.ncell * bk.nnode * sizeof(uint32_t);
        CUDA_TRY(ctx, pool_alloc_t(ctx, &tmp, bytes));
        conn_to_cell_major_kernel<<<1184, 256, 0, st>>>(bk.d_conn, tmp, bk.ncell, bk.nnode);
        ctx->launches++;
        CUDA_TRY(ctx, cudaMemcpyAsync(conn, tmp, bytes, cudaMemcpyDeviceToHost, st));
        CUDA_TRY(ctx, cudaStreamSynchronize(st));
        pool_free(ctx, tmp, bytes);
    }
    return GL_OK;
}

} // extern "C"

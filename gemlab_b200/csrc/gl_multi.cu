// gl_multi.cu -- (1) device-resident output slabs (gl_slab) and (2) the single-process multi-GPU driver of SURVEY.md 8(e):
// one gl_ctx + stream + host thread per device, the cell range split into contiguous pieces
//      piece g of [b, e) over G devices = [b + floor(g n / G), b + floor((g+1) n / G)),   n = e - b
// with NO collective (cells are independent); host outputs land concatenated in cell order, so the result is bit-identical
// to the single-device call; device outputs stay on their devices as one gl_slab per device.
//
// The reference is single threaded (one &mut Scratchpad per call, src/integ/common_args.rs:5-43); a Rust caller of the batched
// boundary would otherwise hand-roll G contexts and G threads -- this is that loop, behind the C ABI.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gl_internal.hpp"

using namespace gl;

struct gl_multi {
    std::vector<gl_ctx *> ctx;
    std::vector<int> device;
    std::string last_error;
};

struct gl_multi_mesh {
    std::vector<gl_mesh *> mesh; // one replica per device (the path shards cells, not points)
};

namespace {

inline void piece(uint64_t b, uint64_t e, int G, int g, uint64_t &pb, uint64_t &pe) {
    const uint64_t n = e - b;
    // 128-bit products: g n can exceed 64 bits only for absurd inputs, but the rule must match partition() exactly
    pb = b + (uint64_t)(((unsigned __int128)n * (unsigned)g) / (unsigned)G);
    pe = b + (uint64_t)(((unsigned __int128)n * (unsigned)(g + 1)) / (unsigned)G);
}

int multi_fail(gl_multi *m, int status, const std::string &msg) {
    if (m) m->last_error = msg;
    return status;
}

// Gauss points per cell of the kinds in [b, e) (PER_GAUSS records are laid out cell by cell)
int range_ngauss(const gl_mesh *mesh, uint64_t b, uint64_t e, const gl_args *a, int *ng_out) {
    int ng_common = -1;
    for (auto &bk : mesh->buckets) {
        bool hit = mesh->homogeneous ? (e > b) : false;
        if (!mesh->homogeneous)
            for (uint32_t id : bk.h_cell_ids)
                if (id >= b && id < e) { hit = true; break; }
        if (!hit) continue;
        int ng = 0;
        const double *data = nullptr;
        const int st = gauss_rule(bk.kind, a->ngauss, &ng, &data);
        if (st) return st;
        if (ng_common >= 0 && ng != ng_common) return GL_ERR_COEF;
        ng_common = ng;
    }
    *ng_out = std::max(ng_common, 1);
    return GL_OK;
}

// the coefficient of piece [pb, pe) of a call over [b, e): per-cell / per-Gauss host records advance with the piece
int piece_coef(const gl_mesh *mesh, int op, uint64_t b, uint64_t e, uint64_t pb, const gl_args *a, const gl_coef *coef, gl_coef *out) {
    *out = *coef;
    if (coef->mode == GL_COEF_PER_CELL || coef->mode == GL_COEF_PER_GAUSS) {
        uint64_t per_cell = gl_op_coef_size(op, mesh->ndim);
        if (coef->mode == GL_COEF_PER_GAUSS) {
            int ng = 1;
            const int st = range_ngauss(mesh, b, e, a, &ng);
            if (st) return st;
            per_cell *= (uint64_t)ng;
        }
        out->data = coef->data + (pb - b) * per_cell;
    }
    return GL_OK;
}

} // namespace

extern "C" {

// =====================================================================================================
// device-resident output slabs
// =====================================================================================================

int gl_integ_slab(gl_ctx *ctx, const gl_mesh *mesh, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef, gl_slab *slab) {
    if (!ctx || !mesh || !a || !slab) return GL_ERR_INVALID_ARG;
    uint64_t total = 0;
    int st = gl_out_total(mesh, op, b, e, a, &total);
    if (st) return st;
    cudaSetDevice(ctx->device);
    if (slab->ptr == nullptr) {
        double *ptr = nullptr;
        if (cudaMalloc(&ptr, std::max<uint64_t>(total, 2) * sizeof(double)) != cudaSuccess) {
            cudaGetLastError();
            return GL_ERR_CUDA;
        }
        slab->ptr = ptr;
        slab->capacity = std::max<uint64_t>(total, 2);
        slab->owned = 1;
        if (!a->clear) cudaMemsetAsync(ptr, 0, total * sizeof(double), (cudaStream_t)ctx->stream); // accumulate into zeros
    }
    slab->device = ctx->device;
    slab->cell_begin = b;
    slab->cell_end = e;
    slab->size = total;
    gl_out out{slab->ptr, GL_DEVICE, 0, slab->capacity};
    st = gl_integ(ctx, mesh, op, b, e, a, coef, &out);
    if (st && slab->owned) gl_slab_free(slab); // nothing half-made is handed back
    return st;
}

int gl_slab_download(gl_ctx *ctx, const gl_slab *slab, double *host) {
    if (!ctx || !slab || !host) return GL_ERR_INVALID_ARG;
    if (slab->device != ctx->device) return GL_ERR_INVALID_ARG;
    cudaSetDevice(ctx->device);
    if (slab->size &&
        cudaMemcpyAsync(host, slab->ptr, slab->size * sizeof(double), cudaMemcpyDeviceToHost, (cudaStream_t)ctx->stream) != cudaSuccess)
        return GL_ERR_CUDA;
    return gl_ctx_sync(ctx);
}

void gl_slab_free(gl_slab *slab) {
    if (!slab) return;
    if (slab->owned && slab->ptr) {
        int cur = 0;
        cudaGetDevice(&cur);
        cudaSetDevice(slab->device);
        cudaFree(slab->ptr);
        cudaSetDevice(cur);
    }
    slab->ptr = nullptr;
    slab->capacity = slab->size = 0;
    slab->owned = 0;
}

// =====================================================================================================
// single-process multi-GPU driver
// =====================================================================================================

int gl_multi_create(const int *devices, int ndevice, gl_multi **out) {
    if (!devices || ndevice < 1 || !out) return GL_ERR_INVALID_ARG;
    auto *m = new gl_multi();
    for (int g = 0; g < ndevice; g++) {
        gl_ctx *ctx = nullptr;
        cudaStream_t stream = nullptr;
        int st = GL_OK;
        if (cudaSetDevice(devices[g]) != cudaSuccess || cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking) != cudaSuccess) {
            cudaGetLastError();
            st = GL_ERR_NO_DEVICE;
        }
        if (!st) st = gl_ctx_create(devices[g], stream, &ctx);
        if (st) {
            for (auto *c : m->ctx) {
                cudaStream_t s = (cudaStream_t)c->stream;
                gl_ctx_destroy(c);
                if (s) cudaStreamDestroy(s);
            }
            if (stream) cudaStreamDestroy(stream);
            delete m;
            return st;
        }
        m->ctx.push_back(ctx);
        m->device.push_back(devices[g]);
    }
    *out = m;
    return GL_OK;
}

void gl_multi_destroy(gl_multi *m) {
    if (!m) return;
    for (auto *c : m->ctx) {
        cudaSetDevice(c->device);
        cudaStream_t s = (cudaStream_t)c->stream;
        gl_ctx_destroy(c);
        if (s) cudaStreamDestroy(s);
    }
    delete m;
}

int gl_multi_ndevice(const gl_multi *m) { return m ? (int)m->ctx.size() : 0; }
gl_ctx *gl_multi_ctx(gl_multi *m, int g) { return (m && g >= 0 && g < (int)m->ctx.size()) ? m->ctx[g] : nullptr; }
const char *gl_multi_last_error(const gl_multi *m) { return m ? m->last_error.c_str() : ""; }

void gl_multi_partition(uint64_t cell_begin, uint64_t cell_end, int ndevice, int g, uint64_t *pb, uint64_t *pe) {
    uint64_t a = cell_begin, b = cell_begin;
    if (ndevice > 0 && g >= 0 && g < ndevice && cell_end >= cell_begin) piece(cell_begin, cell_end, ndevice, g, a, b);
    if (pb) *pb = a;
    if (pe) *pe = b;
}

int gl_multi_mesh_upload(gl_multi *m, int ndim, uint64_t npoint, const double *coords, uint64_t ncell, const uint8_t *kinds,
                         const int32_t *markers, const uint64_t *conn_off, const uint64_t *conn, gl_multi_mesh **out) {
    if (!m || !out) return GL_ERR_INVALID_ARG;
    const int G = (int)m->ctx.size();
    auto *mm = new gl_multi_mesh();
    mm->mesh.assign(G, nullptr);
    std::vector<int> st(G, GL_OK);
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
        th.emplace_back([&, g] { st[g] = gl_mesh_upload(m->ctx[g], ndim, npoint, coords, ncell, kinds, markers, conn_off, conn, &mm->mesh[g]); });
    for (auto &t : th) t.join();
    for (int g = 0; g < G; g++)
        if (st[g]) {
            const int bad = st[g];
            multi_fail(m, bad, std::string("device ") + std::to_string(m->device[g]) + ": " + gl_last_error(m->ctx[g]));
            for (int k = 0; k < G; k++)
                if (mm->mesh[k]) gl_mesh_free(m->ctx[k], mm->mesh[k]);
            delete mm;
            return bad;
        }
    *out = mm;
    return GL_OK;
}

void gl_multi_mesh_free(gl_multi *m, gl_multi_mesh *mm) {
    if (!m || !mm) return;
    for (size_t g = 0; g < mm->mesh.size() && g < m->ctx.size(); g++)
        if (mm->mesh[g]) gl_mesh_free(m->ctx[g], mm->mesh[g]);
    delete mm;
}

const gl_mesh *gl_multi_mesh_replica(const gl_multi_mesh *mm, int g) { return (mm && g >= 0 && g < (int)mm->mesh.size()) ? mm->mesh[g] : nullptr; }

static int multi_run(gl_multi *m, const gl_multi_mesh *mm, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef,
                     gl_out *host_out, gl_slab *slabs) {
    if (!m || !mm || !a || !coef) return GL_ERR_INVALID_ARG;
    const int G = (int)m->ctx.size();
    if ((int)mm->mesh.size() != G) return multi_fail(m, GL_ERR_INVALID_ARG, "mesh was uploaded through another gl_multi");
    const gl_mesh *m0 = mm->mesh[0];
    if (b > e || e > m0->ncell) return multi_fail(m, GL_ERR_CELL_RANGE, gl_status_string(GL_ERR_CELL_RANGE));
    if (coef->location == GL_DEVICE && G > 1 && coef->mode != GL_COEF_UNIFORM)
        return multi_fail(m, GL_ERR_COEF, "device-resident coefficient records belong to one device: pass host records to gl_multi_*");
    uint64_t total = 0;
    int st = gl_out_total(m0, op, b, e, a, &total);
    if (st) return multi_fail(m, st, gl_status_string(st));
    if (host_out) {
        if (host_out->location != GL_HOST) return multi_fail(m, GL_ERR_INVALID_ARG, "gl_multi_integ writes a HOST buffer; use gl_multi_integ_slabs for device outputs");
        if (e > b && (!host_out->ptr || host_out->capacity < total)) return multi_fail(m, GL_ERR_OUT_CAPACITY, gl_status_string(GL_ERR_OUT_CAPACITY));
    }
    std::vector<int> status(G, GL_OK);
    std::vector<uint64_t> pb(G), pe(G), off(G, 0);
    std::vector<gl_coef> pc(G);
    for (int g = 0; g < G; g++) {
        piece(b, e, G, g, pb[g], pe[g]);
        st = gl_out_total(m0, op, b, pb[g], a, &off[g]);
        if (!st) st = piece_coef(m0, op, b, e, pb[g], a, coef, &pc[g]);
        if (st) return multi_fail(m, st, gl_status_string(st));
    }
    std::vector<std::thread> th;
    for (int g = 0; g < G; g++)
        th.emplace_back([&, g] {
            gl_ctx *ctx = m->ctx[g];
            if (host_out) {
                uint64_t mine = 0;
                int s = gl_out_total(mm->mesh[g], op, pb[g], pe[g], a, &mine);
                gl_out o{host_out->ptr + off[g], GL_HOST, 0, mine};
                if (!s) s = gl_integ(ctx, mm->mesh[g], op, pb[g], pe[g], a, &pc[g], &o);
                status[g] = s;
            } else {
                status[g] = gl_integ_slab(ctx, mm->mesh[g], op, pb[g], pe[g], a, &pc[g], &slabs[g]);
            }
        });
    for (auto &t : th) t.join();
    // the smallest failing cell over all devices is the one the single-device call would report
    int bad = GL_OK;
    for (int g = 0; g < G; g++)
        if (status[g] && !bad) {
            bad = status[g];
            multi_fail(m, bad, std::string(gl_last_error(m->ctx[g])));
        }
    return bad;
}

int gl_multi_integ(gl_multi *m, const gl_multi_mesh *mm, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef, gl_out *out) {
    if (!out) return GL_ERR_INVALID_ARG;
    return multi_run(m, mm, op, b, e, a, coef, out, nullptr);
}

int gl_multi_integ_slabs(gl_multi *m, const gl_multi_mesh *mm, int op, uint64_t b, uint64_t e, const gl_args *a, const gl_coef *coef,
                         gl_slab *slabs) {
    if (!slabs) return GL_ERR_INVALID_ARG;
    return multi_run(m, mm, op, b, e, a, coef, nullptr, slabs);
}

int gl_multi_sync(gl_multi *m) {
    if (!m) return GL_ERR_INVALID_ARG;
    int bad = GL_OK;
    for (auto *c : m->ctx) {
        const int st = gl_ctx_sync(c);
        if (st && !bad) {
            bad = st;
            m->last_error = gl_last_error(c);
        }
    }
    return bad;
}

} // extern "C"

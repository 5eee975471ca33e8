// gemlab.hpp -- header-only C++ host mirror of gemlab's integration interface over the C ABI (include/gemlab_cuda.h).
//
// The reference is compiled Rust and its toolchain is not in the authoring image, so the compiled-language host side is
// written in C++ with the reference's names and argument meaning:
//   gemlab::GeoKind                         src/shapes/enums.rs:116-176 (VALUES order)
//   gemlab::Point / Cell / Mesh             src/mesh/mesh.rs:27-54, 110-135 (array-of-structures, exactly as the reference)
//   gemlab::integ::CommonArgs               src/integ/common_args.rs:5-43 (pad and gauss are implied by the mesh / ngauss)
//   gemlab::integ::batch::mat_10_bdb(...)   src/integ/mat_10_bdb.rs:121  -- and the siblings mat_01_nsn, mat_03_btb,
//                                           vec_01_ns, vec_02_nv, vec_03_bv, vec_04_bt
// Errors are thrown as gemlab::StrError carrying the reference's strings.
#pragma once

#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "../include/gemlab_cuda.h"

namespace gemlab {

enum class GeoKind : uint8_t {
    Lin2, Lin3, Lin4, Lin5, Tri3, Tri6, Tri10, Tri15, Qua4, Qua8, Qua9, Qua12, Qua16, Qua17, Tet4, Tet10, Tet20, Hex8, Hex20, Hex32
};

struct StrError : std::runtime_error {
    int status;
    uint64_t cell;
    StrError(int st, const std::string &msg, uint64_t c = ~0ull) : std::runtime_error(msg), status(st), cell(c) {}
};

struct Point {
    size_t id;
    int marker;
    std::vector<double> coords;
};

struct Cell {
    size_t id;
    int marker;
    GeoKind kind;
    std::vector<size_t> points;
};

struct Mesh {
    size_t ndim = 0;
    std::vector<Point> points;
    std::vector<Cell> cells;
};

class Context;

class DeviceMesh {
  public:
    DeviceMesh(const DeviceMesh &) = delete;
    DeviceMesh(DeviceMesh &&o) noexcept : ctx_(o.ctx_), mesh_(o.mesh_) { o.mesh_ = nullptr; }
    ~DeviceMesh() {
        if (mesh_) gl_mesh_free(ctx_, mesh_);
    }
    size_t ncell() const { return (size_t)gl_mesh_ncell(mesh_); }
    const gl_mesh *handle() const { return mesh_; }

  private:
    friend class Context;
    DeviceMesh(gl_ctx *ctx, gl_mesh *mesh) : ctx_(ctx), mesh_(mesh) {}
    gl_ctx *ctx_;
    gl_mesh *mesh_;
};

// The batched output of one call, resident on its device (struct gl_slab): element e of [cell_begin, cell_end) at
// ptr + off[e - cell_begin] (gl_out_offsets; homogeneous meshes: (e - cell_begin) * block).  Feed it to gl_triplet_indices, a sparse
// solver or your own kernels; download() brings it to the host.
class DeviceSlab {
  public:
    DeviceSlab() { slab_ = gl_slab{}; }
    DeviceSlab(const DeviceSlab &) = delete;
    DeviceSlab(DeviceSlab &&o) noexcept : slab_(o.slab_) { o.slab_ = gl_slab{}; }
    DeviceSlab &operator=(DeviceSlab &&o) noexcept {
        gl_slab_free(&slab_);
        slab_ = o.slab_;
        o.slab_ = gl_slab{};
        return *this;
    }
    ~DeviceSlab() { gl_slab_free(&slab_); }
    double *ptr() const { return slab_.ptr; }
    size_t size() const { return (size_t)slab_.size; }
    int device() const { return slab_.device; }
    std::pair<size_t, size_t> cells() const { return {(size_t)slab_.cell_begin, (size_t)slab_.cell_end}; }
    std::vector<double> download(gl_ctx *ctx) const {
        std::vector<double> host(slab_.size);
        if (int st = gl_slab_download(ctx, &slab_, host.data())) throw StrError(st, gl_last_error(ctx), gl_last_error_cell(ctx));
        return host;
    }
    gl_slab *raw() { return &slab_; }

  private:
    gl_slab slab_;
};

class Context {
  public:
    explicit Context(int device = 0, void *stream = nullptr) {
        int st = gl_ctx_create(device, stream, &ctx_);
        if (st) throw StrError(st, gl_status_string(st));
    }
    Context(const Context &) = delete;
    ~Context() { gl_ctx_destroy(ctx_); }

    // flattens the AoS Mesh and uploads it (replaces the per-cell Mesh::set_pad, src/mesh/mesh.rs:448-456)
    DeviceMesh upload(const Mesh &mesh) {
        std::vector<double> coords;
        coords.reserve(mesh.points.size() * mesh.ndim);
        for (const auto &p : mesh.points) coords.insert(coords.end(), p.coords.begin(), p.coords.begin() + mesh.ndim);
        std::vector<uint8_t> kinds;
        std::vector<int32_t> markers;
        std::vector<uint64_t> off{0}, conn;
        for (const auto &c : mesh.cells) {
            kinds.push_back((uint8_t)c.kind);
            markers.push_back(c.marker);
            for (size_t p : c.points) conn.push_back((uint64_t)p);
            off.push_back(conn.size());
        }
        gl_mesh *h = nullptr;
        int st = gl_mesh_upload(ctx_, (int)mesh.ndim, mesh.points.size(), coords.data(), mesh.cells.size(), kinds.data(),
                                markers.data(), off.data(), conn.data(), &h);
        check(st);
        return DeviceMesh(ctx_, h);
    }
    // Block::new(box).set_ndiv(ndiv).subdivide(kind) built in HBM (gl_mesh_block): Qua4 / Qua8 / Qua9 / Hex8 / Hex20
    DeviceMesh block(GeoKind kind, const std::vector<uint32_t> &ndiv, const std::vector<double> &xmin, const std::vector<double> &xmax,
                     int marker = 1) {
        uint32_t nd[3] = {1, 1, 1};
        double lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
        for (size_t d = 0; d < ndiv.size() && d < 3; d++) nd[d] = ndiv[d], lo[d] = xmin[d], hi[d] = xmax[d];
        gl_mesh *h = nullptr;
        check(gl_mesh_block(ctx_, (int)kind, nd, lo, hi, marker, &h));
        return DeviceMesh(ctx_, h);
    }
    // ... and of a general block given by its corners or its Qua8 / Hex20 nodes (gl_mesh_block_general; control: ncontrol x ndim)
    DeviceMesh block_general(GeoKind kind, const std::vector<uint32_t> &ndiv, const std::vector<double> &control, int ncontrol, int marker = 1) {
        uint32_t nd[3] = {1, 1, 1};
        for (size_t d = 0; d < ndiv.size() && d < 3; d++) nd[d] = ndiv[d];
        gl_mesh *h = nullptr;
        check(gl_mesh_block_general(ctx_, (int)kind, nd, ncontrol, control.data(), marker, &h));
        return DeviceMesh(ctx_, h);
    }
    // Mesh::read / Mesh::from_text (MSH) and Mesh::read_json behind the C ABI, parsed and uploaded in one call
    DeviceMesh mesh_from_text(const std::string &text, int format = GL_MESH_FORMAT_MSH) {
        gl_mesh *h = nullptr;
        check(gl_mesh_from_text(ctx_, text.data(), text.size(), format, &h));
        return DeviceMesh(ctx_, h);
    }
    void sync() { check(gl_ctx_sync(ctx_)); }
    std::string last_kernel() const { return gl_ctx_last_kernel(ctx_); }
    gl_ctx *handle() const { return ctx_; }
    void check(int st) const {
        if (st) throw StrError(st, gl_last_error(ctx_), gl_last_error_cell(ctx_));
    }

  private:
    gl_ctx *ctx_ = nullptr;
};

namespace integ {

// CommonArgs::new defaults: clear = true, axisymmetric = false, alpha = 1, ii0 = jj0 = 0
struct CommonArgs {
    int ngauss = 0; // 0 = Gauss::new(kind)
    bool clear = true;
    bool axisymmetric = false;
    double alpha = 1.0;
    unsigned ii0 = 0, jj0 = 0, nrow = 0, ncol = 0;
    gl_args c() const {
        gl_args a;
        gl_args_default(&a);
        a.ngauss = ngauss;
        a.clear = clear;
        a.axisymmetric = axisymmetric;
        a.alpha = alpha;
        a.ii0 = ii0;
        a.jj0 = jj0;
        a.nrow = nrow;
        a.ncol = ncol;
        return a;
    }
};

// coefficient data replacing the reference's closures
struct Coef {
    gl_coef_mode mode = GL_COEF_UNIFORM;
    std::vector<double> data;
    std::vector<int32_t> markers;
    static Coef uniform(std::vector<double> record) { return Coef{GL_COEF_UNIFORM, std::move(record), {}}; }
    static Coef per_marker(std::vector<int32_t> m, std::vector<double> records) { return Coef{GL_COEF_PER_MARKER, std::move(records), std::move(m)}; }
    static Coef per_cell(std::vector<double> records) { return Coef{GL_COEF_PER_CELL, std::move(records), {}}; }
    static Coef per_gauss(std::vector<double> records) { return Coef{GL_COEF_PER_GAUSS, std::move(records), {}}; }
};

// LinElasticity::new(E, nu, two_dim, plane_stress).get_modulus().matrix(), column-major
inline std::vector<double> lin_elasticity(double young, double poisson, bool two_dim, bool plane_stress = false) {
    std::vector<double> dd(two_dim ? 16 : 36);
    gl_lin_elasticity(young, poisson, two_dim, plane_stress, dd.data());
    return dd;
}

namespace batch {

inline std::vector<double> run(int op, Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args,
                               const Coef *coef) {
    gl_args a = args.c();
    uint64_t total = 0;
    if (int st = gl_out_total(mesh.handle(), op, cells.first, cells.second, &a, &total)) throw StrError(st, gl_status_string(st));
    std::vector<double> out(total);
    gl_out o{out.data(), GL_HOST, 0, total};
    gl_coef c{};
    if (coef) {
        c.mode = coef->mode;
        c.location = GL_HOST;
        c.data = coef->data.data();
        c.markers = coef->markers.empty() ? nullptr : coef->markers.data();
        c.nrecord = coef->markers.size();
    }
    ctx.check(gl_integ(ctx.handle(), mesh.handle(), op, cells.first, cells.second, &a, coef ? &c : nullptr, &o));
    return out;
}

// element e of the range occupies out[e*size .. (e+1)*size) (homogeneous mesh), column-major, ii = i + m*ndim
inline std::vector<double> mat_10_bdb(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &dd) {
    return run(GL_OP_MAT_10_BDB, ctx, mesh, cells, args, &dd);
}
inline std::vector<double> mat_01_nsn(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &s) {
    return run(GL_OP_MAT_01_NSN, ctx, mesh, cells, args, &s);
}
inline std::vector<double> mat_03_btb(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &tt) {
    return run(GL_OP_MAT_03_BTB, ctx, mesh, cells, args, &tt);
}
inline std::vector<double> mat_02_bvn(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &v) {
    return run(GL_OP_MAT_02_BVN, ctx, mesh, cells, args, &v); // src/integ/mat_02_bvn.rs:48
}
inline std::vector<double> mat_08_ntn(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &tt) {
    return run(GL_OP_MAT_08_NTN, ctx, mesh, cells, args, &tt); // src/integ/mat_08_ntn.rs:56
}
inline std::vector<double> mat_09_nvb(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &v) {
    return run(GL_OP_MAT_09_NVB, ctx, mesh, cells, args, &v); // src/integ/mat_09_nvb.rs:54
}
inline std::vector<double> vec_01_ns(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &s) {
    return run(GL_OP_VEC_01_NS, ctx, mesh, cells, args, &s);
}
inline std::vector<double> vec_02_nv(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &v) {
    return run(GL_OP_VEC_02_NV, ctx, mesh, cells, args, &v);
}
inline std::vector<double> vec_03_bv(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &w) {
    return run(GL_OP_VEC_03_BV, ctx, mesh, cells, args, &w);
}
inline std::vector<double> vec_04_bt(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &sig) {
    return run(GL_OP_VEC_04_BT, ctx, mesh, cells, args, &sig);
}

inline gl_coef to_c(const Coef &coef) {
    gl_coef c{};
    c.mode = coef.mode;
    c.location = GL_HOST;
    c.data = coef.data.data();
    c.markers = coef.markers.empty() ? nullptr : coef.markers.data();
    c.nrecord = coef.markers.size();
    return c;
}

// the same integration with the output LEFT ON THE DEVICE (gl_integ_slab; asynchronous: ctx.sync() reports zero determinants)
inline DeviceSlab integ_device(int op, Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args,
                               const Coef &coef) {
    gl_args a = args.c();
    gl_coef c = to_c(coef);
    DeviceSlab slab;
    ctx.check(gl_integ_slab(ctx.handle(), mesh.handle(), op, cells.first, cells.second, &a, &c, slab.raw()));
    return slab;
}
inline DeviceSlab mat_10_bdb_device(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args, const Coef &dd) {
    return integ_device(GL_OP_MAT_10_BDB, ctx, mesh, cells, args, dd);
}

// several cases from one geometry pass (gl_integ_fused); outputs on the host, one vector per item
inline std::vector<std::vector<double>> fused(Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args,
                                              const std::vector<std::pair<int, const Coef *>> &items) {
    gl_args a = args.c();
    std::vector<std::vector<double>> outs(items.size());
    std::vector<gl_coef> coefs(items.size());
    std::vector<gl_out> gouts(items.size());
    std::vector<gl_fused_item> its(items.size());
    for (size_t i = 0; i < items.size(); i++) {
        uint64_t total = 0;
        if (int st = gl_out_total(mesh.handle(), items[i].first, cells.first, cells.second, &a, &total)) throw StrError(st, gl_status_string(st));
        outs[i].resize(total);
        coefs[i] = to_c(*items[i].second);
        gouts[i] = gl_out{outs[i].data(), GL_HOST, 0, total};
        its[i] = gl_fused_item{items[i].first, 0, &coefs[i], &gouts[i]};
    }
    ctx.check(gl_integ_fused(ctx.handle(), mesh.handle(), cells.first, cells.second, &a, its.data(), (int)its.size()));
    return outs;
}

// planned assembly (gl_assemble_plan_create / gl_assemble_planned): the slot of every element-block entry in the CSR values is
// found once; every later assemble() is one kernel per GeoKind for the stiffness matrix -- the element blocks never reach HBM.
// rowptr / colidx / vals are DEVICE pointers, eq is the HOST point -> equation table.
class AssemblyPlan {
  public:
    AssemblyPlan(int op, Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args,
                 const std::vector<int32_t> &eq, const int64_t *rowptr, const int32_t *colidx)
        : ctx_(ctx) {
        gl_args a = args.c();
        uint64_t missing = 0;
        ctx.check(gl_assemble_plan_create(ctx.handle(), mesh.handle(), op, cells.first, cells.second, &a, eq.data(), GL_HOST, rowptr, colidx,
                                          &missing, &plan_));
    }
    AssemblyPlan(const AssemblyPlan &) = delete;
    ~AssemblyPlan() { gl_assemble_plan_free(ctx_.handle(), plan_); }
    void assemble(const Coef &coef, double *vals) {
        gl_coef c = to_c(coef);
        ctx_.check(gl_assemble_planned(ctx_.handle(), plan_, &c, vals));
    }

  private:
    Context &ctx_;
    gl_asm_plan *plan_ = nullptr;
};

// assembly hand-off (gl_assemble): integrates `op` over the cells and adds the element blocks into the CSR values
// (matrix cases; DEVICE pointers rowptr / colidx / vals) or into the global vector `vals` (vector cases; rowptr = colidx =
// nullptr).  `eq` is the HOST point -> equation table eq[point * dofs_per_node + i]; negative = prescribed dof.
inline void assemble(int op, Context &ctx, const DeviceMesh &mesh, std::pair<size_t, size_t> cells, const CommonArgs &args,
                     const Coef &coef, const std::vector<int32_t> &eq, const int64_t *rowptr, const int32_t *colidx, double *vals) {
    gl_args a = args.c();
    gl_coef c{};
    c.mode = coef.mode;
    c.location = GL_HOST;
    c.data = coef.data.data();
    c.markers = coef.markers.empty() ? nullptr : coef.markers.data();
    c.nrecord = coef.markers.size();
    uint64_t missing = 0;
    ctx.check(gl_assemble(ctx.handle(), mesh.handle(), op, cells.first, cells.second, &a, &c, eq.data(), GL_HOST, rowptr, colidx, vals, &missing));
}

} // namespace batch
} // namespace integ

// Single-process multi-GPU driver (gl_multi_*): one context, stream and host thread per device inside the library; the cell range
// is split into contiguous pieces, host outputs come back concatenated in cell order (bit-identical to one device).
class MultiContext {
  public:
    explicit MultiContext(const std::vector<int> &devices) {
        int st = gl_multi_create(devices.data(), (int)devices.size(), &multi_);
        if (st) throw StrError(st, gl_status_string(st));
    }
    MultiContext(const MultiContext &) = delete;
    ~MultiContext() {
        if (mesh_) gl_multi_mesh_free(multi_, mesh_);
        gl_multi_destroy(multi_);
    }
    int ndevice() const { return gl_multi_ndevice(multi_); }
    void upload(const Mesh &mesh) {
        std::vector<double> coords;
        for (const auto &p : mesh.points) coords.insert(coords.end(), p.coords.begin(), p.coords.begin() + mesh.ndim);
        std::vector<uint8_t> kinds;
        std::vector<int32_t> markers;
        std::vector<uint64_t> off{0}, conn;
        for (const auto &c : mesh.cells) {
            kinds.push_back((uint8_t)c.kind);
            markers.push_back(c.marker);
            for (size_t p : c.points) conn.push_back((uint64_t)p);
            off.push_back(conn.size());
        }
        if (mesh_) gl_multi_mesh_free(multi_, mesh_);
        mesh_ = nullptr;
        check(gl_multi_mesh_upload(multi_, (int)mesh.ndim, mesh.points.size(), coords.data(), mesh.cells.size(), kinds.data(), markers.data(),
                                   off.data(), conn.data(), &mesh_));
    }
    std::vector<double> integ(int op, std::pair<size_t, size_t> cells, const integ::CommonArgs &args, const integ::Coef &coef) {
        gl_args a = args.c();
        gl_coef c = integ::batch::to_c(coef);
        uint64_t total = 0;
        if (int st = gl_out_total(gl_multi_mesh_replica(mesh_, 0), op, cells.first, cells.second, &a, &total)) throw StrError(st, gl_status_string(st));
        std::vector<double> out(total);
        gl_out o{out.data(), GL_HOST, 0, total};
        check(gl_multi_integ(multi_, mesh_, op, cells.first, cells.second, &a, &c, &o));
        return out;
    }
    // one DeviceSlab per device, in piece order (device-resident; synchronised before returning)
    std::vector<DeviceSlab> integ_slabs(int op, std::pair<size_t, size_t> cells, const integ::CommonArgs &args, const integ::Coef &coef) {
        gl_args a = args.c();
        gl_coef c = integ::batch::to_c(coef);
        std::vector<gl_slab> raw(ndevice(), gl_slab{});
        check(gl_multi_integ_slabs(multi_, mesh_, op, cells.first, cells.second, &a, &c, raw.data()));
        check(gl_multi_sync(multi_));
        std::vector<DeviceSlab> out(raw.size());
        for (size_t g = 0; g < raw.size(); g++) *out[g].raw() = raw[g];
        return out;
    }
    gl_ctx *ctx(int g) { return gl_multi_ctx(multi_, g); }

  private:
    void check(int st) const {
        if (st) throw StrError(st, gl_multi_last_error(multi_));
    }
    gl_multi *multi_ = nullptr;
    gl_multi_mesh *mesh_ = nullptr;
};

} // namespace gemlab

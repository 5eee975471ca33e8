//! Raw bindings to the C ABI declared in `include/gemlab_cuda.h` (one `extern "C"` item per `GL_API` symbol).
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)]
pub struct gl_ctx {
    _private: [u8; 0],
}
#[repr(C)]
pub struct gl_mesh {
    _private: [u8; 0],
}

pub const GL_OK: c_int = 0;
pub const GL_ERR_ZERO_DET: c_int = 2;

pub const GL_OP_MAT_10_BDB: c_int = 0;
pub const GL_OP_MAT_01_NSN: c_int = 1;
pub const GL_OP_MAT_03_BTB: c_int = 2;
pub const GL_OP_VEC_01_NS: c_int = 3;
pub const GL_OP_VEC_03_BV: c_int = 4;
pub const GL_OP_VEC_02_NV: c_int = 5;
pub const GL_OP_VEC_04_BT: c_int = 6;
pub const GL_OP_JACOBIAN: c_int = 8;
pub const GL_OP_MAT_02_BVN: c_int = 9;
pub const GL_OP_MAT_08_NTN: c_int = 10;
pub const GL_OP_MAT_09_NVB: c_int = 11;
pub const GL_OP_MAT_04_NSB: c_int = 12;
pub const GL_OP_MAT_05_BTN: c_int = 13;
pub const GL_OP_MAT_06_NVN: c_int = 14;
pub const GL_OP_MAT_07_BSN: c_int = 15;
pub const GL_OP_MAT_01_NSN_BRY: c_int = 16;
pub const GL_OP_VEC_01_NS_BRY: c_int = 17;
pub const GL_OP_VEC_02_NV_BRY: c_int = 18;
pub const GL_OP_NORMAL: c_int = 19;

pub const GL_HOST: i32 = 0;
pub const GL_DEVICE: i32 = 1;

pub const GL_COEF_UNIFORM: i32 = 0;
pub const GL_COEF_PER_MARKER: i32 = 1;
pub const GL_COEF_PER_CELL: i32 = 2;
pub const GL_COEF_PER_GAUSS: i32 = 3;

/// `struct gl_args` -- mirrors `gemlab::integ::CommonArgs` (src/integ/common_args.rs:5-43)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct gl_args {
    pub ngauss: i32,
    pub axisymmetric: i32,
    pub clear: i32,
    /// 1 = element matrices leave as packed upper triangles (batched boundary only)
    pub packed: i32,
    pub alpha: c_double,
    pub ii0: u32,
    pub jj0: u32,
    pub nrow: u32,
    pub ncol: u32,
    pub ksi: [c_double; 3],
    /// two-pad cases (mat_04_nsb .. mat_07_bsn): GeoKind ordinal of the second Scratchpad; -1 otherwise
    pub kind_b: i32,
    pub reserved: i32,
}

#[repr(C)]
pub struct gl_coef {
    pub mode: i32,
    pub location: i32,
    pub data: *const c_double,
    pub markers: *const i32,
    pub nrecord: u64,
}

#[repr(C)]
pub struct gl_out {
    pub ptr: *mut c_double,
    pub location: i32,
    pub reserved: i32,
    pub capacity: u64,
}

/// `struct gl_slab` -- a batched output that stays on its device
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct gl_slab {
    pub ptr: *mut c_double,
    pub capacity: u64,
    pub size: u64,
    pub cell_begin: u64,
    pub cell_end: u64,
    pub device: i32,
    pub owned: i32,
}

#[repr(C)]
pub struct gl_host_mesh {
    _private: [u8; 0],
}
pub const GL_MESH_FORMAT_MSH: c_int = 0;
pub const GL_MESH_FORMAT_JSON: c_int = 1;
#[repr(C)]
pub struct gl_asm_plan {
    _private: [u8; 0],
}
#[repr(C)]
pub struct gl_multi {
    _private: [u8; 0],
}
#[repr(C)]
pub struct gl_multi_mesh {
    _private: [u8; 0],
}

#[repr(C)]
pub struct gl_fused_item {
    pub op: i32,
    pub reserved: i32,
    pub coef: *const gl_coef,
    pub out: *mut gl_out,
}

extern "C" {
    pub fn gl_integ_fused(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, items: *const gl_fused_item, nitems: c_int) -> c_int;
    pub fn gl_ctx_create(device: c_int, stream: *mut c_void, ctx: *mut *mut gl_ctx) -> c_int;
    pub fn gl_ctx_destroy(ctx: *mut gl_ctx);
    pub fn gl_ctx_set_stream(ctx: *mut gl_ctx, stream: *mut c_void) -> c_int;
    pub fn gl_ctx_sync(ctx: *mut gl_ctx) -> c_int;
    pub fn gl_last_error(ctx: *const gl_ctx) -> *const c_char;
    pub fn gl_last_error_cell(ctx: *const gl_ctx) -> u64;
    pub fn gl_status_string(status: c_int) -> *const c_char;
    pub fn gl_ctx_launch_count(ctx: *const gl_ctx) -> u64;
    pub fn gl_ctx_last_kernel(ctx: *const gl_ctx) -> *const c_char;

    pub fn gl_kind_nnode(kind: c_int) -> c_int;
    pub fn gl_kind_geo_ndim(kind: c_int) -> c_int;
    pub fn gl_kind_class(kind: c_int) -> c_int;
    pub fn gl_kind_from_name(name: *const c_char) -> c_int;
    pub fn gl_kind_name(kind: c_int) -> *const c_char;
    pub fn gl_kind_supported(kind: c_int) -> c_int;
    pub fn gl_gauss_rule(kind: c_int, ngauss: c_int, npoint: *mut c_int, data: *mut *const c_double) -> c_int;
    pub fn gl_shape_eval(kind: c_int, ksi: *const c_double, nn: *mut c_double, ll: *mut c_double) -> c_int;
    pub fn gl_kind_reference_coords(kind: c_int, m: c_int, ksi: *mut c_double) -> c_int;
    pub fn gl_extrap_matrix(kind: c_int, ngauss: c_int, ee: *mut c_double) -> c_int;

    pub fn gl_mesh_upload(
        ctx: *mut gl_ctx,
        ndim: c_int,
        npoint: u64,
        coords: *const c_double,
        ncell: u64,
        kinds: *const u8,
        markers: *const i32,
        conn_off: *const u64,
        conn: *const u64,
        mesh: *mut *mut gl_mesh,
    ) -> c_int;
    pub fn gl_mesh_upload_homogeneous(
        ctx: *mut gl_ctx,
        ndim: c_int,
        npoint: u64,
        coords: *const c_double,
        ncell: u64,
        kind: c_int,
        markers: *const i32,
        conn: *const u32,
        mesh: *mut *mut gl_mesh,
    ) -> c_int;
    pub fn gl_mesh_update_coords(ctx: *mut gl_ctx, mesh: *mut gl_mesh, coords: *const c_double, location: c_int) -> c_int;
    pub fn gl_mesh_free(ctx: *mut gl_ctx, mesh: *mut gl_mesh);
    pub fn gl_mesh_ncell(mesh: *const gl_mesh) -> u64;
    pub fn gl_mesh_npoint(mesh: *const gl_mesh) -> u64;
    pub fn gl_mesh_ndim(mesh: *const gl_mesh) -> c_int;

    pub fn gl_op_out_size(op: c_int, kind: c_int, ndim: c_int) -> u64;
    pub fn gl_op_coef_size(op: c_int, ndim: c_int) -> u64;
    pub fn gl_out_total(mesh: *const gl_mesh, op: c_int, cell_begin: u64, cell_end: u64, args: *const gl_args, total: *mut u64) -> c_int;
    pub fn gl_out_offsets(mesh: *const gl_mesh, op: c_int, cell_begin: u64, cell_end: u64, args: *const gl_args, off: *mut u64) -> c_int;

    pub fn gl_integ(
        ctx: *mut gl_ctx,
        mesh: *const gl_mesh,
        op: c_int,
        cell_begin: u64,
        cell_end: u64,
        args: *const gl_args,
        coef: *const gl_coef,
        out: *mut gl_out,
    ) -> c_int;
    pub fn gl_mat_10_bdb(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_01_nsn(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_02_bvn(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_08_ntn(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_09_nvb(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_03_btb(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_01_ns(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_02_nv(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_03_bv(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_04_bt(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_jacobian(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, out: *mut gl_out) -> c_int;
    pub fn gl_mat_04_nsb(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_05_btn(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_06_nvn(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_07_bsn(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_mat_01_nsn_bry(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_01_ns_bry(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_02_nv_bry(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_normal_vectors(ctx: *mut gl_ctx, mesh: *const gl_mesh, cell_begin: u64, cell_end: u64, args: *const gl_args, out: *mut gl_out) -> c_int;
    pub fn gl_mat_10_gdg(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_vec_10_gs(ctx: *mut gl_ctx, mesh: *const gl_mesh, b: u64, e: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;

    pub fn gl_args_default(args: *mut gl_args);
    pub fn gl_lin_elasticity(young: c_double, poisson: c_double, two_dim: c_int, plane_stress: c_int, dd: *mut c_double);
    pub fn gl_bench_fp64_peak(ctx: *mut gl_ctx, kind: c_int, iters: c_int, tflops: *mut c_double) -> c_int;
    pub fn gl_mesh_block(ctx: *mut gl_ctx, kind: c_int, ndiv: *const u32, xmin: *const c_double, xmax: *const c_double, marker: i32, mesh: *mut *mut gl_mesh) -> c_int;
    pub fn gl_mesh_block_general(ctx: *mut gl_ctx, kind: c_int, ndiv: *const u32, ncontrol: c_int, control: *const c_double, marker: i32, mesh: *mut *mut gl_mesh) -> c_int;
    pub fn gl_mesh_download(ctx: *mut gl_ctx, mesh: *const gl_mesh, coords: *mut c_double, conn: *mut u32) -> c_int;
    pub fn gl_bench_hbm_write(ctx: *mut gl_ctx, kind: c_int, dev: *mut c_double, nbytes: u64, warps_per_sm: c_int, gbs: *mut c_double) -> c_int;

    // assembly hand-off (element blocks -> global sparse matrix / vector on the device)
    pub fn gl_op_dofs_per_node(op: c_int, ndim: c_int) -> c_int;
    pub fn gl_triplet_indices(
        ctx: *mut gl_ctx,
        mesh: *const gl_mesh,
        op: c_int,
        cell_begin: u64,
        cell_end: u64,
        args: *const gl_args,
        eq: *const i32,
        eq_location: c_int,
        rows: *mut i32,
        cols: *mut i32,
    ) -> c_int;
    pub fn gl_assemble(
        ctx: *mut gl_ctx,
        mesh: *const gl_mesh,
        op: c_int,
        cell_begin: u64,
        cell_end: u64,
        args: *const gl_args,
        coef: *const gl_coef,
        eq: *const i32,
        eq_location: c_int,
        rowptr: *const i64,
        colidx: *const i32,
        vals: *mut c_double,
        missing: *mut u64,
    ) -> c_int;

    // device-resident output slabs
    pub fn gl_integ_slab(ctx: *mut gl_ctx, mesh: *const gl_mesh, op: c_int, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, slab: *mut gl_slab) -> c_int;
    pub fn gl_slab_download(ctx: *mut gl_ctx, slab: *const gl_slab, host: *mut c_double) -> c_int;
    pub fn gl_slab_free(slab: *mut gl_slab);

    // single-process multi-GPU driver
    pub fn gl_multi_create(devices: *const c_int, ndevice: c_int, multi: *mut *mut gl_multi) -> c_int;
    pub fn gl_multi_destroy(multi: *mut gl_multi);
    pub fn gl_multi_ndevice(multi: *const gl_multi) -> c_int;
    pub fn gl_multi_ctx(multi: *mut gl_multi, g: c_int) -> *mut gl_ctx;
    pub fn gl_multi_last_error(multi: *const gl_multi) -> *const c_char;
    pub fn gl_multi_partition(cell_begin: u64, cell_end: u64, ndevice: c_int, g: c_int, piece_begin: *mut u64, piece_end: *mut u64);
    pub fn gl_multi_mesh_upload(
        multi: *mut gl_multi,
        ndim: c_int,
        npoint: u64,
        coords: *const c_double,
        ncell: u64,
        kinds: *const u8,
        markers: *const i32,
        conn_off: *const u64,
        conn: *const u64,
        mesh: *mut *mut gl_multi_mesh,
    ) -> c_int;
    pub fn gl_multi_mesh_free(multi: *mut gl_multi, mesh: *mut gl_multi_mesh);
    pub fn gl_multi_mesh_replica(mesh: *const gl_multi_mesh, g: c_int) -> *const gl_mesh;
    pub fn gl_multi_integ(multi: *mut gl_multi, mesh: *const gl_multi_mesh, op: c_int, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, out: *mut gl_out) -> c_int;
    pub fn gl_multi_integ_slabs(multi: *mut gl_multi, mesh: *const gl_multi_mesh, op: c_int, cell_begin: u64, cell_end: u64, args: *const gl_args, coef: *const gl_coef, slabs: *mut gl_slab) -> c_int;
    pub fn gl_multi_sync(multi: *mut gl_multi) -> c_int;

    // planned assembly (position map computed once; stiffness blocks go from shared memory straight into the CSR values)
    pub fn gl_assemble_plan_create(
        ctx: *mut gl_ctx,
        mesh: *const gl_mesh,
        op: c_int,
        cell_begin: u64,
        cell_end: u64,
        args: *const gl_args,
        eq: *const i32,
        eq_location: c_int,
        rowptr: *const i64,
        colidx: *const i32,
        missing: *mut u64,
        plan: *mut *mut gl_asm_plan,
    ) -> c_int;
    pub fn gl_assemble_planned(ctx: *mut gl_ctx, plan: *const gl_asm_plan, coef: *const gl_coef, vals: *mut c_double) -> c_int;
    pub fn gl_assemble_plan_free(ctx: *mut gl_ctx, plan: *mut gl_asm_plan);

    // mesh ingestion: gemlab's MSH text and JSON formats, parsed on the host (the reference's grammar and error strings)
    pub fn gl_host_mesh_parse(text: *const c_char, len: u64, format: c_int, mesh: *mut *mut gl_host_mesh) -> c_int;
    pub fn gl_host_mesh_read(path: *const c_char, format: c_int, mesh: *mut *mut gl_host_mesh) -> c_int;
    pub fn gl_host_mesh_free(mesh: *mut gl_host_mesh);
    pub fn gl_mesh_text_last_error() -> *const c_char;
    pub fn gl_host_mesh_ndim(mesh: *const gl_host_mesh) -> c_int;
    pub fn gl_host_mesh_npoint(mesh: *const gl_host_mesh) -> u64;
    pub fn gl_host_mesh_ncell(mesh: *const gl_host_mesh) -> u64;
    pub fn gl_host_mesh_nmarked_edge(mesh: *const gl_host_mesh) -> u64;
    pub fn gl_host_mesh_nmarked_face(mesh: *const gl_host_mesh) -> u64;
    pub fn gl_host_mesh_coords(mesh: *const gl_host_mesh) -> *const c_double;
    pub fn gl_host_mesh_point_markers(mesh: *const gl_host_mesh) -> *const i32;
    pub fn gl_host_mesh_kinds(mesh: *const gl_host_mesh) -> *const u8;
    pub fn gl_host_mesh_markers(mesh: *const gl_host_mesh) -> *const i32;
    pub fn gl_host_mesh_conn_off(mesh: *const gl_host_mesh) -> *const u64;
    pub fn gl_host_mesh_conn(mesh: *const gl_host_mesh) -> *const u64;
    pub fn gl_host_mesh_marked_edges(mesh: *const gl_host_mesh) -> *const i64;
    pub fn gl_host_mesh_marked_faces(mesh: *const gl_host_mesh) -> *const i64;
    pub fn gl_mesh_upload_host(ctx: *mut gl_ctx, host: *const gl_host_mesh, mesh: *mut *mut gl_mesh) -> c_int;
    pub fn gl_mesh_from_text(ctx: *mut gl_ctx, text: *const c_char, len: u64, format: c_int, mesh: *mut *mut gl_mesh) -> c_int;
}

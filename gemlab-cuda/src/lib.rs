//! `gemlab-cuda`: batched element integration for gemlab meshes on NVIDIA B200 (sm_100a).
//!
//! gemlab integrates ONE cell per call (`integ::mat_10_bdb(&mut kk, &mut args, |dd,p,N,B| ..)`, src/integ/mat_10_bdb.rs:121)
//! inside a user loop `for cell in &mesh.cells { mesh.set_pad(&mut pad, &cell.points); integ::X(..)?; }`.
//! This crate keeps gemlab's `Mesh`, `GeoKind`, `Gauss`, `Scratchpad` and `integ::*` untouched and adds
//! `integ::batch::*`, which takes the `Mesh` plus a cell range and returns all element matrices/vectors at once:
//!
//! ```ignore
//! use gemlab::prelude::*;
//! use gemlab_cuda::{integ::batch, BatchArgs, Coef, Context};
//! use russell_tensor::LinElasticity;
//!
//! let mesh = Block::new_cube(1.0).set_ndiv(&[200, 200, 200])?.subdivide(GeoKind::Hex8)?;
//! let ctx = Context::new(0)?;                      // one per GPU
//! let model = LinElasticity::new(480.0, 1.0 / 3.0, false, false);
//! let dd = Coef::uniform_tensor4(model.get_modulus());
//! // the one-call form north_star words: &Mesh + cell range (uploads, integrates, frees)
//! let kk = batch::for_mesh::mat_10_bdb(&ctx, &mesh, 0..mesh.cells.len(), &BatchArgs::new(), &dd)?;
//! // or keep the mesh on the device across calls (a Newton loop):
//! let dmesh = ctx.upload(&mesh)?;                  // SoA device mesh (replaces Mesh::set_pad per cell)
//! // element e occupies kk[e*576 .. (e+1)*576], column-major 24x24, row ii = i + m*3, col jj = j + n*3
//! let kk = batch::mat_10_bdb(&ctx, &dmesh, 0..mesh.cells.len(), &BatchArgs::new(), &dd)?;
//! // ... or leave the element matrices on the device (DeviceSlab) / assemble them there (AssemblyPlan)
//! let slab = batch::device::mat_10_bdb(&ctx, &dmesh, 0..mesh.cells.len(), &BatchArgs::new(), &dd)?;
//! ```
//!
//! Multi-GPU: `MultiContext::new(&[0, 1, ..])` owns one context, stream and host thread per device inside the library and splits
//! the cell range into the contiguous pieces `partition(ncell, G, g)`; host outputs come back concatenated in cell order
//! (bit-identical to one device); there is no collective on the data path.
//!
//! NOTE: shipped as source; not compiled in the authoring image (no cargo). The ABI is defined by
//! `include/gemlab_cuda.h`; see INTEGRATION.md.

pub mod ffi;

use gemlab::mesh::Mesh;
use gemlab::shapes::GeoKind;
use gemlab::StrError;
use russell_tensor::Tensor4;
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::ops::Range;
use std::ptr;
use std::sync::{Mutex, OnceLock};

/// Mirrors `gemlab::integ::CommonArgs` without `pad`/`gauss` (implied by the mesh and `ngauss`).
#[derive(Clone, Copy, Debug)]
pub struct BatchArgs {
    /// None = `Gauss::new(kind)`; Some(n) = `Gauss::new_sized(kind.class(), n)`
    pub ngauss: Option<usize>,
    pub clear: bool,
    pub axisymmetric: bool,
    pub alpha: f64,
    pub ii0: usize,
    pub jj0: usize,
    /// per-element block dims; 0 = tight
    pub nrow: usize,
    pub ncol: usize,
    /// element matrices leave as packed upper triangles (n(n+1)/2 doubles; K(i,j), i <= j, at j(j+1)/2 + i)
    pub packed: bool,
    /// two-pad cases (`mat_04_nsb` .. `mat_07_bsn`): the `GeoKind` of the second `Scratchpad` (its nodes are the cell's first ones)
    pub kind_b: Option<GeoKind>,
}

impl BatchArgs {
    /// Same defaults as `CommonArgs::new`: clear = true, axisymmetric = false, alpha = 1, ii0 = jj0 = 0
    pub fn new() -> Self {
        BatchArgs { ngauss: None, clear: true, axisymmetric: false, alpha: 1.0, ii0: 0, jj0: 0, nrow: 0, ncol: 0, packed: false, kind_b: None }
    }
    fn to_c(&self) -> ffi::gl_args {
        ffi::gl_args {
            ngauss: self.ngauss.unwrap_or(0) as i32,
            axisymmetric: self.axisymmetric as i32,
            clear: self.clear as i32,
            packed: self.packed as i32,
            alpha: self.alpha,
            ii0: self.ii0 as u32,
            jj0: self.jj0 as u32,
            nrow: self.nrow as u32,
            ncol: self.ncol as u32,
            ksi: [0.0; 3],
            kind_b: self.kind_b.map_or(-1, |k| k as i32),
            reserved: 0,
        }
    }
}

/// Coefficient data replacing gemlab's per-Gauss-point closures (`fn_dd`, `fn_s`, `fn_tt`, `fn_w`, ...).
pub enum Coef {
    Uniform(Vec<f64>),
    PerMarker { markers: Vec<i32>, records: Vec<f64> },
    PerCell(Vec<f64>),
    /// record `(e - cell_begin) * ngauss + p` for Gauss point `p` of cell `e`: what the closure would have returned there
    PerGauss(Vec<f64>),
}

impl Coef {
    /// `dd.set_tensor(1.0, model.get_modulus())` for every Gauss point: the Mandel matrix, column-major as stored by
    /// russell_lab::Matrix
    pub fn uniform_tensor4(dd: &Tensor4) -> Self {
        Coef::Uniform(dd.matrix().as_data().clone())
    }
    fn to_c(&self) -> ffi::gl_coef {
        match self {
            Coef::Uniform(d) => ffi::gl_coef { mode: ffi::GL_COEF_UNIFORM, location: ffi::GL_HOST, data: d.as_ptr(), markers: ptr::null(), nrecord: 1 },
            Coef::PerMarker { markers, records } => ffi::gl_coef {
                mode: ffi::GL_COEF_PER_MARKER,
                location: ffi::GL_HOST,
                data: records.as_ptr(),
                markers: markers.as_ptr(),
                nrecord: markers.len() as u64,
            },
            Coef::PerCell(d) => ffi::gl_coef { mode: ffi::GL_COEF_PER_CELL, location: ffi::GL_HOST, data: d.as_ptr(), markers: ptr::null(), nrecord: 0 },
            Coef::PerGauss(d) => ffi::gl_coef { mode: ffi::GL_COEF_PER_GAUSS, location: ffi::GL_HOST, data: d.as_ptr(), markers: ptr::null(), nrecord: 0 },
        }
    }
}

/// Integration cases (`gl_op`); the numbering follows the reference's file names (src/integ/mod.rs:304-353)
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum Op {
    Mat10Bdb = 0,
    Mat01Nsn = 1,
    Mat03Btb = 2,
    Vec01Ns = 3,
    Vec03Bv = 4,
    Vec02Nv = 5,
    Vec04Bt = 6,
    Mat02Bvn = 9,
    Mat08Ntn = 10,
    Mat09Nvb = 11,
    Mat04Nsb = 12,
    Mat05Btn = 13,
    Mat06Nvn = 14,
    Mat07Bsn = 15,
    Mat01NsnBry = 16,
    Vec01NsBry = 17,
    Vec02NvBry = 18,
}

// ---- errors ------------------------------------------------------------------------------------------------------------------
// gemlab's StrError is `&'static str`.  The library hands out the reference's own strings; each DISTINCT message is interned once
// (a bounded set: the reference's error strings), so repeated failures do not leak.
fn intern(msg: &str) -> StrError {
    static TABLE: OnceLock<Mutex<HashMap<String, &'static str>>> = OnceLock::new();
    let mut table = TABLE.get_or_init(|| Mutex::new(HashMap::new())).lock().unwrap();
    if let Some(s) = table.get(msg) {
        return *s;
    }
    let leaked: &'static str = Box::leak(msg.to_string().into_boxed_str());
    table.insert(msg.to_string(), leaked);
    leaked
}

fn static_error(status: i32) -> StrError {
    let s = unsafe { CStr::from_ptr(ffi::gl_status_string(status)) };
    intern(s.to_str().unwrap_or("libgemlab_cuda error"))
}

fn ctx_error(ctx: *const ffi::gl_ctx, status: i32) -> StrError {
    let s = unsafe { CStr::from_ptr(ffi::gl_last_error(ctx)) };
    match s.to_str() {
        Ok(v) if !v.is_empty() => intern(v),
        _ => static_error(status),
    }
}

// ---- context, meshes --------------------------------------------------------------------------------------------------------
pub struct Context {
    ctx: *mut ffi::gl_ctx,
}

// a context may move between threads but must not be shared (gemlab's &mut Scratchpad discipline)
unsafe impl Send for Context {}

pub struct DeviceMesh<'a> {
    ctx: &'a Context,
    mesh: *mut ffi::gl_mesh,
    pub ncell: usize,
}

/// Flat arrays of a `Mesh` in the layout `gl_mesh_upload` takes
struct FlatMesh {
    ndim: usize,
    coords: Vec<f64>,
    kinds: Vec<u8>,
    markers: Vec<i32>,
    conn_off: Vec<u64>,
    conn: Vec<u64>,
}

fn flatten(mesh: &Mesh) -> FlatMesh {
    let ndim = mesh.ndim;
    let mut coords = Vec::with_capacity(mesh.points.len() * ndim);
    for p in &mesh.points {
        coords.extend_from_slice(&p.coords[..ndim]);
    }
    let mut kinds = Vec::with_capacity(mesh.cells.len());
    let mut markers = Vec::with_capacity(mesh.cells.len());
    let mut conn_off = Vec::with_capacity(mesh.cells.len() + 1);
    let mut conn: Vec<u64> = Vec::new();
    conn_off.push(0u64);
    for c in &mesh.cells {
        kinds.push(c.kind as u8); // GeoKind::VALUES order
        markers.push(c.marker as i32);
        conn.extend(c.points.iter().map(|&p| p as u64));
        conn_off.push(conn.len() as u64);
    }
    FlatMesh { ndim, coords, kinds, markers, conn_off, conn }
}

impl Context {
    pub fn new(device: i32) -> Result<Self, StrError> {
        let mut ctx = ptr::null_mut();
        let st = unsafe { ffi::gl_ctx_create(device, ptr::null_mut(), &mut ctx) };
        if st != 0 {
            return Err(static_error(st));
        }
        Ok(Context { ctx })
    }

    fn check(&self, st: i32) -> Result<(), StrError> {
        if st != 0 {
            Err(ctx_error(self.ctx, st))
        } else {
            Ok(())
        }
    }

    /// Flattens `Mesh.points` / `Mesh.cells` and uploads them (the library transposes to structure-of-arrays and
    /// buckets cells by GeoKind). Replaces the per-cell `Mesh::set_pad` (src/mesh/mesh.rs:448-456).
    pub fn upload<'a>(&'a self, mesh: &Mesh) -> Result<DeviceMesh<'a>, StrError> {
        let f = flatten(mesh);
        let mut handle = ptr::null_mut();
        let st = unsafe {
            ffi::gl_mesh_upload(
                self.ctx,
                f.ndim as i32,
                mesh.points.len() as u64,
                f.coords.as_ptr(),
                mesh.cells.len() as u64,
                f.kinds.as_ptr(),
                f.markers.as_ptr(),
                f.conn_off.as_ptr(),
                f.conn.as_ptr(),
                &mut handle,
            )
        };
        self.check(st)?;
        Ok(DeviceMesh { ctx: self, mesh: handle, ncell: mesh.cells.len() })
    }

    /// `Block::new(box).set_ndiv(ndiv).subdivide(kind)` for an axis-aligned box, generated directly on the device with the
    /// reference's cell order and first-touch point numbering (`gl_mesh_block`; Qua4 / Qua8 / Qua9 / Hex8 / Hex20 targets).
    pub fn block<'a>(&'a self, kind: GeoKind, ndiv: &[usize], xmin: &[f64], xmax: &[f64], marker: i32) -> Result<DeviceMesh<'a>, StrError> {
        let mut nd = [1u32; 3];
        let (mut lo, mut hi) = ([0.0f64; 3], [0.0f64; 3]);
        for d in 0..ndiv.len().min(3) {
            nd[d] = ndiv[d] as u32;
            lo[d] = xmin[d];
            hi[d] = xmax[d];
        }
        let mut handle = ptr::null_mut();
        let st = unsafe { ffi::gl_mesh_block(self.ctx, kind as i32, nd.as_ptr(), lo.as_ptr(), hi.as_ptr(), marker, &mut handle) };
        self.check(st)?;
        let ncell = unsafe { ffi::gl_mesh_ncell(handle) } as usize;
        Ok(DeviceMesh { ctx: self, mesh: handle, ncell })
    }

    /// The same for a general block: `control` holds its 4 corners or 8 Qua8 nodes (2D), 8 corners or 20 Hex20 nodes (3D),
    /// `ncontrol x ndim` row-major (`gl_mesh_block_general`; `Block::subdivide`, src/mesh/block.rs:661-812).
    pub fn block_general<'a>(&'a self, kind: GeoKind, ndiv: &[usize], control: &[f64], ncontrol: usize, marker: i32) -> Result<DeviceMesh<'a>, StrError> {
        let mut nd = [1u32; 3];
        for d in 0..ndiv.len().min(3) {
            nd[d] = ndiv[d] as u32;
        }
        let mut handle = ptr::null_mut();
        let st = unsafe { ffi::gl_mesh_block_general(self.ctx, kind as i32, nd.as_ptr(), ncontrol as i32, control.as_ptr(), marker, &mut handle) };
        self.check(st)?;
        let ncell = unsafe { ffi::gl_mesh_ncell(handle) } as usize;
        Ok(DeviceMesh { ctx: self, mesh: handle, ncell })
    }

    /// `Mesh::from_text` (MSH, `json = false`) or the JSON image of `Mesh` (`Mesh::read_json`, `json = true`) parsed by the library
    /// and uploaded in one call (`gl_mesh_from_text`); errors are the reference's strings.
    pub fn mesh_from_text<'a>(&'a self, text: &str, json: bool) -> Result<DeviceMesh<'a>, StrError> {
        let c = CString::new(text).map_err(|_| "text contains a NUL byte")?;
        let mut handle = ptr::null_mut();
        let format = if json { ffi::GL_MESH_FORMAT_JSON } else { ffi::GL_MESH_FORMAT_MSH };
        let st = unsafe { ffi::gl_mesh_from_text(self.ctx, c.as_ptr(), text.len() as u64, format, &mut handle) };
        self.check(st)?;
        let ncell = unsafe { ffi::gl_mesh_ncell(handle) } as usize;
        Ok(DeviceMesh { ctx: self, mesh: handle, ncell })
    }

    /// Copies a single-kind device mesh back: (coords `npoint x ndim` row-major, connectivity `ncell x nnode` row-major)
    pub fn download(&self, mesh: &DeviceMesh, nnode: usize) -> Result<(Vec<f64>, Vec<u32>), StrError> {
        let npoint = unsafe { ffi::gl_mesh_npoint(mesh.mesh) } as usize;
        let ndim = unsafe { ffi::gl_mesh_ndim(mesh.mesh) } as usize;
        let mut coords = vec![0.0f64; npoint * ndim];
        let mut conn = vec![0u32; mesh.ncell * nnode];
        let st = unsafe { ffi::gl_mesh_download(self.ctx, mesh.mesh, coords.as_mut_ptr(), conn.as_mut_ptr()) };
        self.check(st)?;
        Ok((coords, conn))
    }

    pub fn sync(&self) -> Result<(), StrError> {
        let st = unsafe { ffi::gl_ctx_sync(self.ctx) };
        self.check(st)
    }

    /// Smallest id of a cell with a zero Jacobian determinant found by the last failing call
    pub fn last_error_cell(&self) -> Option<usize> {
        let c = unsafe { ffi::gl_last_error_cell(self.ctx) };
        if c == u64::MAX {
            None
        } else {
            Some(c as usize)
        }
    }

    /// Which kernel the most recent integration launch used ("hex8_bdb<pat=2>", "syrk_dmma<nn=20,pat=2>", ...)
    pub fn last_kernel(&self) -> String {
        unsafe { CStr::from_ptr(ffi::gl_ctx_last_kernel(self.ctx)) }.to_string_lossy().into_owned()
    }
}

impl Drop for Context {
    fn drop(&mut self) {
        unsafe { ffi::gl_ctx_destroy(self.ctx) }
    }
}

impl<'a> Drop for DeviceMesh<'a> {
    fn drop(&mut self) {
        unsafe { ffi::gl_mesh_free(self.ctx.ctx, self.mesh) }
    }
}

/// The batched output of one call, resident on its device (`gl_slab`): element `e` of `cells` starts at
/// `ptr + off[e - cells.start]`, `off` from `offsets()` (homogeneous meshes: `(e - cells.start) * block`).  Hand `ptr` to a sparse
/// solver, to `batch::triplet_indices` (the slab IS the COO value array) or to your own kernels; `download` copies it to the host.
pub struct DeviceSlab {
    slab: ffi::gl_slab,
}

impl DeviceSlab {
    fn empty() -> Self {
        DeviceSlab { slab: ffi::gl_slab { ptr: ptr::null_mut(), capacity: 0, size: 0, cell_begin: 0, cell_end: 0, device: 0, owned: 0 } }
    }
    /// Device pointer of the first element block
    pub fn ptr(&self) -> *mut f64 {
        self.slab.ptr
    }
    pub fn len(&self) -> usize {
        self.slab.size as usize
    }
    pub fn is_empty(&self) -> bool {
        self.slab.size == 0
    }
    pub fn device(&self) -> i32 {
        self.slab.device
    }
    pub fn cells(&self) -> Range<usize> {
        self.slab.cell_begin as usize..self.slab.cell_end as usize
    }
    pub fn download(&self, ctx: &Context) -> Result<Vec<f64>, StrError> {
        let mut host = vec![0.0f64; self.len()];
        let st = unsafe { ffi::gl_slab_download(ctx.ctx, &self.slab, host.as_mut_ptr()) };
        ctx.check(st)?;
        Ok(host)
    }
}

impl Drop for DeviceSlab {
    fn drop(&mut self) {
        unsafe { ffi::gl_slab_free(&mut self.slab) }
    }
}

unsafe impl Send for DeviceSlab {}

/// Contiguous cell range of rank `g` out of `world`: [floor(g*ncell/G), floor((g+1)*ncell/G))
pub fn partition(ncell: usize, world: usize, rank: usize) -> Range<usize> {
    (rank * ncell) / world..((rank + 1) * ncell) / world
}

// ---- single-process multi-GPU driver ------------------------------------------------------------------------------------------
/// One context + stream + host thread per device inside the library (`gl_multi_*`).  The mesh is replicated, the cell range is
/// split into `partition` pieces, and there is no collective: cells are independent.
pub struct MultiContext {
    multi: *mut ffi::gl_multi,
}

unsafe impl Send for MultiContext {}

pub struct MultiMesh<'a> {
    multi: &'a MultiContext,
    mesh: *mut ffi::gl_multi_mesh,
    pub ncell: usize,
}

impl MultiContext {
    /// `devices` may name a device more than once (every entry gets its own context and stream)
    pub fn new(devices: &[i32]) -> Result<Self, StrError> {
        let mut multi = ptr::null_mut();
        let st = unsafe { ffi::gl_multi_create(devices.as_ptr(), devices.len() as i32, &mut multi) };
        if st != 0 {
            return Err(static_error(st));
        }
        Ok(MultiContext { multi })
    }
    pub fn ndevice(&self) -> usize {
        unsafe { ffi::gl_multi_ndevice(self.multi) as usize }
    }
    fn check(&self, st: i32) -> Result<(), StrError> {
        if st == 0 {
            return Ok(());
        }
        let s = unsafe { CStr::from_ptr(ffi::gl_multi_last_error(self.multi)) };
        match s.to_str() {
            Ok(v) if !v.is_empty() => Err(intern(v)),
            _ => Err(static_error(st)),
        }
    }
    pub fn upload<'a>(&'a self, mesh: &Mesh) -> Result<MultiMesh<'a>, StrError> {
        let f = flatten(mesh);
        let mut handle = ptr::null_mut();
        let st = unsafe {
            ffi::gl_multi_mesh_upload(
                self.multi,
                f.ndim as i32,
                mesh.points.len() as u64,
                f.coords.as_ptr(),
                mesh.cells.len() as u64,
                f.kinds.as_ptr(),
                f.markers.as_ptr(),
                f.conn_off.as_ptr(),
                f.conn.as_ptr(),
                &mut handle,
            )
        };
        self.check(st)?;
        Ok(MultiMesh { multi: self, mesh: handle, ncell: mesh.cells.len() })
    }
    /// Host output of the whole range, pieces concatenated in cell order: bit-identical to the single-device call
    pub fn integ(&self, op: Op, mesh: &MultiMesh, cells: Range<usize>, args: &BatchArgs, coef: &Coef) -> Result<Vec<f64>, StrError> {
        let a = args.to_c();
        let c = coef.to_c();
        let (b, e) = (cells.start as u64, cells.end as u64);
        let mut total = 0u64;
        let st = unsafe { ffi::gl_out_total(ffi::gl_multi_mesh_replica(mesh.mesh, 0), op as i32, b, e, &a, &mut total) };
        if st != 0 {
            return Err(static_error(st));
        }
        let mut out = vec![0.0f64; total as usize];
        let mut o = ffi::gl_out { ptr: out.as_mut_ptr(), location: ffi::GL_HOST, reserved: 0, capacity: total };
        let st = unsafe { ffi::gl_multi_integ(self.multi, mesh.mesh, op as i32, b, e, &a, &c, &mut o) };
        self.check(st)?;
        Ok(out)
    }
    /// One `DeviceSlab` per device, in piece order; returns after every device has finished
    pub fn integ_slabs(&self, op: Op, mesh: &MultiMesh, cells: Range<usize>, args: &BatchArgs, coef: &Coef) -> Result<Vec<DeviceSlab>, StrError> {
        let a = args.to_c();
        let c = coef.to_c();
        let mut slabs: Vec<DeviceSlab> = (0..self.ndevice()).map(|_| DeviceSlab::empty()).collect();
        // gl_slab is repr(C) and DeviceSlab is a transparent wrapper around one: pass a contiguous array of raw slabs
        let mut raw: Vec<ffi::gl_slab> = slabs.iter().map(|s| s.slab).collect();
        let st = unsafe { ffi::gl_multi_integ_slabs(self.multi, mesh.mesh, op as i32, cells.start as u64, cells.end as u64, &a, &c, raw.as_mut_ptr()) };
        for (s, r) in slabs.iter_mut().zip(raw.iter()) {
            s.slab = *r; // ownership of the library-allocated buffers moves into the DeviceSlabs (freed on drop)
        }
        self.check(st)?;
        let st = unsafe { ffi::gl_multi_sync(self.multi) };
        self.check(st)?;
        Ok(slabs)
    }
}

impl Drop for MultiContext {
    fn drop(&mut self) {
        unsafe { ffi::gl_multi_destroy(self.multi) }
    }
}

impl<'a> Drop for MultiMesh<'a> {
    fn drop(&mut self) {
        unsafe { ffi::gl_multi_mesh_free(self.multi.multi, self.mesh) }
    }
}

// ---- planned assembly ---------------------------------------------------------------------------------------------------------
/// A device buffer the caller owns (e.g. allocated by its sparse-solver crate): pointer + length, nothing more
#[derive(Clone, Copy)]
pub struct DevicePtr<T> {
    pub ptr: *mut T,
    pub len: usize,
}

/// `gl_asm_plan`: the slot of every element-block entry in the global CSR matrix (or vector), found once.  Every `assemble` is
/// then ONE kernel per GeoKind for the stiffness matrix -- the element blocks go from shared memory straight into `vals`.
pub struct AssemblyPlan<'a> {
    ctx: &'a Context,
    plan: *mut ffi::gl_asm_plan,
}

impl<'a> AssemblyPlan<'a> {
    /// `eq[point * dofs_per_node + i]` is the host point -> equation table (negative = prescribed dof, not assembled); the
    /// local-dof rule is the reference's (`ii = i + m*ndim`, src/integ/mat_10_bdb.rs:31-46).  `rowptr` (i64, nrow + 1) and
    /// `colidx` (i32, ascending within a row) describe the CSR pattern on the device; pass `None` for vector cases.
    pub fn new(
        ctx: &'a Context,
        mesh: &DeviceMesh,
        op: Op,
        cells: Range<usize>,
        args: &BatchArgs,
        eq: &[i32],
        pattern: Option<(DevicePtr<i64>, DevicePtr<i32>)>,
    ) -> Result<Self, StrError> {
        let a = args.to_c();
        let (rowptr, colidx) = match pattern {
            Some((r, c)) => (r.ptr as *const i64, c.ptr as *const i32),
            None => (ptr::null(), ptr::null()),
        };
        let mut missing = 0u64;
        let mut plan = ptr::null_mut();
        let st = unsafe {
            ffi::gl_assemble_plan_create(ctx.ctx, mesh.mesh, op as i32, cells.start as u64, cells.end as u64, &a, eq.as_ptr(), ffi::GL_HOST, rowptr, colidx, &mut missing, &mut plan)
        };
        ctx.check(st)?;
        Ok(AssemblyPlan { ctx, plan })
    }
    /// Integrates and ADDS into `vals` (CSR values or the global vector, on the device).  Asynchronous: `Context::sync` waits and
    /// reports zero determinants.
    pub fn assemble(&self, coef: &Coef, vals: DevicePtr<f64>) -> Result<(), StrError> {
        let c = coef.to_c();
        let st = unsafe { ffi::gl_assemble_planned(self.ctx.ctx, self.plan, &c, vals.ptr) };
        self.ctx.check(st)
    }
}

impl<'a> Drop for AssemblyPlan<'a> {
    fn drop(&mut self) {
        unsafe { ffi::gl_assemble_plan_free(self.ctx.ctx, self.plan) }
    }
}

pub mod integ {
    /// Batched counterparts of `gemlab::integ::*`: same names, a `Mesh` (uploaded) plus a cell range instead of one
    /// `Scratchpad`, coefficient data instead of a closure, `Vec<f64>` of packed column-major element blocks as result.
    pub mod batch {
        use crate::ffi;
        use crate::{static_error, BatchArgs, Coef, Context, DeviceMesh, DevicePtr, Op};
        use gemlab::StrError;
        use std::ops::Range;
        use std::ptr;

        pub(crate) fn run(op: i32, ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, coef: Option<&Coef>) -> Result<Vec<f64>, StrError> {
            let a = args.to_c();
            let (b, e) = (cells.start as u64, cells.end as u64);
            let mut total = 0u64;
            let st = unsafe { ffi::gl_out_total(mesh.mesh, op, b, e, &a, &mut total) };
            if st != 0 {
                return Err(static_error(st));
            }
            let mut out = vec![0.0f64; total as usize];
            let mut o = ffi::gl_out { ptr: out.as_mut_ptr(), location: ffi::GL_HOST, reserved: 0, capacity: total };
            let c = coef.map(|c| c.to_c());
            let cptr = c.as_ref().map_or(ptr::null(), |c| c as *const ffi::gl_coef);
            let st = unsafe { ffi::gl_integ(ctx.ctx, mesh.mesh, op, b, e, &a, cptr, &mut o) };
            ctx.check(st)?;
            Ok(out)
        }

        /// Start of every element block of the range in the batched output (`gl_out_offsets`; `off[n]` = total)
        pub fn offsets(mesh: &DeviceMesh, op: Op, cells: Range<usize>, args: &BatchArgs) -> Result<Vec<u64>, StrError> {
            let a = args.to_c();
            let mut off = vec![0u64; cells.len() + 1];
            let st = unsafe { ffi::gl_out_offsets(mesh.mesh, op as i32, cells.start as u64, cells.end as u64, &a, off.as_mut_ptr()) };
            if st != 0 {
                return Err(static_error(st));
            }
            Ok(off)
        }

        /// K = ∫ B·D·B (stiffness); replaces `integ::mat_10_bdb` (src/integ/mat_10_bdb.rs:121)
        pub fn mat_10_bdb(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, dd: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_10_BDB, ctx, mesh, cells, args, Some(dd))
        }
        /// K = ∫ N s N; replaces `integ::mat_01_nsn` (src/integ/mat_01_nsn.rs:48)
        pub fn mat_01_nsn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_01_NSN, ctx, mesh, cells, args, Some(s))
        }
        /// K = ∫ B·T·B; replaces `integ::mat_03_btb` (src/integ/mat_03_btb.rs:50)
        pub fn mat_03_btb(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, tt: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_03_BTB, ctx, mesh, cells, args, Some(tt))
        }
        /// K = ∫ (B·v) N; replaces `integ::mat_02_bvn` (src/integ/mat_02_bvn.rs:48)
        pub fn mat_02_bvn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, v: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_02_BVN, ctx, mesh, cells, args, Some(v))
        }
        /// K = ∫ N T N; replaces `integ::mat_08_ntn` (src/integ/mat_08_ntn.rs:56)
        pub fn mat_08_ntn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, tt: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_08_NTN, ctx, mesh, cells, args, Some(tt))
        }
        /// K = ∫ N v B; replaces `integ::mat_09_nvb` (src/integ/mat_09_nvb.rs:54)
        pub fn mat_09_nvb(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, v: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_09_NVB, ctx, mesh, cells, args, Some(v))
        }
        /// Two-pad coupling matrices (`args.kind_b` names the second pad): `integ::mat_04_nsb` (src/integ/mat_04_nsb.rs:66),
        /// `mat_05_btn` (:68), `mat_06_nvn` (:69), `mat_07_bsn` (:69)
        pub fn mat_04_nsb(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_04_NSB, ctx, mesh, cells, args, Some(s))
        }
        pub fn mat_05_btn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, tt: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_05_BTN, ctx, mesh, cells, args, Some(tt))
        }
        pub fn mat_06_nvn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, v: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_06_NVN, ctx, mesh, cells, args, Some(v))
        }
        pub fn mat_07_bsn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_07_BSN, ctx, mesh, cells, args, Some(s))
        }
        /// Boundary integrals over cells with geo_ndim = space_ndim - 1: `integ::mat_01_nsn_bry` (src/integ/mat_01_nsn_bry.rs:48),
        /// `vec_01_ns_bry` (:54), `vec_02_nv_bry` (:61); records are (s0, w) with s = s0 + w·un, resp. (v, qn) with v + qn un
        pub fn mat_01_nsn_bry(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_MAT_01_NSN_BRY, ctx, mesh, cells, args, Some(s))
        }
        pub fn vec_01_ns_bry(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_VEC_01_NS_BRY, ctx, mesh, cells, args, Some(s))
        }
        pub fn vec_02_nv_bry(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, v: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_VEC_02_NV_BRY, ctx, mesh, cells, args, Some(v))
        }
        /// a = ∫ N s; replaces `integ::vec_01_ns` (src/integ/vec_01_ns.rs:83)
        pub fn vec_01_ns(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_VEC_01_NS, ctx, mesh, cells, args, Some(s))
        }
        /// b = ∫ N v; replaces `integ::vec_02_nv` (src/integ/vec_02_nv.rs:85)
        pub fn vec_02_nv(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, v: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_VEC_02_NV, ctx, mesh, cells, args, Some(v))
        }
        /// c = ∫ B·w; replaces `integ::vec_03_bv` (src/integ/vec_03_bv.rs:84)
        pub fn vec_03_bv(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, w: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_VEC_03_BV, ctx, mesh, cells, args, Some(w))
        }
        /// d = ∫ B·σ; replaces `integ::vec_04_bt` (src/integ/vec_04_bt.rs:93)
        pub fn vec_04_bt(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, sig: &Coef) -> Result<Vec<f64>, StrError> {
            run(ffi::GL_OP_VEC_04_BT, ctx, mesh, cells, args, Some(sig))
        }
        /// det J at `ksi` for every cell; replaces the loop of `Mesh::check_jacobian` (src/mesh/check.rs:43-60)
        pub fn jacobian(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, ksi: &[f64]) -> Result<Vec<f64>, StrError> {
            let mut args = BatchArgs::new().to_c();
            for (i, v) in ksi.iter().take(3).enumerate() {
                args.ksi[i] = *v;
            }
            let (b, e) = (cells.start as u64, cells.end as u64);
            let mut out = vec![0.0f64; cells.len()];
            let mut o = ffi::gl_out { ptr: out.as_mut_ptr(), location: ffi::GL_HOST, reserved: 0, capacity: out.len() as u64 };
            let st = unsafe { ffi::gl_jacobian(ctx.ctx, mesh.mesh, b, e, &args, &mut o) };
            ctx.check(st)?;
            Ok(out)
        }
        /// Unit normals and magnitudes at the Gauss points of boundary cells: `ngauss x (un[space_ndim], mag_n)` per cell
        /// (`Scratchpad::calc_normal_vector`, src/shapes/scratchpad_calc_normal_vector.rs:111)
        pub fn normal_vectors(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs) -> Result<Vec<f64>, StrError> {
            let a = args.to_c();
            let (b, e) = (cells.start as u64, cells.end as u64);
            let mut total = 0u64;
            let st = unsafe { ffi::gl_out_total(mesh.mesh, ffi::GL_OP_NORMAL, b, e, &a, &mut total) };
            if st != 0 {
                return Err(static_error(st));
            }
            let mut out = vec![0.0f64; total as usize];
            let mut o = ffi::gl_out { ptr: out.as_mut_ptr(), location: ffi::GL_HOST, reserved: 0, capacity: total };
            let st = unsafe { ffi::gl_normal_vectors(ctx.ctx, mesh.mesh, b, e, &a, &mut o) };
            ctx.check(st)?;
            Ok(out)
        }

        /// Several cases over ONE geometry pass (`gl_integ_fused`): e.g. the mass matrix and the body-force vector, or the
        /// stiffness matrix and the internal-force vector of a Hex8 mesh (`Op::Mat10Bdb` with a uniform / per-cell modulus plus
        /// `Op::Vec04Bt` with `Coef::PerGauss(sigma)`).  Returns one host vector per item, in item order.
        pub fn fused(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, items: &[(Op, &Coef)]) -> Result<Vec<Vec<f64>>, StrError> {
            let a = args.to_c();
            let (b, e) = (cells.start as u64, cells.end as u64);
            let mut outs: Vec<Vec<f64>> = Vec::with_capacity(items.len());
            for (op, _) in items {
                let mut total = 0u64;
                let st = unsafe { ffi::gl_out_total(mesh.mesh, *op as i32, b, e, &a, &mut total) };
                if st != 0 {
                    return Err(static_error(st));
                }
                outs.push(vec![0.0f64; total as usize]);
            }
            let coefs: Vec<ffi::gl_coef> = items.iter().map(|(_, c)| c.to_c()).collect();
            let mut gouts: Vec<ffi::gl_out> =
                outs.iter_mut().map(|o| ffi::gl_out { ptr: o.as_mut_ptr(), location: ffi::GL_HOST, reserved: 0, capacity: o.len() as u64 }).collect();
            let its: Vec<ffi::gl_fused_item> = items
                .iter()
                .enumerate()
                .map(|(i, (op, _))| ffi::gl_fused_item { op: *op as i32, reserved: 0, coef: &coefs[i] as *const ffi::gl_coef, out: &mut gouts[i] as *mut ffi::gl_out })
                .collect();
            let st = unsafe { ffi::gl_integ_fused(ctx.ctx, mesh.mesh, b, e, &a, its.as_ptr(), its.len() as i32) };
            ctx.check(st)?;
            Ok(outs)
        }

        /// (rows, cols) of every double of the batched output of `op` (device arrays the caller owns, i32, one entry per double):
        /// together with a `DeviceSlab` this is the COO triplet list a `CooMatrix::put` loop would build (`gl_triplet_indices`).
        /// `cols` may be `None` for vector cases.
        pub fn triplet_indices(
            ctx: &Context,
            mesh: &DeviceMesh,
            op: Op,
            cells: Range<usize>,
            args: &BatchArgs,
            eq: &[i32],
            rows: DevicePtr<i32>,
            cols: Option<DevicePtr<i32>>,
        ) -> Result<(), StrError> {
            let a = args.to_c();
            let st = unsafe {
                ffi::gl_triplet_indices(
                    ctx.ctx,
                    mesh.mesh,
                    op as i32,
                    cells.start as u64,
                    cells.end as u64,
                    &a,
                    eq.as_ptr(),
                    ffi::GL_HOST,
                    rows.ptr,
                    cols.map_or(ptr::null_mut(), |c| c.ptr),
                )
            };
            ctx.check(st)
        }

        /// Integrates `op` and ADDS the element blocks into the CSR values `vals` of the pattern (`rowptr` i64, `colidx` i32:
        /// device buffers) -- `gl_assemble`.  For repeated assemblies build an `AssemblyPlan` instead (5x faster: no search, no
        /// element-matrix slab).
        pub fn assemble(
            ctx: &Context,
            mesh: &DeviceMesh,
            op: Op,
            cells: Range<usize>,
            args: &BatchArgs,
            coef: &Coef,
            eq: &[i32],
            rowptr: DevicePtr<i64>,
            colidx: DevicePtr<i32>,
            vals: DevicePtr<f64>,
        ) -> Result<(), StrError> {
            let a = args.to_c();
            let c = coef.to_c();
            let mut missing = 0u64;
            let st = unsafe {
                ffi::gl_assemble(ctx.ctx, mesh.mesh, op as i32, cells.start as u64, cells.end as u64, &a, &c, eq.as_ptr(), ffi::GL_HOST, rowptr.ptr, colidx.ptr, vals.ptr, &mut missing)
            };
            ctx.check(st)
        }

        /// The vector cases of `assemble`: adds the element vectors into the global vector `rhs` on the device
        pub fn assemble_vec(ctx: &Context, mesh: &DeviceMesh, op: Op, cells: Range<usize>, args: &BatchArgs, coef: &Coef, eq: &[i32], rhs: DevicePtr<f64>) -> Result<(), StrError> {
            let a = args.to_c();
            let c = coef.to_c();
            let st = unsafe {
                ffi::gl_assemble(ctx.ctx, mesh.mesh, op as i32, cells.start as u64, cells.end as u64, &a, &c, eq.as_ptr(), ffi::GL_HOST, ptr::null(), ptr::null(), rhs.ptr, ptr::null_mut())
            };
            ctx.check(st)
        }

        /// Outputs that STAY ON THE DEVICE: the 1.4 G elements/s path (no PCIe); asynchronous, `Context::sync` waits
        pub mod device {
            use crate::ffi;
            use crate::{BatchArgs, Coef, Context, DeviceMesh, DeviceSlab, Op};
            use gemlab::StrError;
            use std::ops::Range;

            pub fn integ(ctx: &Context, mesh: &DeviceMesh, op: Op, cells: Range<usize>, args: &BatchArgs, coef: &Coef) -> Result<DeviceSlab, StrError> {
                let a = args.to_c();
                let c = coef.to_c();
                let mut slab = DeviceSlab::empty();
                let st = unsafe { ffi::gl_integ_slab(ctx.ctx, mesh.mesh, op as i32, cells.start as u64, cells.end as u64, &a, &c, &mut slab.slab) };
                ctx.check(st)?;
                Ok(slab)
            }
            pub fn mat_10_bdb(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, dd: &Coef) -> Result<DeviceSlab, StrError> {
                integ(ctx, mesh, Op::Mat10Bdb, cells, args, dd)
            }
            pub fn mat_01_nsn(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<DeviceSlab, StrError> {
                integ(ctx, mesh, Op::Mat01Nsn, cells, args, s)
            }
            pub fn mat_03_btb(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, tt: &Coef) -> Result<DeviceSlab, StrError> {
                integ(ctx, mesh, Op::Mat03Btb, cells, args, tt)
            }
            pub fn vec_04_bt(ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, sig: &Coef) -> Result<DeviceSlab, StrError> {
                integ(ctx, mesh, Op::Vec04Bt, cells, args, sig)
            }
        }

        /// The one-call form: `&Mesh` + cell range (uploads the mesh, integrates, frees the device mesh).  For more than one call
        /// on the same mesh keep a `DeviceMesh` (`Context::upload`) and use the functions above.
        pub mod for_mesh {
            use crate::{BatchArgs, Coef, Context, Op};
            use gemlab::mesh::Mesh;
            use gemlab::StrError;
            use std::ops::Range;

            pub fn integ(ctx: &Context, mesh: &Mesh, op: Op, cells: Range<usize>, args: &BatchArgs, coef: &Coef) -> Result<Vec<f64>, StrError> {
                let dmesh = ctx.upload(mesh)?;
                super::run(op as i32, ctx, &dmesh, cells, args, Some(coef))
            }
            pub fn mat_10_bdb(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, dd: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Mat10Bdb, cells, args, dd)
            }
            pub fn mat_01_nsn(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Mat01Nsn, cells, args, s)
            }
            pub fn mat_03_btb(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, tt: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Mat03Btb, cells, args, tt)
            }
            pub fn vec_01_ns(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, s: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Vec01Ns, cells, args, s)
            }
            pub fn vec_02_nv(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, v: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Vec02Nv, cells, args, v)
            }
            pub fn vec_03_bv(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, w: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Vec03Bv, cells, args, w)
            }
            pub fn vec_04_bt(ctx: &Context, mesh: &Mesh, cells: Range<usize>, args: &BatchArgs, sig: &Coef) -> Result<Vec<f64>, StrError> {
                integ(ctx, mesh, Op::Vec04Bt, cells, args, sig)
            }
        }

        /// legacy (gemlab <= 1.x) names
        pub use mat_10_bdb as mat_10_gdg;
        pub use vec_03_bv as vec_10_gs;
    }
}

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
//! let dmesh = ctx.upload(&mesh)?;                  // SoA device mesh (replaces Mesh::set_pad per cell)
//! let model = LinElasticity::new(480.0, 1.0 / 3.0, false, false);
//! let dd = Coef::uniform_tensor4(model.get_modulus());
//! // element e occupies kk[e*576 .. (e+1)*576], column-major 24x24, row ii = i + m*3, col jj = j + n*3
//! let kk = batch::mat_10_bdb(&ctx, &dmesh, 0..mesh.cells.len(), &BatchArgs::new(), &dd)?;
//! ```
//!
//! Multi-GPU: create one `Context` per device (each on its own thread), give rank `g` of `G` the contiguous range
//! `partition(ncell, G, g)` and concatenate the outputs in rank order; there is no collective on the data path.
//!
//! NOTE: shipped as source; not compiled in the authoring image (no cargo). The ABI is defined by
//! `include/gemlab_cuda.h`; see INTEGRATION.md.

pub mod ffi;

use gemlab::mesh::Mesh;
use gemlab::StrError;
use russell_tensor::Tensor4;
use std::ffi::CStr;
use std::ops::Range;
use std::ptr;

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
    pub kind_b: Option<gemlab::shapes::GeoKind>,
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
    PerGauss(Vec<f64>),
}

impl Coef {
    /// `dd.set_tensor(1.0, model.get_modulus())` for every Gauss point: the Mandel matrix, column-major as stored by
    /// russell_lab::Matrix
    pub fn uniform_tensor4(dd: &Tensor4) -> Self {
        Coef::Uniform(dd.matrix().as_data().clone())
    }
}

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

fn static_error(status: i32) -> StrError {
    // the library returns the reference's own static strings for the reference's error conditions
    let s = unsafe { CStr::from_ptr(ffi::gl_status_string(status)) };
    match s.to_str() {
        Ok(v) => Box::leak(v.to_string().into_boxed_str()),
        Err(_) => "libgemlab_cuda error",
    }
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

    /// Flattens `Mesh.points` / `Mesh.cells` and uploads them (the library transposes to structure-of-arrays and
    /// buckets cells by GeoKind). Replaces the per-cell `Mesh::set_pad` (src/mesh/mesh.rs:448-456).
    pub fn upload<'a>(&'a self, mesh: &Mesh) -> Result<DeviceMesh<'a>, StrError> {
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
        let mut handle = ptr::null_mut();
        let st = unsafe {
            ffi::gl_mesh_upload(
                self.ctx,
                ndim as i32,
                mesh.points.len() as u64,
                coords.as_ptr(),
                mesh.cells.len() as u64,
                kinds.as_ptr(),
                markers.as_ptr(),
                conn_off.as_ptr(),
                conn.as_ptr(),
                &mut handle,
            )
        };
        if st != 0 {
            return Err(static_error(st));
        }
        Ok(DeviceMesh { ctx: self, mesh: handle, ncell: mesh.cells.len() })
    }

    pub fn sync(&self) -> Result<(), StrError> {
        let st = unsafe { ffi::gl_ctx_sync(self.ctx) };
        if st != 0 {
            return Err(static_error(st));
        }
        Ok(())
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

/// Contiguous cell range of rank `g` out of `world`: [floor(g*ncell/G), floor((g+1)*ncell/G))
pub fn partition(ncell: usize, world: usize, rank: usize) -> Range<usize> {
    (rank * ncell) / world..((rank + 1) * ncell) / world
}

pub mod integ {
    /// Batched counterparts of `gemlab::integ::*`: same names, a `Mesh` (uploaded) plus a cell range instead of one
    /// `Scratchpad`, coefficient data instead of a closure, `Vec<f64>` of packed column-major element blocks as result.
    pub mod batch {
        use crate::ffi;
        use crate::{static_error, BatchArgs, Coef, Context, DeviceMesh};
        use gemlab::StrError;
        use std::ops::Range;
        use std::ptr;

        fn run(op: i32, ctx: &Context, mesh: &DeviceMesh, cells: Range<usize>, args: &BatchArgs, coef: Option<&Coef>) -> Result<Vec<f64>, StrError> {
            let a = args.to_c();
            let (b, e) = (cells.start as u64, cells.end as u64);
            let mut total = 0u64;
            let st = unsafe { ffi::gl_out_total(mesh.mesh, op, b, e, &a, &mut total) };
            if st != 0 {
                return Err(static_error(st));
            }
            let mut out = vec![0.0f64; total as usize];
            let mut o = ffi::gl_out { ptr: out.as_mut_ptr(), location: ffi::GL_HOST, reserved: 0, capacity: total };
            let c = coef.map(|c| match c {
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
            });
            let cptr = c.as_ref().map_or(ptr::null(), |c| c as *const ffi::gl_coef);
            let st = unsafe { ffi::gl_integ(ctx.ctx, mesh.mesh, op, b, e, &a, cptr, &mut o) };
            if st != 0 {
                return Err(static_error(st));
            }
            Ok(out)
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
            if st != 0 {
                return Err(static_error(st));
            }
            Ok(out)
        }
        /// Assembly hand-off: integrates `op` over the cells and ADDS the element blocks into a global object that lives on
        /// the device -- the CSR values `vals` of the pattern (`rowptr`, `colidx`: device pointers) for matrix cases, the
        /// global vector `vals` (null `rowptr` / `colidx`) for vector cases.  `eq[point * dofs_per_node + i]` is the host
        /// point -> equation table (negative = prescribed dof); the local-dof rule is the reference's
        /// (src/integ/mat_10_bdb.rs:31-46).  Nothing but the matrix values ever needs to leave the GPU.
        ///
        /// # Safety
        /// `rowptr`, `colidx` and `vals` must be valid device pointers of the sizes the pattern implies.
        pub unsafe fn assemble(
            op: i32,
            ctx: &Context,
            mesh: &DeviceMesh,
            cells: Range<usize>,
            args: &BatchArgs,
            coef: &[f64],
            eq: &[i32],
            rowptr: *const i64,
            colidx: *const i32,
            vals: *mut f64,
        ) -> Result<(), StrError> {
            let a = args.to_c();
            let c = ffi::gl_coef { mode: ffi::GL_COEF_UNIFORM, location: ffi::GL_HOST, data: coef.as_ptr(), markers: ptr::null(), nrecord: 1 };
            let mut missing = 0u64;
            let st = ffi::gl_assemble(ctx.ctx, mesh.mesh, op, cells.start as u64, cells.end as u64, &a, &c, eq.as_ptr(), ffi::GL_HOST, rowptr, colidx, vals, &mut missing);
            if st != 0 {
                return Err(static_error(st));
            }
            Ok(())
        }
        /// legacy (gemlab <= 1.x) names
        pub use mat_10_bdb as mat_10_gdg;
        pub use vec_03_bv as vec_10_gs;
    }
}

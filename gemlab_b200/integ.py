"""Batched `integ` entry points -- host-side mirror of gemlab::integ for a Mesh plus a cell range.

The reference integrates ONE cell per call: `integ::mat_10_bdb(&mut kk, &mut CommonArgs{pad, gauss, ..}, |dd,p,N,B| ..)`
(src/integ/mat_10_bdb.rs:121; siblings mat_01_nsn.rs:48, mat_03_btb.rs:50, vec_01_ns.rs:83, vec_02_nv.rs:85,
vec_03_bv.rs:84, vec_04_bt.rs:93) inside a user loop `for cell in &mesh.cells { mesh.set_pad(..); integ::X(..) }`.
Here the same functions take the whole mesh and a cell range and run on the GPU through the C ABI
(include/gemlab_cuda.h); names, argument meaning and error strings follow the reference:

    ctx   = Context(device=0)                       # one per GPU
    dmesh = ctx.upload(mesh)                        # SoA device mesh (replaces Mesh::set_pad's gather)
    args  = CommonArgs()                            # clear=True, axisymmetric=False, alpha=1, ii0=jj0=0, ngauss=None
    kk    = integ.mat_10_bdb(ctx, dmesh, (0, ncell), args, Coef.uniform(dd))

Outputs: element e of the range occupies out[off[e-begin] : off[e-begin]+size], column-major inside the element.
`out` may be None (a new numpy array is returned), a numpy array (host path, copies inside the call) or a CUDA
torch tensor (device path, asynchronous on the ctx stream).  PyTorch is plumbing only (device memory, streams).
"""
import ctypes as C

import numpy as np

from ._lib import GlArgs, GlCoef, GlFusedItem, GlOut, lib
from .geo import GemlabError, GeoKind
from .mesh import Mesh

GL_HOST, GL_DEVICE = 0, 1
OPS = {"mat_02_bvn": 9, "mat_08_ntn": 10, "mat_09_nvb": 11, "mat_10_bdb": 0, "mat_01_nsn": 1, "mat_03_btb": 2, "vec_01_ns": 3, "vec_03_bv": 4, "vec_02_nv": 5,
       "vec_04_bt": 6, "jacobian": 8, "mat_04_nsb": 12, "mat_05_btn": 13, "mat_06_nvn": 14, "mat_07_bsn": 15,
       "mat_01_nsn_bry": 16, "vec_01_ns_bry": 17, "vec_02_nv_bry": 18, "normal": 19}
COEF_UNIFORM, COEF_PER_MARKER, COEF_PER_CELL, COEF_PER_GAUSS = 0, 1, 2, 3


class CommonArgs:
    """gemlab::integ::CommonArgs (src/integ/common_args.rs:5-43) without pad/gauss (implied by the mesh and `ngauss`)."""

    def __init__(self, ngauss=None, clear=True, axisymmetric=False, alpha=1.0, ii0=0, jj0=0, nrow=0, ncol=0, packed=False, kind_b=None):
        self.ngauss = ngauss          # None = Gauss::new(kind), else Gauss::new_sized(kind.class(), ngauss)
        self.clear = clear
        self.axisymmetric = axisymmetric
        self.alpha = alpha
        self.ii0 = ii0
        self.jj0 = jj0
        self.nrow = nrow              # 0 = tight element blocks
        self.ncol = ncol
        self.ksi = (0.0, 0.0, 0.0)    # integ.jacobian only
        self.packed = packed          # batched boundary only: matrices leave as packed upper triangles (gl_args.packed)
        self.kind_b = kind_b          # two-pad cases (mat_04_nsb .. mat_07_bsn): GeoKind of the second Scratchpad (pad_b)

    def _c(self):
        a = GlArgs()
        a.ngauss = 0 if self.ngauss is None else int(self.ngauss)
        a.axisymmetric = int(bool(self.axisymmetric))
        a.clear = int(bool(self.clear))
        a.packed = int(bool(self.packed))
        a.kind_b = -1 if self.kind_b is None else int(self.kind_b)
        a.alpha = float(self.alpha)
        a.ii0, a.jj0, a.nrow, a.ncol = int(self.ii0), int(self.jj0), int(self.nrow), int(self.ncol)
        for i in range(3):
            a.ksi[i] = float(self.ksi[i]) if i < len(self.ksi) else 0.0
        return a


class Coef:
    """Coefficient data replacing the reference's per-Gauss-point closures (see struct gl_coef)."""

    def __init__(self, mode, data, markers=None):
        self.mode = mode
        self.data = data
        self.markers = None if markers is None else np.ascontiguousarray(markers, dtype=np.int32)

    @staticmethod
    def uniform(record):
        """One record for every cell and Gauss point (e.g. `dd.set_tensor(1.0, model.get_modulus())`)."""
        return Coef(COEF_UNIFORM, np.asarray(record, dtype=np.float64))

    @staticmethod
    def per_marker(markers, records):
        return Coef(COEF_PER_MARKER, np.asarray(records, dtype=np.float64), markers)

    @staticmethod
    def per_cell(records):
        """records[(e - cell_begin)] ; numpy array (host) or CUDA torch tensor (device)."""
        return Coef(COEF_PER_CELL, records)

    @staticmethod
    def per_gauss(records):
        """records[(e - cell_begin) * ngauss + p]."""
        return Coef(COEF_PER_GAUSS, records)


def mandel_matrix(dd):
    """Flattens an (ns, ns) Mandel matrix (Tensor4::matrix()) COLUMN-major as the ABI expects."""
    return np.asarray(dd, dtype=np.float64).ravel(order="F").copy()


def lin_elasticity(young, poisson, two_dim, plane_stress=False):
    """LinElasticity::new(E, nu, two_dim, plane_stress).get_modulus().matrix() -> (ns, ns)."""
    ns = 4 if two_dim else 6
    dd = np.zeros(ns * ns)
    lib().gl_lin_elasticity(float(young), float(poisson), int(two_dim), int(plane_stress), dd.ctypes.data)
    return dd.reshape(ns, ns, order="F")


def _is_torch_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


class DeviceMesh:
    """Handle of a mesh resident on one GPU (struct gl_mesh)."""

    def __init__(self, ctx, handle, ndim, ncell, npoint, kinds):
        self.ctx = ctx
        self._h = handle
        self.ndim = ndim
        self.ncell = ncell
        self.npoint = npoint
        self.kinds = kinds  # host copy (uint8) -- used for sizes only

    def update_coords(self, coords):
        """Replaces the point coordinates (host array npoint x ndim, as Mesh.coords) of this device mesh (gl_mesh_update_coords)."""
        arr = np.ascontiguousarray(coords, dtype=np.float64)
        if arr.shape != (self.npoint, self.ndim):
            raise ValueError("coords must be npoint x ndim")
        st = lib().gl_mesh_update_coords(self.ctx._h, self._h, arr.ctypes.data, GL_HOST)
        if st:
            self.ctx._raise(st)

    def free(self):
        if self._h:
            lib().gl_mesh_free(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """One per GPU (struct gl_ctx).  `stream` is a raw cudaStream_t handle; by default torch's current stream of that
    device is used so that torch.cuda.Event timings bracket the kernels."""

    def __init__(self, device=0, stream=None):
        self.device = int(device)
        if stream is None:
            try:
                import torch

                stream = torch.cuda.current_stream(self.device).cuda_stream
            except Exception:
                stream = 0
        h = C.c_void_p()
        st = lib().gl_ctx_create(self.device, C.c_void_p(stream), C.byref(h))
        if st:
            raise GemlabError(st)
        self._h = h

    def close(self):
        if self._h:
            lib().gl_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, status):
        msg = lib().gl_last_error(self._h).decode("utf-8")
        cell = lib().gl_last_error_cell(self._h)
        raise GemlabError(status, msg or None, None if cell == 0xFFFFFFFFFFFFFFFF else cell)

    def set_stream(self, stream):
        lib().gl_ctx_set_stream(self._h, C.c_void_p(stream))

    def sync(self):
        st = lib().gl_ctx_sync(self._h)
        if st:
            self._raise(st)

    def launch_count(self):
        return int(lib().gl_ctx_launch_count(self._h))

    def last_kernel(self):
        """Kernel family (with its template parameters) of the most recent integration launch (gl_ctx_last_kernel)."""
        return lib().gl_ctx_last_kernel(self._h).decode("ascii")

    def upload(self, mesh):
        """Mesh -> device SoA (gl_mesh_upload); replaces the per-cell Mesh::set_pad gather (src/mesh/mesh.rs:448-456)."""
        if not isinstance(mesh, Mesh):
            raise TypeError("expected a gemlab_b200.Mesh")
        h = C.c_void_p()
        # the library detects single-kind meshes itself and transposes those on the device (no host-side copies here)
        st = lib().gl_mesh_upload(self._h, mesh.ndim, mesh.npoint, mesh.coords.ctypes.data, mesh.ncell,
                                  mesh.kinds.ctypes.data, mesh.markers.ctypes.data, mesh.conn_off.ctypes.data,
                                  mesh.conn.ctypes.data, C.byref(h))
        if st:
            self._raise(st)
        return DeviceMesh(self, h, mesh.ndim, mesh.ncell, mesh.npoint, mesh.kinds.copy())

    def mesh_from_text(self, text, fmt="msh"):
        """gl_mesh_from_text: parse gemlab's MSH text (fmt="msh") or JSON (fmt="json") mesh format with the library's own parser and
        upload it (Mesh::read / Mesh::read_json followed by the per-cell Mesh::set_pad gather, in one call)."""
        data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        st = lib().gl_mesh_from_text(self._h, data, len(data), {"msh": 0, "json": 1}[fmt], C.byref(h))
        if st:
            self._raise(st)
        ncell, npoint, ndim = int(lib().gl_mesh_ncell(h)), int(lib().gl_mesh_npoint(h)), int(lib().gl_mesh_ndim(h))
        kinds = Mesh.from_text_native(text, fmt).kinds
        return DeviceMesh(self, h, ndim, ncell, npoint, kinds)

    def block(self, kind, ndiv, xmin=None, xmax=None, marker=1):
        """Device-side lattice generator (gl_mesh_block): the mesh of Block::new(box).set_ndiv(ndiv).subdivide(kind) for an
        axis-aligned box, built in HBM with the reference's cell order and first-touch point numbering."""
        kind = GeoKind(kind)
        ndim = kind.ndim()
        nd = (C.c_uint32 * 3)(*[int(v) for v in list(ndiv) + [1] * (3 - len(ndiv))])
        lo = (C.c_double * 3)(*([0.0] * 3 if xmin is None else [float(v) for v in list(xmin) + [0.0] * (3 - len(xmin))]))
        hi = (C.c_double * 3)(*([1.0] * 3 if xmax is None else [float(v) for v in list(xmax) + [0.0] * (3 - len(xmax))]))
        h = C.c_void_p()
        st = lib().gl_mesh_block(self._h, int(kind), nd, lo, hi, int(marker), C.byref(h))
        if st:
            self._raise(st)
        ncell = int(lib().gl_mesh_ncell(h))
        npoint = int(lib().gl_mesh_npoint(h))
        return DeviceMesh(self, h, ndim, ncell, npoint, np.full(ncell, int(kind), dtype=np.uint8))

    def block_general(self, kind, ndiv, control, marker=1):
        """gl_mesh_block_general: Block::new(control).set_ndiv(ndiv).subdivide(kind) on the device for any Block target and a
        straight-edged (corners) or curved (Qua8 / Hex20 nodes) block."""
        kind = GeoKind(kind)
        ndim = kind.ndim()
        ctrl = np.ascontiguousarray(control, dtype=np.float64)
        nd = (C.c_uint32 * 3)(*[int(v) for v in list(ndiv) + [1] * (3 - len(ndiv))])
        h = C.c_void_p()
        st = lib().gl_mesh_block_general(self._h, int(kind), nd, int(ctrl.shape[0]), ctrl.ctypes.data, int(marker), C.byref(h))
        if st:
            self._raise(st)
        ncell, npoint = int(lib().gl_mesh_ncell(h)), int(lib().gl_mesh_npoint(h))
        return DeviceMesh(self, h, ndim, ncell, npoint, np.full(ncell, int(kind), dtype=np.uint8))

    def download(self, dmesh):
        """Device mesh -> host Mesh (gl_mesh_download; single-kind meshes)."""
        kind = GeoKind(int(dmesh.kinds[0]))
        coords = np.empty((dmesh.npoint, dmesh.ndim))
        conn = np.empty((dmesh.ncell, kind.nnode()), dtype=np.uint32)
        st = lib().gl_mesh_download(self._h, dmesh._h, coords.ctypes.data, conn.ctypes.data)
        if st:
            self._raise(st)
        return Mesh.homogeneous(dmesh.ndim, coords, kind, conn.astype(np.uint64))

    def fp64_peak(self, kind="dfma", iters=4096):
        """Measured FP64 peak of this device in TFLOP/s (kind: 'dfma', 'dmma' or 'mixed' = both interleaved)."""
        t = C.c_double()
        st = lib().gl_bench_fp64_peak(self._h, {"dfma": 0, "dmma": 1, "mixed": 2}[kind], int(iters), C.byref(t))
        if st:
            self._raise(st)
        return t.value


def _hbm_write_peak(self, buf, kind="bulk", warps_per_sm=8, mode=0):
    """Measured write bandwidth (GB/s) into the device tensor `buf` (kind: 'bulk' = cp.async.bulk shared->global in
    4,608-byte blocks, 'stg' = 16-byte register stores, 'memset' = cudaMemsetAsync)."""
    t = C.c_double()
    st = lib().gl_bench_hbm_write(self._h, {"bulk": 0, "stg": 1, "memset": 2}[kind] | (int(mode) << 4), buf.data_ptr(),
                                  buf.numel() * buf.element_size(), int(warps_per_sm), C.byref(t))
    if st:
        self._raise(st)
    return t.value


Context.hbm_write_peak = _hbm_write_peak


def out_total(dmesh, op, cell_range, args):
    b, e = _range(dmesh, cell_range)
    a = args._c()
    total = C.c_uint64()
    st = lib().gl_out_total(dmesh._h, OPS[op], b, e, C.byref(a), C.byref(total))
    if st:
        raise GemlabError(st)
    return total.value


def out_offsets(dmesh, op, cell_range, args):
    """off[0..n]: start of every element block of the range (gl_out_offsets)."""
    b, e = _range(dmesh, cell_range)
    a = args._c()
    off = np.zeros(e - b + 1, dtype=np.uint64)
    st = lib().gl_out_offsets(dmesh._h, OPS[op], b, e, C.byref(a), off.ctypes.data)
    if st:
        raise GemlabError(st)
    return off


def _range(dmesh, cell_range):
    if cell_range is None:
        return 0, dmesh.ncell
    if isinstance(cell_range, range):
        if cell_range.step != 1:
            raise ValueError("cell range must be contiguous")
        return cell_range.start, cell_range.stop
    b, e = cell_range
    return int(b), int(e)


def _marshal_coef(coef, keep):
    c = GlCoef()
    c.mode = coef.mode
    if _is_torch_tensor(coef.data):
        if not coef.data.is_cuda:
            raise ValueError("coefficient tensor must be a CUDA tensor or a numpy array")
        t = coef.data.contiguous()
        keep.append(t)
        c.location, c.data = GL_DEVICE, t.data_ptr()
    else:
        arr = np.ascontiguousarray(coef.data, dtype=np.float64)
        keep.append(arr)
        c.location, c.data = GL_HOST, arr.ctypes.data
    if coef.markers is not None:
        c.markers = coef.markers.ctypes.data
        c.nrecord = coef.markers.shape[0]
    return c


def _marshal_out(out, total, clear):
    o = GlOut()
    result = out
    if out is None:
        result = np.empty(total, dtype=np.float64) if clear else np.zeros(total, dtype=np.float64)
        o.ptr, o.location, o.capacity = result.ctypes.data, GL_HOST, total
    elif _is_torch_tensor(out):
        if not out.is_contiguous():
            raise ValueError("output tensor must be contiguous")
        o.ptr, o.capacity = out.data_ptr(), out.numel()
        o.location = GL_DEVICE if out.is_cuda else GL_HOST
    else:
        if not (isinstance(out, np.ndarray) and out.dtype == np.float64 and out.flags.c_contiguous):
            raise ValueError("output must be a contiguous float64 numpy array")
        o.ptr, o.location, o.capacity = out.ctypes.data, GL_HOST, out.size
    return o, result


def _total(dmesh, op, b, e, a):
    total = C.c_uint64()
    st = lib().gl_out_total(dmesh._h, OPS[op], b, e, C.byref(a), C.byref(total))
    if st:
        raise GemlabError(st)
    return total.value


def _integ(op, ctx, dmesh, cell_range, args, coef, out):
    b, e = _range(dmesh, cell_range)
    a = args._c()
    total = _total(dmesh, op, b, e, a)
    keep = []
    c = None if coef is None else _marshal_coef(coef, keep)
    o, result = _marshal_out(out, total, args.clear)
    st = lib().gl_integ(ctx._h, dmesh._h, OPS[op], b, e, C.byref(a), None if c is None else C.byref(c), C.byref(o))
    if st:
        ctx._raise(st)
    return result


def fused(ctx, dmesh, cell_range, args, items):
    """Several integration cases from ONE geometry pass (gl_integ_fused).  `items` is a list of (op name, Coef, out);
    returns the list of outputs.  Equivalent to calling each `integ.<op>` separately; the library fuses mat_01_nsn,
    mat_03_btb, vec_01_ns and vec_03_bv (uniform coefficients, CUDA-tensor outputs) into one kernel."""
    b, e = _range(dmesh, cell_range)
    a = args._c()
    keep, results = [], []
    arr = (GlFusedItem * len(items))()
    for i, (op, coef, out) in enumerate(items):
        total = _total(dmesh, op, b, e, a)
        c = _marshal_coef(coef, keep)
        o, result = _marshal_out(out, total, args.clear)
        keep.extend([c, o])
        arr[i].op = OPS[op]
        arr[i].coef = C.pointer(c)
        arr[i].out = C.pointer(o)
        results.append(result)
    st = lib().gl_integ_fused(ctx._h, dmesh._h, b, e, C.byref(a), arr, len(items))
    if st:
        ctx._raise(st)
    return results


def mat_10_bdb(ctx, dmesh, cell_range, args, coef, out=None):
    """K = int B.D.B (stiffness), ndof x ndof per cell -- integ::mat_10_bdb (src/integ/mat_10_bdb.rs:121-233)."""
    return _integ("mat_10_bdb", ctx, dmesh, cell_range, args, coef, out)


mat_10_gdg = mat_10_bdb  # legacy (gemlab <= 1.x) name used by BASELINE.json


def mat_01_nsn(ctx, dmesh, cell_range, args, coef, out=None):
    """K = int N s N, nnode x nnode -- integ::mat_01_nsn (src/integ/mat_01_nsn.rs:48-102)."""
    return _integ("mat_01_nsn", ctx, dmesh, cell_range, args, coef, out)


def mat_03_btb(ctx, dmesh, cell_range, args, coef, out=None):
    """K = int B.T.B, nnode x nnode -- integ::mat_03_btb (src/integ/mat_03_btb.rs:50-131)."""
    return _integ("mat_03_btb", ctx, dmesh, cell_range, args, coef, out)


def mat_02_bvn(ctx, dmesh, cell_range, args, coef, out=None):
    """K = int (B.v) N, nnode x nnode -- integ::mat_02_bvn (src/integ/mat_02_bvn.rs:48-120); coefficient v (space_ndim)."""
    return _integ("mat_02_bvn", ctx, dmesh, cell_range, args, coef, out)


def mat_08_ntn(ctx, dmesh, cell_range, args, coef, out=None):
    """K = int N T N, ndof x ndof -- integ::mat_08_ntn (src/integ/mat_08_ntn.rs:56-135); coefficient Mandel vector of T."""
    return _integ("mat_08_ntn", ctx, dmesh, cell_range, args, coef, out)


def mat_09_nvb(ctx, dmesh, cell_range, args, coef, out=None):
    """K = int N v B, ndof x ndof -- integ::mat_09_nvb (src/integ/mat_09_nvb.rs:54-125); coefficient v (space_ndim)."""
    return _integ("mat_09_nvb", ctx, dmesh, cell_range, args, coef, out)


def vec_01_ns(ctx, dmesh, cell_range, args, coef, out=None):
    """a = int N s, nnode -- integ::vec_01_ns (src/integ/vec_01_ns.rs:83-130)."""
    return _integ("vec_01_ns", ctx, dmesh, cell_range, args, coef, out)


def vec_02_nv(ctx, dmesh, cell_range, args, coef, out=None):
    """b = int N v, ndof -- integ::vec_02_nv (src/integ/vec_02_nv.rs:85-140)."""
    return _integ("vec_02_nv", ctx, dmesh, cell_range, args, coef, out)


def vec_03_bv(ctx, dmesh, cell_range, args, coef, out=None):
    """c = int B.w, nnode -- integ::vec_03_bv (src/integ/vec_03_bv.rs:84-141)."""
    return _integ("vec_03_bv", ctx, dmesh, cell_range, args, coef, out)


vec_10_gs = vec_03_bv  # legacy name used by BASELINE.json


def vec_04_bt(ctx, dmesh, cell_range, args, coef, out=None):
    """d = int B.sigma, ndof -- integ::vec_04_bt (src/integ/vec_04_bt.rs:93-170)."""
    return _integ("vec_04_bt", ctx, dmesh, cell_range, args, coef, out)


def mat_04_nsb(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::mat_04_nsb (src/integ/mat_04_nsb.rs:66): K = int Nb s B, nnode_b x ndof; args.kind_b names the second pad."""
    return _integ("mat_04_nsb", ctx, dmesh, cell_range, args, coef, out)


def mat_05_btn(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::mat_05_btn (src/integ/mat_05_btn.rs:68): K = int Bb.T N, nnode_b x ndof; coefficient = Mandel vector of T."""
    return _integ("mat_05_btn", ctx, dmesh, cell_range, args, coef, out)


def mat_06_nvn(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::mat_06_nvn (src/integ/mat_06_nvn.rs:69): K = int N v Nb, ndof x nnode_b."""
    return _integ("mat_06_nvn", ctx, dmesh, cell_range, args, coef, out)


def mat_07_bsn(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::mat_07_bsn (src/integ/mat_07_bsn.rs:69): K = int B s Nb, ndof x nnode_b."""
    return _integ("mat_07_bsn", ctx, dmesh, cell_range, args, coef, out)


def mat_01_nsn_bry(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::mat_01_nsn_bry (src/integ/mat_01_nsn_bry.rs:48) over boundary cells; record (s0, w..): s = s0 + w.un."""
    return _integ("mat_01_nsn_bry", ctx, dmesh, cell_range, args, coef, out)


def vec_01_ns_bry(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::vec_01_ns_bry (src/integ/vec_01_ns_bry.rs:54) over boundary cells; record (s0, w..): s = s0 + w.un."""
    return _integ("vec_01_ns_bry", ctx, dmesh, cell_range, args, coef, out)


def vec_02_nv_bry(ctx, dmesh, cell_range, args, coef, out=None):
    """integ::vec_02_nv_bry (src/integ/vec_02_nv_bry.rs:61) over boundary cells; record (v.., qn): v = v + qn un."""
    return _integ("vec_02_nv_bry", ctx, dmesh, cell_range, args, coef, out)


def normal_vectors(ctx, dmesh, cell_range, args, out=None):
    """Scratchpad::calc_normal_vector at every Gauss point (scratchpad_calc_normal_vector.rs:111): per cell ng x (un.., mag_n)."""
    return _integ("normal", ctx, dmesh, cell_range, args, None, out)


def jacobian(ctx, dmesh, cell_range, ksi, out=None):
    """det J (SOLID), |J| (CABLE) or -1 (SHELL) of every cell at the reference point ksi --
    the Scratchpad::calc_jacobian loop of Mesh::check_jacobian (src/mesh/check.rs:43-60)."""
    args = CommonArgs()
    args.ksi = tuple(ksi)
    return _integ("jacobian", ctx, dmesh, cell_range, args, None, out)


def unpack_symmetric(packed, n):
    """Packed upper triangles (gl_args.packed: K(i,j), i <= j, at j(j+1)/2 + i) -> full symmetric blocks [ncell][n*n] (column-major)."""
    tri = n * (n + 1) // 2
    packed = np.asarray(packed).reshape(-1, tri)
    jj, ii = np.triu_indices(n)  # row-major upper of the transposed = column-major upper: pairs (j >= i) sorted by j then i
    order = np.lexsort((jj, ii))  # sort by column (ii here is the larger index) then row
    cols, rows = ii[order], jj[order]
    full = np.empty((packed.shape[0], n, n))
    full[:, cols, rows] = packed   # [cell][col][row]: column-major blocks
    full[:, rows, cols] = packed
    return full.reshape(packed.shape[0], n * n)


def partition(ncell, world_size, rank):
    """Contiguous cell range of `rank`: [floor(r*ncell/G), floor((r+1)*ncell/G)); the concatenation of
    the ranks' outputs in rank order equals the single-GPU output."""
    return (rank * ncell) // world_size, ((rank + 1) * ncell) // world_size


def op_out_size(op, kind, ndim):
    return int(lib().gl_op_out_size(OPS[op], int(GeoKind(kind)), int(ndim)))


def op_coef_size(op, ndim):
    return int(lib().gl_op_coef_size(OPS[op], int(ndim)))


# ---- assembly hand-off (include/gemlab_cuda.h: gl_triplet_indices, gl_assemble) ---------------------------------------

def dofs_per_node(op, ndim):
    return int(lib().gl_op_dofs_per_node(OPS[op], int(ndim)))


def _marshal_eq(eq, keep):
    if _is_torch_tensor(eq):
        import torch

        if eq.dtype != torch.int32 or not eq.is_contiguous():
            raise ValueError("eq must be a contiguous int32 tensor")
        keep.append(eq)
        return eq.data_ptr(), (GL_DEVICE if eq.is_cuda else GL_HOST)
    arr = np.ascontiguousarray(eq, dtype=np.int32)
    keep.append(arr)
    return arr.ctypes.data, GL_HOST


def triplet_indices(ctx, dmesh, op, cell_range, args, eq):
    """(rows, cols) CUDA int32 tensors: global row / column of every double of the batched output of `op` over the range,
    so that (rows, cols, out) is the COO triplet list of a CooMatrix::put loop.  `eq[point*npd + i]` is the point ->
    equation table (negative = not assembled); cols is None for vector cases."""
    import torch

    b, e = _range(dmesh, cell_range)
    a = args._c()
    total = _total(dmesh, op, b, e, a)
    keep = []
    ptr, loc = _marshal_eq(eq, keep)
    dev = torch.device("cuda", ctx.device)
    rows = torch.empty(total, dtype=torch.int32, device=dev)
    cols = torch.empty(total, dtype=torch.int32, device=dev) if op.startswith("mat") else None
    st = lib().gl_triplet_indices(ctx._h, dmesh._h, OPS[op], b, e, C.byref(a), ptr, loc, rows.data_ptr(),
                                  None if cols is None else cols.data_ptr())
    if st:
        ctx._raise(st)
    return rows, cols


def assemble(ctx, dmesh, op, cell_range, args, coef, eq, vals, rowptr=None, colidx=None):
    """Integrates `op` over the range and adds the element blocks into the global CSR matrix (rowptr int64, colidx int32,
    vals float64: CUDA tensors) or, for vector cases, into the global vector `vals`.  Returns `vals`."""
    import torch

    b, e = _range(dmesh, cell_range)
    a = args._c()
    keep = []
    c = None if coef is None else _marshal_coef(coef, keep)
    ptr, loc = _marshal_eq(eq, keep)
    if not (_is_torch_tensor(vals) and vals.is_cuda and vals.dtype == torch.float64 and vals.is_contiguous()):
        raise ValueError("vals must be a contiguous float64 CUDA tensor")
    if op.startswith("mat"):
        if rowptr.dtype != torch.int64 or colidx.dtype != torch.int32 or not (rowptr.is_cuda and colidx.is_cuda):
            raise ValueError("rowptr must be an int64 and colidx an int32 CUDA tensor")
    missing = C.c_uint64()
    st = lib().gl_assemble(ctx._h, dmesh._h, OPS[op], b, e, C.byref(a), None if c is None else C.byref(c), ptr, loc,
                           None if rowptr is None else rowptr.data_ptr(), None if colidx is None else colidx.data_ptr(),
                           vals.data_ptr(), C.byref(missing))
    if st:
        ctx._raise(st)
    return vals


class AssemblyPlan:
    """gl_asm_plan: the slot of every element-block entry in the global CSR matrix / vector, computed once (gl_assemble_plan_create).
    `assemble(coef, vals)` integrates and adds (gl_assemble_planned; the stiffness matrix of solid kinds in ONE kernel per GeoKind,
    element blocks never written to HBM).  Asynchronous on the ctx stream."""

    def __init__(self, ctx, dmesh, op, cell_range, args, eq, rowptr=None, colidx=None):
        b, e = _range(dmesh, cell_range)
        a = args._c()
        keep = []
        ptr, loc = _marshal_eq(eq, keep)
        missing = C.c_uint64()
        h = C.c_void_p()
        st = lib().gl_assemble_plan_create(ctx._h, dmesh._h, OPS[op], b, e, C.byref(a), ptr, loc,
                                           None if rowptr is None else rowptr.data_ptr(), None if colidx is None else colidx.data_ptr(),
                                           C.byref(missing), C.byref(h))
        if st:
            ctx._raise(st)
        self.ctx, self._h, self._dmesh = ctx, h, dmesh

    def assemble(self, coef, vals):
        keep = []
        c = _marshal_coef(coef, keep)
        st = lib().gl_assemble_planned(self.ctx._h, self._h, C.byref(c), vals.data_ptr())
        if st:
            self.ctx._raise(st)
        return vals

    def free(self):
        if self._h:
            lib().gl_assemble_plan_free(self.ctx._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

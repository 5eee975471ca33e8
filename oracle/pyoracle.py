"""ctypes binding of the CPU oracle (oracle/libgemlab_oracle.so).

TEST INFRASTRUCTURE ONLY -- imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (gemlab_b200) must never import this module.

The Python classes mirror the reference's names (Scratchpad, Gauss, CommonArgs, integ.*) so that the
oracle's own tests read like gemlab's Rust tests (src/integ/*.rs `mod tests`).
"""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libgemlab_oracle.so"

KINDS = ["lin2", "lin3", "lin4", "lin5", "tri3", "tri6", "tri10", "tri15", "qua4", "qua8", "qua9", "qua12", "qua16",
         "qua17", "tet4", "tet10", "tet20", "hex8", "hex20", "hex32"]
KIND = {name: i for i, name in enumerate(KINDS)}
CLASS_LIN, CLASS_TRI, CLASS_QUA, CLASS_TET, CLASS_HEX = range(5)
# a representative kind per class, for Gauss::new_sized(class, n)
_CLASS_KIND = {CLASS_LIN: KIND["lin2"], CLASS_TRI: KIND["tri3"], CLASS_QUA: KIND["qua4"], CLASS_TET: KIND["tet4"],
               CLASS_HEX: KIND["hex8"]}

OP_EXT = {"mat_04_nsb": 12, "mat_05_btn": 13, "mat_06_nvn": 14, "mat_07_bsn": 15, "mat_01_nsn_bry": 16, "vec_01_ns_bry": 17,
          "vec_02_nv_bry": 18, "normal": 19}
OP = {"mat_02_bvn": 9, "mat_08_ntn": 10, "mat_09_nvb": 11, "mat_10_bdb": 0, "mat_01_nsn": 1, "mat_03_btb": 2, "vec_01_ns": 3, "vec_03_bv": 4, "vec_02_nv": 5,
      "vec_04_bt": 6, "mat_10_bdb_alt": 7, "jacobian": 8}

GO_MAX_NNODE = 32


def build(force=False):
    """Compile the oracle with the recipe in oracle/Makefile (gcc only)."""
    src_mtime = max((_HERE / f).stat().st_mtime for f in ("gemlab_oracle.c", "gemlab_oracle.h", "gauss_tables.inc"))
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


class _Pad(C.Structure):
    _fields_ = [
        ("kind", C.c_int), ("space_ndim", C.c_int), ("geo_ndim", C.c_int), ("nnode", C.c_int), ("ok_xxt", C.c_int),
        ("xxt", C.c_double * (3 * GO_MAX_NNODE)),
        ("interp", C.c_double * GO_MAX_NNODE),
        ("deriv", C.c_double * (3 * GO_MAX_NNODE)),
        ("jacobian", C.c_double * 9),
        ("inv_jacobian", C.c_double * 9),
        ("gradient", C.c_double * (3 * GO_MAX_NNODE)),
    ]


class _Args(C.Structure):
    _fields_ = [
        ("pad", C.POINTER(_Pad)), ("kind_for_gauss", C.c_int), ("gauss", C.c_void_p), ("ngauss", C.c_int),
        ("clear", C.c_int), ("axisymmetric", C.c_int), ("alpha", C.c_double), ("ii0", C.c_int), ("jj0", C.c_int),
    ]


class _Mesh(C.Structure):
    _fields_ = [
        ("ndim", C.c_int), ("npoint", C.c_size_t), ("ncell", C.c_size_t), ("coords", C.c_void_p),
        ("kinds", C.c_void_p), ("conn_off", C.c_void_p), ("conn", C.c_void_p),
    ]


_CB = C.CFUNCTYPE(C.c_int, C.POINTER(C.c_double), C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_LIB_PATH))
        _lib.go_strerror.restype = C.c_char_p
        _lib.go_kind_name.restype = C.c_char_p
        _lib.go_op_out_size.restype = C.c_size_t
        _lib.go_op_coef_size.restype = C.c_size_t
        _lib.go_lin_elasticity.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_void_p]
        _lib.go_mesh_integ.argtypes = [C.c_int, C.POINTER(_Mesh), C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                       C.c_size_t, C.c_size_t, C.c_double, C.c_int, C.c_void_p, C.c_void_p,
                                       C.c_void_p, C.c_int, C.POINTER(C.c_size_t)]
    return _lib


class OracleError(Exception):
    """Carries the reference's StrError string."""

    def __init__(self, code):
        self.code = code
        super().__init__(lib().go_strerror(code).decode("utf-8"))


def _check(code):
    if code != 0:
        raise OracleError(code)


def kind_id(kind):
    return KIND[kind] if isinstance(kind, str) else int(kind)


def kind_nnode(kind):
    return lib().go_kind_nnode(kind_id(kind))


def kind_geo_ndim(kind):
    return lib().go_kind_geo_ndim(kind_id(kind))


def kind_class(kind):
    return lib().go_kind_class(kind_id(kind))


def reference_coords(kind, m):
    ksi = (C.c_double * 3)()
    _check(lib().go_kind_reference_coords(kind_id(kind), m, ksi))
    return np.array(ksi[: kind_geo_ndim(kind)])


def interp(kind, ksi):
    k = kind_id(kind)
    out = np.zeros(kind_nnode(k))
    ksi3 = np.zeros(3)
    ksi3[: len(ksi)] = ksi
    _check(lib().go_interp(k, ksi3.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return out


def deriv(kind, ksi):
    k = kind_id(kind)
    out = np.zeros((kind_nnode(k), kind_geo_ndim(k)))
    ksi3 = np.zeros(3)
    ksi3[: len(ksi)] = ksi
    _check(lib().go_deriv(k, ksi3.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)))
    return out


class Gauss:
    """gemlab::integ::Gauss (src/integ/gauss.rs:48-70)."""

    def __init__(self, kind, ngauss=0):
        ptr = C.c_void_p()
        n = C.c_int()
        _check(lib().go_gauss(kind_id(kind), int(ngauss), C.byref(ptr), C.byref(n)))
        self._ptr = ptr
        self._n = n.value
        self.data = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(self._n, 4)).copy()

    @staticmethod
    def new(kind):
        return Gauss(kind, 0)

    @staticmethod
    def new_sized(geo_class, ngauss):
        return Gauss(_CLASS_KIND[geo_class], ngauss)

    def npoint(self):
        return self._n

    def coords(self, p):
        return self.data[p]

    def weight(self, p):
        return self.data[p, 3]


class Scratchpad:
    """gemlab::shapes::Scratchpad (src/shapes/scratchpad.rs:61-116)."""

    def __init__(self, space_ndim, kind):
        self._pad = _Pad()
        _check(lib().go_pad_init(C.byref(self._pad), space_ndim, kind_id(kind)))
        self.kind = kind_id(kind)

    @staticmethod
    def new(space_ndim, kind):
        return Scratchpad(space_ndim, kind)

    @property
    def nnode(self):
        return self._pad.nnode

    @property
    def space_ndim(self):
        return self._pad.space_ndim

    @property
    def geo_ndim(self):
        return self._pad.geo_ndim

    @property
    def ok_xxt(self):
        return bool(self._pad.ok_xxt)

    def set_xx(self, m, j, value):
        lib().go_pad_set_xx(C.byref(self._pad), m, j, C.c_double(value))

    def set_coords(self, xx):
        xx = np.asarray(xx, dtype=float)
        for m in range(xx.shape[0]):
            for j in range(xx.shape[1]):
                self.set_xx(m, j, xx[m, j])
        return self

    @property
    def xxt(self):
        return np.array(self._pad.xxt[: self.space_ndim * self.nnode]).reshape(self.space_ndim, self.nnode)

    @property
    def interp(self):
        return np.array(self._pad.interp[: self.nnode])

    @property
    def deriv(self):
        return np.array(self._pad.deriv[: self.nnode * self.geo_ndim]).reshape(self.nnode, self.geo_ndim)

    @property
    def jacobian(self):
        return np.array(self._pad.jacobian[: self.space_ndim * self.geo_ndim]).reshape(self.space_ndim, self.geo_ndim)

    @property
    def inv_jacobian(self):
        return np.array(self._pad.inv_jacobian[: self.space_ndim**2]).reshape(self.space_ndim, self.space_ndim)

    @property
    def gradient(self):
        return np.array(self._pad.gradient[: self.nnode * self.space_ndim]).reshape(self.nnode, self.space_ndim)

    def _ksi(self, ksi):
        k = np.zeros(4)
        k[: len(ksi)] = ksi
        return k

    def calc_interp(self, ksi):
        k = self._ksi(ksi)
        lib().go_pad_calc_interp(C.byref(self._pad), k.ctypes.data_as(C.c_void_p))

    def calc_jacobian(self, ksi):
        k = self._ksi(ksi)
        det = C.c_double()
        _check(lib().go_pad_calc_jacobian(C.byref(self._pad), k.ctypes.data_as(C.c_void_p), C.byref(det)))
        return det.value

    def calc_gradient(self, ksi):
        k = self._ksi(ksi)
        det = C.c_double()
        _check(lib().go_pad_calc_gradient(C.byref(self._pad), k.ctypes.data_as(C.c_void_p), C.byref(det)))
        return det.value

    def calc_coords(self, ksi):
        k = self._ksi(ksi)
        x = np.zeros(3)
        if not self.ok_xxt:
            raise OracleError(1)
        lib().go_pad_calc_coords(C.byref(self._pad), k.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p))
        return x[: self.space_ndim]


class CommonArgs:
    """gemlab::integ::CommonArgs (src/integ/common_args.rs:5-43)."""

    def __init__(self, pad, gauss):
        self.pad = pad
        self.gauss = gauss
        self.clear = True
        self.axisymmetric = False
        self.alpha = 1.0
        self.ii0 = 0
        self.jj0 = 0

    def _c(self):
        a = _Args()
        a.pad = C.pointer(self.pad._pad)
        a.gauss = self.gauss._ptr
        a.ngauss = self.gauss._n
        a.clear = int(self.clear)
        a.axisymmetric = int(self.axisymmetric)
        a.alpha = float(self.alpha)
        a.ii0 = int(self.ii0)
        a.jj0 = int(self.jj0)
        return a


def get_points_coords(pad, gauss):
    """recovery::get_points_coords (src/recovery/get_point_coords.rs:61): real coords of the Gauss points."""
    return [pad.calc_coords(gauss.coords(p)[:3]) for p in range(gauss.npoint())]


def _wrap_cb(fn, out_len, pad, pass_bb=True):
    """Adapt a Python callable fn(out, p, N, B) -> None | raises to the C callback signature."""
    nnode, sd = pad.nnode, pad.space_ndim

    def cb(out_ptr, p, nn_ptr, bb_ptr, _user):
        out = np.ctypeslib.as_array(out_ptr, shape=(out_len,))
        nn = np.ctypeslib.as_array(nn_ptr, shape=(nnode,)) if nn_ptr else None
        bb = np.ctypeslib.as_array(bb_ptr, shape=(nnode, sd)) if (bb_ptr and pass_bb) else None
        try:
            fn(out, p, nn, bb)
        except StopIteration:
            return 1
        return 0

    return _CB(cb)


def _mat(fname, kk, args, fn, coef_len):
    kk_f = np.asfortranarray(kk, dtype=float)
    cb = _wrap_cb(fn, coef_len, args.pad)
    a = args._c()
    code = getattr(lib(), fname)(kk_f.ctypes.data_as(C.c_void_p), kk_f.shape[0], kk_f.shape[1], C.byref(a), cb, None)
    _check(code)
    kk[...] = kk_f
    return kk


def _vec(fname, v, args, fn, coef_len):
    v_c = np.ascontiguousarray(v, dtype=float)
    cb = _wrap_cb(fn, coef_len, args.pad)
    a = args._c()
    code = getattr(lib(), fname)(v_c.ctypes.data_as(C.c_void_p), v_c.shape[0], C.byref(a), cb, None)
    _check(code)
    v[...] = v_c
    return v


def mat_10_bdb(kk, args, fn_dd):
    """fn_dd(dd, p, N, B): dd is the flat (ns*ns) COLUMN-MAJOR Mandel matrix; use dd[:] = D.ravel(order='F')."""
    ns = 2 * args.pad.space_ndim
    return _mat("go_mat_10_bdb", kk, args, fn_dd, ns * ns)


def mat_10_bdb_alt(kk, args, fn_dd):
    ns = 2 * args.pad.space_ndim
    return _mat("go_mat_10_bdb_alt", kk, args, fn_dd, ns * ns)


def mat_01_nsn(kk, args, fn_s):
    """fn_s(s, p, N, B): set s[0]."""
    return _mat("go_mat_01_nsn", kk, args, fn_s, 1)


def mat_03_btb(kk, args, fn_tt):
    """fn_tt(t, p, N, B): t is the Mandel vector (ns)."""
    return _mat("go_mat_03_btb", kk, args, fn_tt, 2 * args.pad.space_ndim)


def mat_02_bvn(kk, args, fn_v):
    """fn_v(v, p, N, B): v has space_ndim entries."""
    return _mat("go_mat_02_bvn", kk, args, fn_v, args.pad.space_ndim)


def mat_08_ntn(kk, args, fn_tt):
    return _mat("go_mat_08_ntn", kk, args, fn_tt, 2 * args.pad.space_ndim)


def mat_09_nvb(kk, args, fn_v):
    return _mat("go_mat_09_nvb", kk, args, fn_v, args.pad.space_ndim)


def vec_01_ns(a, args, fn_s):
    return _vec("go_vec_01_ns", a, args, fn_s, 1)


def vec_02_nv(b, args, fn_v):
    return _vec("go_vec_02_nv", b, args, fn_v, args.pad.space_ndim)


def vec_03_bv(c, args, fn_w):
    return _vec("go_vec_03_bv", c, args, fn_w, args.pad.space_ndim)


def vec_04_bt(d, args, fn_sig):
    return _vec("go_vec_04_bt", d, args, fn_sig, 2 * args.pad.space_ndim)


def scalar_field(pad, gauss, fn_s):
    """fn_s(p) -> float."""
    def fn(out, p, _nn, _bb):
        out[0] = fn_s(p)

    cb = _wrap_cb(fn, 1, pad)
    res = C.c_double()
    _check(lib().go_scalar_field(C.byref(res), C.byref(pad._pad), gauss._ptr, gauss._n, cb, None))
    return res.value


def lin_elasticity(young, poisson, two_dim, plane_stress):
    """LinElasticity::new(E, nu, two_dim, plane_stress).get_modulus().matrix() as an (ns, ns) array."""
    ns = 4 if two_dim else 6
    dd = np.zeros((ns, ns), order="F")
    lib().go_lin_elasticity(young, poisson, int(two_dim), int(plane_stress), dd.ctypes.data_as(C.c_void_p))
    return np.array(dd)


def op_out_size(op, kind, ndim):
    return lib().go_op_out_size(OP[op], kind_id(kind), ndim)


def op_coef_size(op, ndim):
    return lib().go_op_coef_size(OP[op], ndim)


def mesh_integ(op, ndim, coords, kinds, conn_off, conn, cell_begin=0, cell_end=None, ngauss=0, coef=None,
               coef_cell_stride=0, coef_gp_stride=0, alpha=1.0, axisymmetric=False, ksi=None, out_off=None,
               nthreads=1, out=None):
    """The reference's user loop `for cell in cells { mesh.set_pad(..); integ::op(..) }` over a cell range.

    Returns the packed outputs (float64 1-D array).  coords: (npoint, ndim) AoS; kinds: uint8 (ncell);
    conn_off: uint64 (ncell+1); conn: uint64 point ids.
    """
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
    conn_off = np.ascontiguousarray(conn_off, dtype=np.uint64)
    conn = np.ascontiguousarray(conn, dtype=np.uint64)
    ncell = kinds.shape[0]
    if cell_end is None:
        cell_end = ncell
    m = _Mesh(ndim, coords.shape[0], ncell, coords.ctypes.data, kinds.ctypes.data, conn_off.ctypes.data, conn.ctypes.data)
    n = cell_end - cell_begin
    if out_off is None:
        sizes = np.array([lib().go_op_out_size(OP[op], int(k), ndim) for k in np.unique(kinds[cell_begin:cell_end])])
        if len(sizes) > 1:
            per = np.array([lib().go_op_out_size(OP[op], int(k), ndim) for k in kinds[cell_begin:cell_end]], dtype=np.uint64)
            out_off = np.concatenate([[0], np.cumsum(per)]).astype(np.uint64)
            total = int(out_off[-1])
        else:
            total = int(sizes[0]) * n if n else 0
    else:
        out_off = np.ascontiguousarray(out_off, dtype=np.uint64)
        total = int(out_off[-1])
    if out is None:
        out = np.full(total, np.nan)
    coef_arr = None if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
    ksi_arr = None if ksi is None else np.ascontiguousarray(np.concatenate([ksi, np.zeros(4 - len(ksi))]), dtype=np.float64)
    err_cell = C.c_size_t()
    code = lib().go_mesh_integ(OP[op], C.byref(m), cell_begin, cell_end, int(ngauss),
                               None if coef_arr is None else coef_arr.ctypes.data, coef_cell_stride, coef_gp_stride,
                               float(alpha), int(axisymmetric), None if ksi_arr is None else ksi_arr.ctypes.data,
                               out.ctypes.data, None if out_off is None else out_off.ctypes.data, int(nthreads),
                               C.byref(err_cell))
    if code != 0:
        e = OracleError(code)
        e.cell = err_cell.value
        raise e
    return out


def mesh_integ_ext(op, ndim, coords, kinds, conn_off, conn, kind_b=None, cell_begin=0, cell_end=None, ngauss=0, coef=None,
                   coef_cell_stride=0, coef_gp_stride=0, alpha=1.0, axisymmetric=False):
    """The user loop for the two-pad coupling matrices (mat_04_nsb .. mat_07_bsn; kind_b = GeoKind of the second pad, whose nodes are
    the first nnode_b nodes of every cell), the boundary integrals (*_bry) and the batched calc_normal_vector ("normal").
    Coefficient records: see go_mesh_integ_ext in gemlab_oracle.h."""
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
    conn_off = np.ascontiguousarray(conn_off, dtype=np.uint64)
    conn = np.ascontiguousarray(conn, dtype=np.uint64)
    ncell = kinds.shape[0]
    if cell_end is None:
        cell_end = ncell
    m = _Mesh(ndim, coords.shape[0], ncell, coords.ctypes.data, kinds.ctypes.data, conn_off.ctypes.data, conn.ctypes.data)
    kb = -1 if kind_b is None else int(kind_b)
    lib().go_op_out_size_ext.restype = C.c_size_t
    total = sum(int(lib().go_op_out_size_ext(OP_EXT[op], int(k), kb, ndim, int(ngauss))) for k in kinds[cell_begin:cell_end])
    out = np.full(total, np.nan)
    coef_arr = None if coef is None else np.ascontiguousarray(coef, dtype=np.float64)
    err_cell = C.c_size_t()
    code = lib().go_mesh_integ_ext(OP_EXT[op], C.byref(m), C.c_size_t(cell_begin), C.c_size_t(cell_end), int(ngauss), kb,
                                   None if coef_arr is None else C.c_void_p(coef_arr.ctypes.data), C.c_size_t(coef_cell_stride),
                                   C.c_size_t(coef_gp_stride), C.c_double(alpha), int(axisymmetric), C.c_void_p(out.ctypes.data), 1,
                                   C.byref(err_cell))
    if code != 0:
        e = OracleError(code)
        e.cell = err_cell.value
        raise e
    return out


def ncores():
    return len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)

"""GeoKind / GeoClass / Gauss -- host-side mirror of gemlab::shapes::{GeoKind, GeoClass} and gemlab::integ::Gauss.

Reference: src/shapes/enums.rs:84-176 (enums), :217-244 (GeoKind::from), :1467-1493 (VALUES order);
src/integ/gauss.rs:48-70, 87-121 (Gauss::new), 189-246 (Gauss::new_sized).  All data comes from the C library.
"""
import ctypes as C
import enum

import numpy as np

from ._lib import lib


class GeoClass(enum.IntEnum):
    Lin = 0
    Tri = 1
    Qua = 2
    Tet = 3
    Hex = 4

    def ndim(self):
        return (1, 2, 2, 3, 3)[int(self)]


class GeoKind(enum.IntEnum):
    Lin2 = 0
    Lin3 = 1
    Lin4 = 2
    Lin5 = 3
    Tri3 = 4
    Tri6 = 5
    Tri10 = 6
    Tri15 = 7
    Qua4 = 8
    Qua8 = 9
    Qua9 = 10
    Qua12 = 11
    Qua16 = 12
    Qua17 = 13
    Tet4 = 14
    Tet10 = 15
    Tet20 = 16
    Hex8 = 17
    Hex20 = 18
    Hex32 = 19

    @staticmethod
    def from_str(name):
        """GeoKind::from (enums.rs:217-244)."""
        k = lib().gl_kind_from_name(name.encode())
        if k < 0:
            raise ValueError("string representation of GeoKind is incorrect")
        return GeoKind(k)

    def to_string(self):
        return lib().gl_kind_name(int(self)).decode()

    def nnode(self):
        return lib().gl_kind_nnode(int(self))

    def ndim(self):
        return lib().gl_kind_geo_ndim(int(self))

    def geo_class(self):
        return GeoClass(lib().gl_kind_class(int(self)))

    def supported(self):
        return bool(lib().gl_kind_supported(int(self)))

    def interp_deriv(self, ksi):
        """N (nnode) and L = dN/dksi (nnode, geo_ndim) as tabulated for the device."""
        nn = np.zeros(self.nnode())
        ll = np.zeros((self.nnode(), self.ndim()))
        k = np.zeros(3)
        k[: len(ksi)] = ksi
        st = lib().gl_shape_eval(int(self), k.ctypes.data, nn.ctypes.data, ll.ctypes.data)
        if st:
            raise GemlabError(st)
        return nn, ll


_CLASS_KIND = {GeoClass.Lin: GeoKind.Lin2, GeoClass.Tri: GeoKind.Tri3, GeoClass.Qua: GeoKind.Qua4,
               GeoClass.Tet: GeoKind.Tet4, GeoClass.Hex: GeoKind.Hex8}


class GemlabError(Exception):
    """Carries the reference's StrError string (gl_last_error / gl_status_string)."""

    def __init__(self, status, message=None, cell=None):
        self.status = int(status)
        self.cell = cell
        if message is None:
            message = lib().gl_status_string(int(status)).decode("utf-8")
        super().__init__(message)


class Gauss:
    """gemlab::integ::Gauss: rows of {r, s, t, w}."""

    def __init__(self, kind, ngauss=0):
        n = C.c_int()
        ptr = C.c_void_p()
        st = lib().gl_gauss_rule(int(kind), int(ngauss), C.byref(n), C.byref(ptr))
        if st:
            raise GemlabError(st)
        self.kind = GeoKind(int(kind))
        self.data = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(n.value, 4)).copy()

    @staticmethod
    def new(kind):
        return Gauss(kind, 0)

    @staticmethod
    def new_sized(geo_class, ngauss):
        return Gauss(_CLASS_KIND[GeoClass(geo_class)], ngauss)

    @staticmethod
    def new_or_sized(kind, ngauss=None):
        return Gauss(kind, 0 if ngauss is None else ngauss)

    def npoint(self):
        return self.data.shape[0]

    def coords(self, p):
        return self.data[p]

    def weight(self, p):
        return self.data[p, 3]

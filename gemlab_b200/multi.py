"""Device-resident output slabs (struct gl_slab) and the single-process multi-GPU driver (gl_multi_*).

The path shards by contiguous cell ranges with no collective (cells are independent): `MultiContext([0, 1, ..])` owns one
context, stream and host thread per device inside the library; `integ(..)` returns the concatenation in cell order -- bit
identical to the single-device result -- and `integ_slabs(..)` leaves one DeviceSlab per device.
"""
import ctypes as C

import numpy as np

from ._lib import GlSlab, lib
from .geo import GemlabError
from .integ import GL_HOST, OPS, CommonArgs, Coef, _marshal_coef, _marshal_out, _range
from .mesh import Mesh


class DeviceSlab:
    """The batched output of one call, resident on one device: element e of [cell_begin, cell_end) at ptr + off[e - cell_begin]."""

    def __init__(self, c_slab):
        self._c = c_slab

    ptr = property(lambda self: self._c.ptr)
    size = property(lambda self: int(self._c.size))
    device = property(lambda self: int(self._c.device))
    cells = property(lambda self: (int(self._c.cell_begin), int(self._c.cell_end)))

    def download(self, ctx_handle):
        host = np.empty(self.size, dtype=np.float64)
        st = lib().gl_slab_download(ctx_handle, C.byref(self._c), host.ctypes.data)
        if st:
            raise GemlabError(st)
        return host

    def free(self):
        lib().gl_slab_free(C.byref(self._c))

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def integ_slab(ctx, dmesh, op, cell_range, args, coef):
    """gl_integ_slab: integrate into a device slab allocated by the library (asynchronous; ctx.sync() reports zero determinants)."""
    b, e = _range(dmesh, cell_range)
    keep = []
    a = args._c()
    c = _marshal_coef(coef, keep)
    slab = GlSlab()
    st = lib().gl_integ_slab(ctx._h, dmesh._h, OPS[op], b, e, C.byref(a), C.byref(c), C.byref(slab))
    if st:
        ctx._raise(st)
    return DeviceSlab(slab)


class MultiMesh:
    def __init__(self, multi, handle, mesh):
        self.multi, self._h = multi, handle
        self.ndim, self.ncell, self.npoint = mesh.ndim, mesh.ncell, mesh.npoint

    def free(self):
        if self._h:
            lib().gl_multi_mesh_free(self.multi._h, self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MultiContext:
    """gl_multi: `devices` may name a device more than once (each entry gets its own context and stream)."""

    def __init__(self, devices):
        devs = (C.c_int * len(devices))(*[int(d) for d in devices])
        h = C.c_void_p()
        st = lib().gl_multi_create(devs, len(devices), C.byref(h))
        if st:
            raise GemlabError(st)
        self._h = h
        self.ndevice = len(devices)

    def close(self):
        if self._h:
            lib().gl_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, st):
        raise GemlabError(st, lib().gl_multi_last_error(self._h).decode("utf-8") or None)

    def ctx_handle(self, g):
        return lib().gl_multi_ctx(self._h, g)

    def last_kernel(self, g):
        return lib().gl_ctx_last_kernel(self.ctx_handle(g)).decode("ascii")

    @staticmethod
    def partition(cell_begin, cell_end, ndevice, g):
        pb, pe = C.c_uint64(), C.c_uint64()
        lib().gl_multi_partition(cell_begin, cell_end, ndevice, g, C.byref(pb), C.byref(pe))
        return pb.value, pe.value

    def upload(self, mesh):
        if not isinstance(mesh, Mesh):
            raise TypeError("expected a gemlab_b200.Mesh")
        h = C.c_void_p()
        st = lib().gl_multi_mesh_upload(self._h, mesh.ndim, mesh.npoint, mesh.coords.ctypes.data, mesh.ncell, mesh.kinds.ctypes.data,
                                        mesh.markers.ctypes.data, mesh.conn_off.ctypes.data, mesh.conn.ctypes.data, C.byref(h))
        if st:
            self._raise(st)
        return MultiMesh(self, h, mesh)

    def _total(self, mmesh, op, b, e, a):
        total = C.c_uint64()
        st = lib().gl_out_total(lib().gl_multi_mesh_replica(mmesh._h, 0), OPS[op], b, e, C.byref(a), C.byref(total))
        if st:
            raise GemlabError(st)
        return total.value

    def integ(self, op, mmesh, cell_range=None, args=None, coef=None, out=None):
        """gl_multi_integ: host output of the whole range, concatenated in cell order."""
        args = args or CommonArgs()
        b, e = _range(mmesh, cell_range)
        a = args._c()
        keep = []
        c = _marshal_coef(coef if coef is not None else Coef.uniform([0.0]), keep)
        o, result = _marshal_out(out, self._total(mmesh, op, b, e, a), args.clear)
        if o.location != GL_HOST:
            raise ValueError("MultiContext.integ writes host memory; use integ_slabs for device outputs")
        st = lib().gl_multi_integ(self._h, mmesh._h, OPS[op], b, e, C.byref(a), C.byref(c), C.byref(o))
        if st:
            self._raise(st)
        return result

    def integ_slabs(self, op, mmesh, cell_range=None, args=None, coef=None):
        """gl_multi_integ_slabs + gl_multi_sync: one DeviceSlab per device, in piece order."""
        args = args or CommonArgs()
        b, e = _range(mmesh, cell_range)
        a = args._c()
        keep = []
        c = _marshal_coef(coef if coef is not None else Coef.uniform([0.0]), keep)
        slabs = (GlSlab * self.ndevice)()
        st = lib().gl_multi_integ_slabs(self._h, mmesh._h, OPS[op], b, e, C.byref(a), C.byref(c), slabs)
        if st:
            self._raise(st)
        st = lib().gl_multi_sync(self._h)
        if st:
            self._raise(st)
        out = []
        for g in range(self.ndevice):
            s = GlSlab()
            C.memmove(C.byref(s), C.byref(slabs[g]), C.sizeof(GlSlab))
            out.append(DeviceSlab(s))
        return out

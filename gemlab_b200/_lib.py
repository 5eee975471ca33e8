"""ctypes loader of libgemlab_cuda.so (the C ABI in include/gemlab_cuda.h).

The product path has NO CPU fallback: if the shared library is missing this module raises, and if no sm_100 device is
usable `Context()` raises (GL_ERR_NO_DEVICE).  Nothing here imports the oracle.
"""
import ctypes as C
from pathlib import Path

import os

_HERE = Path(__file__).resolve().parent
# GEMLAB_CUDA_TUNING=1 selects the -DGL_TUNING build (make -C gemlab_b200/csrc TUNING=1), the only one that reads the GL_*
# environment knobs; used by the sweep tools, never by tests or bench.py
LIB_PATH = _HERE / ("libgemlab_cuda_tuning.so" if os.environ.get("GEMLAB_CUDA_TUNING") == "1" else "libgemlab_cuda.so")


class GlArgs(C.Structure):
    """struct gl_args (mirrors integ::CommonArgs, src/integ/common_args.rs:5-43)."""

    _fields_ = [
        ("ngauss", C.c_int32), ("axisymmetric", C.c_int32), ("clear", C.c_int32), ("packed", C.c_int32),
        ("alpha", C.c_double), ("ii0", C.c_uint32), ("jj0", C.c_uint32), ("nrow", C.c_uint32), ("ncol", C.c_uint32),
        ("ksi", C.c_double * 3), ("kind_b", C.c_int32), ("reserved", C.c_int32),
    ]


class GlCoef(C.Structure):
    _fields_ = [("mode", C.c_int32), ("location", C.c_int32), ("data", C.c_void_p), ("markers", C.c_void_p),
                ("nrecord", C.c_uint64)]


class GlOut(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("location", C.c_int32), ("reserved", C.c_int32), ("capacity", C.c_uint64)]


class GlFusedItem(C.Structure):
    _fields_ = [("op", C.c_int32), ("reserved", C.c_int32), ("coef", C.POINTER(GlCoef)), ("out", C.POINTER(GlOut))]


class GlSlab(C.Structure):
    """struct gl_slab: a batched output that stays on its device."""

    _fields_ = [("ptr", C.c_void_p), ("capacity", C.c_uint64), ("size", C.c_uint64), ("cell_begin", C.c_uint64),
                ("cell_end", C.c_uint64), ("device", C.c_int32), ("owned", C.c_int32)]


# every symbol include/gemlab_cuda.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "gl_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "gl_ctx_destroy": (None, [_P]),
    "gl_ctx_set_stream": (C.c_int, [_P, _P]),
    "gl_ctx_sync": (C.c_int, [_P]),
    "gl_last_error": (C.c_char_p, [_P]),
    "gl_last_error_cell": (C.c_uint64, [_P]),
    "gl_status_string": (C.c_char_p, [C.c_int]),
    "gl_ctx_launch_count": (C.c_uint64, [_P]),
    "gl_ctx_last_kernel": (C.c_char_p, [_P]),
    "gl_kind_nnode": (C.c_int, [C.c_int]),
    "gl_kind_geo_ndim": (C.c_int, [C.c_int]),
    "gl_kind_class": (C.c_int, [C.c_int]),
    "gl_kind_from_name": (C.c_int, [C.c_char_p]),
    "gl_kind_name": (C.c_char_p, [C.c_int]),
    "gl_kind_supported": (C.c_int, [C.c_int]),
    "gl_gauss_rule": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(_P)]),
    "gl_shape_eval": (C.c_int, [C.c_int, _P, _P, _P]),
    "gl_kind_reference_coords": (C.c_int, [C.c_int, C.c_int, _P]),
    "gl_extrap_matrix": (C.c_int, [C.c_int, C.c_int, _P]),
    "gl_mesh_upload": (C.c_int, [_P, C.c_int, C.c_uint64, _P, C.c_uint64, _P, _P, _P, _P, C.POINTER(_P)]),
    "gl_mesh_upload_homogeneous": (C.c_int, [_P, C.c_int, C.c_uint64, _P, C.c_uint64, C.c_int, _P, _P, C.POINTER(_P)]),
    "gl_mesh_update_coords": (C.c_int, [_P, _P, _P, C.c_int]),
    "gl_mesh_block": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int32, C.POINTER(_P)]),
    "gl_mesh_block_general": (C.c_int, [_P, C.c_int, _P, C.c_int, _P, C.c_int32, C.POINTER(_P)]),
    "gl_mesh_download": (C.c_int, [_P, _P, _P, _P]),
    "gl_mesh_free": (None, [_P, _P]),
    "gl_mesh_ncell": (C.c_uint64, [_P]),
    "gl_mesh_npoint": (C.c_uint64, [_P]),
    "gl_mesh_ndim": (C.c_int, [_P]),
    "gl_op_out_size": (C.c_uint64, [C.c_int, C.c_int, C.c_int]),
    "gl_op_coef_size": (C.c_uint64, [C.c_int, C.c_int]),
    "gl_out_total": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(C.c_uint64)]),
    "gl_out_offsets": (C.c_int, [_P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), _P]),
    "gl_integ": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlCoef), C.POINTER(GlOut)]),
    "gl_jacobian": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlOut)]),
    "gl_integ_fused": (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlFusedItem), C.c_int]),
    "gl_args_default": (None, [C.POINTER(GlArgs)]),
    "gl_lin_elasticity": (None, [C.c_double, C.c_double, C.c_int, C.c_int, _P]),
    "gl_op_dofs_per_node": (C.c_int, [C.c_int, C.c_int]),
    "gl_triplet_indices": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), _P, C.c_int, _P, _P]),
    "gl_assemble": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlCoef), _P, C.c_int, _P, _P, _P,
                              C.POINTER(C.c_uint64)]),
    "gl_bench_fp64_peak": (C.c_int, [_P, C.c_int, C.c_int, C.POINTER(C.c_double)]),
    "gl_bench_hbm_write": (C.c_int, [_P, C.c_int, _P, C.c_uint64, C.c_int, C.POINTER(C.c_double)]),
}
SYMBOLS.update({
    "gl_integ_slab": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlCoef), C.POINTER(GlSlab)]),
    "gl_slab_download": (C.c_int, [_P, C.POINTER(GlSlab), _P]),
    "gl_slab_free": (None, [C.POINTER(GlSlab)]),
    "gl_multi_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_P)]),
    "gl_multi_destroy": (None, [_P]),
    "gl_multi_ndevice": (C.c_int, [_P]),
    "gl_multi_ctx": (_P, [_P, C.c_int]),
    "gl_multi_last_error": (C.c_char_p, [_P]),
    "gl_multi_partition": (None, [C.c_uint64, C.c_uint64, C.c_int, C.c_int, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "gl_multi_mesh_upload": (C.c_int, [_P, C.c_int, C.c_uint64, _P, C.c_uint64, _P, _P, _P, _P, C.POINTER(_P)]),
    "gl_multi_mesh_free": (None, [_P, _P]),
    "gl_multi_mesh_replica": (_P, [_P, C.c_int]),
    "gl_multi_integ": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlCoef), C.POINTER(GlOut)]),
    "gl_multi_integ_slabs": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlCoef), C.POINTER(GlSlab)]),
    "gl_multi_sync": (C.c_int, [_P]),
    "gl_host_mesh_parse": (C.c_int, [C.c_char_p, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "gl_host_mesh_read": (C.c_int, [C.c_char_p, C.c_int, C.POINTER(_P)]),
    "gl_host_mesh_free": (None, [_P]),
    "gl_mesh_text_last_error": (C.c_char_p, []),
    "gl_host_mesh_ndim": (C.c_int, [_P]),
    "gl_host_mesh_npoint": (C.c_uint64, [_P]),
    "gl_host_mesh_ncell": (C.c_uint64, [_P]),
    "gl_host_mesh_nmarked_edge": (C.c_uint64, [_P]),
    "gl_host_mesh_nmarked_face": (C.c_uint64, [_P]),
    "gl_host_mesh_coords": (C.POINTER(C.c_double), [_P]),
    "gl_host_mesh_point_markers": (C.POINTER(C.c_int32), [_P]),
    "gl_host_mesh_kinds": (C.POINTER(C.c_uint8), [_P]),
    "gl_host_mesh_markers": (C.POINTER(C.c_int32), [_P]),
    "gl_host_mesh_conn_off": (C.POINTER(C.c_uint64), [_P]),
    "gl_host_mesh_conn": (C.POINTER(C.c_uint64), [_P]),
    "gl_host_mesh_marked_edges": (C.POINTER(C.c_int64), [_P]),
    "gl_host_mesh_marked_faces": (C.POINTER(C.c_int64), [_P]),
    "gl_mesh_upload_host": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "gl_mesh_from_text": (C.c_int, [_P, C.c_char_p, C.c_uint64, C.c_int, C.POINTER(_P)]),
    "gl_assemble_plan_create": (C.c_int, [_P, _P, C.c_int, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), _P, C.c_int, _P, _P, C.POINTER(C.c_uint64),
                                          C.POINTER(_P)]),
    "gl_assemble_planned": (C.c_int, [_P, _P, C.POINTER(GlCoef), _P]),
    "gl_assemble_plan_free": (None, [_P, _P]),
})
SYMBOLS["gl_normal_vectors"] = (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlOut)])
for _name in ("gl_mat_04_nsb", "gl_mat_05_btn", "gl_mat_06_nvn", "gl_mat_07_bsn", "gl_mat_01_nsn_bry", "gl_vec_01_ns_bry", "gl_vec_02_nv_bry",
              "gl_mat_10_bdb", "gl_mat_01_nsn", "gl_mat_02_bvn", "gl_mat_03_btb", "gl_mat_08_ntn", "gl_mat_09_nvb", "gl_vec_01_ns", "gl_vec_02_nv", "gl_vec_03_bv",
              "gl_vec_04_bt", "gl_mat_10_gdg", "gl_vec_10_gs"):
    SYMBOLS[_name] = (C.c_int, [_P, _P, C.c_uint64, C.c_uint64, C.POINTER(GlArgs), C.POINTER(GlCoef), C.POINTER(GlOut)])

_lib = None


def lib():
    """Loads libgemlab_cuda.so; raises if it has not been built (python __graft_entry__.py or make -C gemlab_b200/csrc)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        handle = C.CDLL(str(LIB_PATH))
        for name, (restype, argtypes) in SYMBOLS.items():
            fn = getattr(handle, name)  # AttributeError if the library does not export a declared symbol
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib

"""gemlab_b200 -- B200-native batched element integration behind gemlab's `integ` interface.

Host-side mirror (Python, over the C ABI of libgemlab_cuda.so) of the reference's types on the hot path:
GeoKind / GeoClass / Gauss (shapes, integ), Mesh / Block (mesh), CommonArgs + integ.* (integ).
The compute path is hand-written sm_100a CUDA (gemlab_b200/csrc); there is no CPU fallback.
"""
from . import integ
from ._lib import LIB_PATH, lib
from .geo import Gauss, GemlabError, GeoClass, GeoKind
from .integ import CommonArgs, Coef, Context, DeviceMesh, lin_elasticity, mandel_matrix, partition
from .multi import DeviceSlab, MultiContext, integ_slab
from .mesh import Block, GemlabTextError, Mesh, jitter_interior_points, split_hex_cells_into_tet10

__all__ = ["integ", "lib", "LIB_PATH", "Gauss", "GemlabError", "GeoClass", "GeoKind", "CommonArgs", "Coef", "Context",
           "DeviceMesh", "DeviceSlab", "MultiContext", "integ_slab", "lin_elasticity", "mandel_matrix", "partition", "Block", "Mesh", "GemlabTextError", "jitter_interior_points",
           "split_hex_cells_into_tet10"]

"""Single-process multi-GPU driver (gl_multi_*) and device-resident slabs (gl_slab): SURVEY.md 8(e).  The cell range is split into
contiguous pieces, one per device context; the concatenation must be BIT-identical to the single-context result (no collective,
no change of arithmetic).  On a one-GPU box the driver is exercised with several contexts on device 0."""
import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Coef, CommonArgs, GeoKind, MultiContext, integ, integ_slab

from test_gpu_parity import make_mesh, random_coef

pytestmark = pytest.mark.gpu


def device_lists():
    import torch

    n = torch.cuda.device_count()
    lists = [[0], [0, 0, 0]]
    if n >= 2:
        lists.append(list(range(n)))
    return lists


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("kind,op", [("hex8", "mat_10_bdb"), ("tet10", "mat_01_nsn"), ("qua8", "mat_03_btb"), ("hex20", "mat_10_bdb"),
                                     ("qua4", "vec_02_nv")])
def test_multi_device_concatenation_is_bit_identical(ctx, kind, op):
    mesh = make_mesh(kind, n=4)
    rng = np.random.default_rng(9)
    dmesh = ctx.upload(mesh)
    ng = g.Gauss.new(GeoKind.from_str(kind)).npoint()
    coefs = {
        "uniform": Coef.uniform(random_coef(op, mesh.ndim, 1, rng)[0]),
        "per_cell": Coef.per_cell(random_coef(op, mesh.ndim, mesh.ncell, rng)),
        "per_gauss": Coef.per_gauss(random_coef(op, mesh.ndim, mesh.ncell * ng, rng)),
    }
    for devices in device_lists():
        multi = MultiContext(devices)
        mm = multi.upload(mesh)
        for name, coef in coefs.items():
            for rng_cells in [None, (3, mesh.ncell - 5)]:
                ref = getattr(integ, op)(ctx, dmesh, rng_cells, CommonArgs(), coef if rng_cells is None else _slice(coef, rng_cells, ng, name))
                got = multi.integ(op, mm, rng_cells, CommonArgs(), coef if rng_cells is None else _slice(coef, rng_cells, ng, name))
                assert np.array_equal(got, ref), f"{kind} {op} {name} devices={devices} range={rng_cells}"
        # device-resident pieces: one slab per device, downloaded and concatenated
        slabs = multi.integ_slabs(op, mm, None, CommonArgs(), coefs["per_cell"])
        ref = getattr(integ, op)(ctx, dmesh, None, CommonArgs(), coefs["per_cell"])
        parts = [s.download(multi.ctx_handle(i)) for i, s in enumerate(slabs)]
        assert np.array_equal(np.concatenate(parts), ref)
        cells = [s.cells for s in slabs]
        assert cells == [MultiContext.partition(0, mesh.ncell, len(devices), i) for i in range(len(devices))]
        assert cells[0][0] == 0 and cells[-1][1] == mesh.ncell
        for s in slabs:
            s.free()
        mm.free()
        multi.close()


def _slice(coef, cells, ng, name):
    b, e = cells
    if name == "per_cell":
        return Coef.per_cell(coef.data[b:e])
    if name == "per_gauss":
        return Coef.per_gauss(coef.data[b * ng:e * ng])
    return coef


def test_partition_rule_matches_the_python_helper():
    for n, G in [(0, 3), (1, 4), (7, 3), (8000000, 8), (4983504, 7)]:
        for r in range(G):
            assert MultiContext.partition(5, 5 + n, G, r) == tuple(5 + v for v in g.partition(n, G, r))


def test_multi_device_errors(ctx):
    mesh = make_mesh("hex8", n=3)
    multi = MultiContext([0, 0])
    mm = multi.upload(mesh)
    with pytest.raises(g.GemlabError):
        multi.integ("mat_10_bdb", mm, (0, mesh.ncell + 1), CommonArgs(), Coef.uniform(np.eye(6).ravel()))
    # a collapsed cell in the second piece: the smallest failing cell id is reported through that device's ctx
    coords = mesh.coords.copy()
    conn = mesh.conn_matrix()
    e_bad = mesh.ncell - 2
    coords[conn[e_bad].astype(np.int64)] = coords[int(conn[e_bad][0])]
    flat = g.Mesh.homogeneous(3, coords, GeoKind.Hex8, conn)
    mm2 = multi.upload(flat)
    with pytest.raises(g.GemlabError) as err:
        multi.integ("mat_10_bdb", mm2, None, CommonArgs(), Coef.uniform(np.eye(6).ravel()))
    assert "cannot compute inverse due to zero determinant" in str(err.value)
    mm.free()
    mm2.free()
    multi.close()


def test_single_context_slab(ctx):
    mesh = make_mesh("tet10", n=3)
    dmesh = ctx.upload(mesh)
    dd = g.mandel_matrix(g.lin_elasticity(100.0, 0.25, False))
    ref = integ.mat_10_bdb(ctx, dmesh, (2, 30), CommonArgs(), Coef.uniform(dd))
    slab = integ_slab(ctx, dmesh, "mat_10_bdb", (2, 30), CommonArgs(), Coef.uniform(dd))
    assert slab.cells == (2, 30) and slab.size == ref.size and slab.device == 0
    assert np.array_equal(slab.download(ctx._h), ref)
    slab.free()

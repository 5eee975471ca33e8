"""Device-side lattice generator (gl_mesh_block) against the host Block, whose cell order and first-touch point numbering
are pinned by the reference's sample meshes (tests/golden/block_samples.json <- src/mesh/samples.rs:1407, 1901)."""
import time

import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Block, Coef, CommonArgs, GeoKind, integ

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("kind,ndiv", [("hex8", (2, 2, 2)), ("hex8", (5, 3, 4)), ("hex8", (1, 1, 1)), ("hex8", (7, 1, 2)),
                                       ("qua4", (2, 2)), ("qua4", (6, 5)), ("qua4", (1, 9))])
def test_generated_mesh_equals_host_block(ctx, kind, ndiv):
    k = GeoKind.from_str(kind)
    host = (Block.new_cube(2.0) if k.ndim() == 3 else Block.new_square(2.0)).set_ndiv(list(ndiv)).subdivide(k)
    dm = ctx.block(k, ndiv, xmin=[0.0] * k.ndim(), xmax=[2.0] * k.ndim())
    got = ctx.download(dm)
    assert np.array_equal(got.conn_matrix(), host.conn_matrix())          # ordering contract: exact
    assert np.max(np.abs(got.coords - host.coords)) <= 4e-16 * 2.0        # coordinates: same lattice
    # and it integrates like the uploaded host mesh
    dd = g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, k.ndim() == 2))
    a = integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(dd))
    b = integ.mat_10_bdb(ctx, ctx.upload(got), None, CommonArgs(), Coef.uniform(dd))
    assert np.array_equal(a, b)


def test_first_touch_numbering_of_the_reference_sample(ctx):
    # Samples::block_3d_eight_hex8 (src/mesh/samples.rs:1901): cell 1 of the 2x2x2 block is [1, 8, 9, 2, 5, 10, 11, 6]
    got = ctx.download(ctx.block(GeoKind.Hex8, (2, 2, 2)))
    assert list(got.conn_matrix()[1]) == [1, 8, 9, 2, 5, 10, 11, 6]
    assert got.npoint == 27


def test_large_block_is_generated_without_host_arrays(ctx):
    t0 = time.perf_counter()
    dm = ctx.block(GeoKind.Hex8, (100, 100, 100))
    ctx.sync()
    dt = time.perf_counter() - t0
    assert dm.ncell == 1_000_000 and dm.npoint == 101 ** 3
    host = Block.new_cube(1.0).set_ndiv([100, 100, 100]).subdivide(GeoKind.Hex8)
    got = ctx.download(dm)
    assert np.array_equal(got.conn_matrix(), host.conn_matrix())
    assert dt < 1.0


@pytest.mark.parametrize("kind,ndiv", [("qua8", (2, 2)), ("qua8", (5, 3)), ("qua9", (2, 2)), ("qua9", (4, 7)), ("hex20", (2, 2, 2)),
                                       ("hex20", (3, 1, 4)), ("qua4", (3, 2)), ("hex8", (2, 3, 2))])
def test_quadratic_targets_and_general_blocks(ctx, kind, ndiv):
    """gl_mesh_block for the quadratic targets and gl_mesh_block_general for straight-edged and curved blocks against the host
    Block (first-touch numbering pinned by Samples::block_2d_four_qua8 / qua9 / block_3d_eight_hex20, src/mesh/samples.rs:1466,
    1530, 2043 through tests/golden/block_samples.json)."""
    k = GeoKind.from_str(kind)
    nd = k.ndim()
    host = (Block.new_cube(2.0) if nd == 3 else Block.new_square(2.0)).set_ndiv(list(ndiv)).subdivide(k)
    got = ctx.download(ctx.block(k, ndiv, xmin=[0.0] * nd, xmax=[2.0] * nd))
    assert got.npoint == host.npoint
    assert np.array_equal(got.conn_matrix(), host.conn_matrix())
    assert np.max(np.abs(got.coords - host.coords)) <= 1e-15 * 2.0
    # a skewed straight-edged block given by its corners
    rng = np.random.default_rng(3)
    ref = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]], dtype=float)
    corners = (ref[:4, :2] if nd == 2 else ref) + 0.2 * rng.uniform(-1, 1, (4 if nd == 2 else 8, nd))
    host = Block(_block_nodes(corners, nd)).set_ndiv(list(ndiv)).subdivide(k)
    got = ctx.download(ctx.block_general(k, ndiv, corners))
    assert np.array_equal(got.conn_matrix(), host.conn_matrix())
    assert np.max(np.abs(got.coords - host.coords)) <= 1e-14
    # a curved block (mid-edge nodes off their edges): the Qua8 / Hex20 serendipity map
    nodes = _block_nodes(corners, nd)
    nodes[4 if nd == 2 else 8:] += 0.05 * rng.uniform(-1, 1, nodes[4 if nd == 2 else 8:].shape)
    host = Block(nodes).set_ndiv(list(ndiv)).subdivide(k)
    got = ctx.download(ctx.block_general(k, ndiv, nodes))
    assert np.array_equal(got.conn_matrix(), host.conn_matrix())
    assert np.max(np.abs(got.coords - host.coords)) <= 1e-14


def _block_nodes(corners, nd):
    """Qua8 / Hex20 nodes of a straight-edged block: corners plus edge midpoints in the reference's local order."""
    if nd == 2:
        edges = [(0, 1), (1, 2), (2, 3), (3, 0)]
    else:
        edges = [(0, 1), (1, 2), (2, 3), (3, 0), (4, 5), (5, 6), (6, 7), (7, 4), (0, 4), (1, 5), (2, 6), (3, 7)]
    return np.vstack([corners, [(corners[a] + corners[b]) / 2 for a, b in edges]])


def test_hex20_block_of_config_2_is_generated_on_the_device(ctx):
    dm = ctx.block(GeoKind.Hex20, (40, 40, 40))
    ctx.sync()
    host = Block.new_cube(1.0).set_ndiv([40, 40, 40]).subdivide(GeoKind.Hex20)
    got = ctx.download(dm)
    assert np.array_equal(got.conn_matrix(), host.conn_matrix())

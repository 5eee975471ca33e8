"""MSH text reader / writer (gemlab_b200.Mesh.from_text / read / to_text): the grammar and error strings of the
reference's reader (src/mesh/read_text_mesh.rs, README.md "MSH file format").  Meshes below are written for this test in
the documented format; when the reference tree is present (authoring container only) its own data/meshes/*.msh are read
as well.  The GPU test integrates a mesh that arrived through the text format."""
from pathlib import Path

import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Coef, CommonArgs, GemlabTextError, GeoKind, Mesh, integ

TWO_CUBES = """
# two stacked unit cubes, the second one with another marker
# ndim npoint ncell nmarked_edge nmarked_face
     3     12     2            2            2

# id marker x y z
   0  0 0.0 0.0 0.0
   1 -1 1.0 0.0 0.0
   2 -1 1.0 1.0 0.0
   3  0 0.0 1.0 0.0
   4  0 0.0 0.0 1.0
   5  0 1.0 0.0 1.0
   6  0 1.0 1.0 1.0
   7 -1 0.0 1.0 1.0
   8  0 0.0 0.0 2.0
   9  0 1.0 0.0 2.0
  10  0 1.0 1.0 2.0
  11  0 0.0 1.0 2.0

# id marker kind points
   0  1  hex8  0 1 2 3 4 5  6  7
   1  2  hex8  4 5 6 7 8 9 10 11

# marker p1 p2
123 7 11
-5 11 10

# marker p1 p2 p3 {p4}
-8 3 2 7 6
-9 8 10 9
"""

MIXED_2D = """
2 6 3 0 0
0 0 0.0 0.0
1 0 1.0 0.0
2 0 2.0 0.0
3 0 0.0 1.0
4 0 1.0 1.0
5 0 2.0 1.5
0 1 qua4 0 1 4 3
1 2 tri3 1 2 4
2 2 tri3 2 5 4
"""


def test_reads_sections_markers_and_mixed_kinds():
    m = Mesh.from_text(TWO_CUBES)
    assert (m.ndim, m.npoint, m.ncell) == (3, 12, 2)
    assert list(m.kinds) == [int(GeoKind.Hex8)] * 2 and list(m.markers) == [1, 2]
    assert list(m.cell_points(1)) == [4, 5, 6, 7, 8, 9, 10, 11]
    assert list(m.point_markers[:3]) == [0, -1, -1] and m.coords[10].tolist() == [1.0, 1.0, 2.0]
    assert m.marked_edges == [(123, 7, 11), (-5, 11, 10)]
    assert m.marked_faces == [(-8, 3, 2, 7, 6), (-9, 8, 10, 9)]
    m2 = Mesh.from_text(MIXED_2D)
    assert [GeoKind(int(k)).to_string() for k in m2.kinds] == ["qua4", "tri3", "tri3"]
    assert list(m2.conn_off) == [0, 4, 7, 10]


def test_write_then_read_round_trip():
    for text in (TWO_CUBES, MIXED_2D):
        a = Mesh.from_text(text)
        b = Mesh.from_text(a.to_text())
        assert np.array_equal(a.coords, b.coords) and np.array_equal(a.conn, b.conn) and np.array_equal(a.kinds, b.kinds)
        assert np.array_equal(a.markers, b.markers) and a.marked_edges == b.marked_edges and a.marked_faces == b.marked_faces


@pytest.mark.parametrize("text,msg", [
    ("", "file is empty or header is missing"),
    ("# only comments\n", "file is empty or header is missing"),
    ("2 4", "cannot read ncell"),
    ("2 x 1 0 0", "cannot parse npoint"),
    ("2 2 1 0 0\n0 0 0.0 0.0\n", "not all points have been found"),
    ("2 2 1 0 0\n0 0 0.0 0.0\n2 0 1.0 0.0\n", "the id and index of points must equal each other"),
    ("2 1 0 0 0\n0 0 0.0 0.0 0.0\n", "point data contains extra values"),
    ("2 1 0 0 0\n0 0 0.0\n", "cannot read point y coordinate"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n", "not all cells have been found"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri4 0 1 2\n", "string representation of GeoKind is incorrect"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1\n", "cannot read cell point id"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1 2 3\n", "cell data contains extra values"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n1 1 tri3 0 1 2\n", "the id and index of cells must equal each other"),
    ("2 3 1 1 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1 2\n", "not all marked edges have been found"),
    ("2 3 1 0 1\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1 2\n-1 0 1\n", "cannot read p3 of marked face"),
])
def test_error_strings_are_the_references(text, msg):
    with pytest.raises(GemlabTextError) as e:
        Mesh.from_text(text)
    assert str(e.value) == msg


def _same(a, b):
    return (a.ndim == b.ndim and np.array_equal(a.coords, b.coords) and np.array_equal(a.conn, b.conn) and np.array_equal(a.conn_off, b.conn_off)
            and np.array_equal(a.kinds, b.kinds) and np.array_equal(a.markers, b.markers) and np.array_equal(a.point_markers, b.point_markers)
            and a.marked_edges == b.marked_edges and a.marked_faces == b.marked_faces)


def test_c_abi_parser_agrees_with_the_python_reader():
    """gl_host_mesh_parse (the parser a Rust / C++ host links) against the Python reader, MSH and JSON."""
    for text in (TWO_CUBES, MIXED_2D):
        a = Mesh.from_text(text)
        b = Mesh.from_text_native(text)
        assert _same(a, b)
        c = Mesh.from_text_native(a.to_json(), "json")  # Mesh::write_json -> Mesh::read_json round trip
        assert _same(a, c)
    js = Mesh.from_text(TWO_CUBES).to_json()
    assert '"kind": "Hex8"' in js and "18446744073709551615" in js  # serde's image: variant names, usize::MAX for a missing p4


@pytest.mark.parametrize("text,msg", [
    ("", "file is empty or header is missing"),
    ("2 4", "cannot read ncell"),
    ("2 x 1 0 0", "cannot parse npoint"),
    ("2 2 1 0 0\n0 0 0.0 0.0\n", "not all points have been found"),
    ("2 2 1 0 0\n0 0 0.0 0.0\n2 0 1.0 0.0\n", "the id and index of points must equal each other"),
    ("2 1 0 0 0\n0 0 0.0 0.0 0.0\n", "point data contains extra values"),
    ("2 1 0 0 0\n0 0 0.0\n", "cannot read point y coordinate"),
    ("2 1 0 0 0\n0 0 0.0 abc\n", "cannot parse point y coordinate"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n", "not all cells have been found"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri4 0 1 2\n", "string representation of GeoKind is incorrect"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1\n", "cannot read cell point id"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1 2 3\n", "cell data contains extra values"),
    ("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n1 1 tri3 0 1 2\n", "the id and index of cells must equal each other"),
    ("2 3 1 1 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1 2\n", "not all marked edges have been found"),
    ("2 3 1 0 1\n0 0 0 0\n1 0 1 0\n2 0 0 1\n0 1 tri3 0 1 2\n-1 0 1\n", "cannot read p3 of marked face"),
])
def test_c_abi_parser_error_strings(text, msg):
    with pytest.raises(GemlabTextError) as e:
        Mesh.from_text_native(text)
    assert str(e.value) == msg


def test_c_abi_json_errors_and_files(tmp_path):
    for bad in ("", "{", '{"ndim": 2}', '{"ndim":2,"points":[],"cells":[{"id":0,"marker":1,"kind":"Tri9","points":[0,1,2]}],"marked_edges":[],"marked_faces":[]}'):
        with pytest.raises(GemlabTextError) as e:
            Mesh.from_text_native(bad, "json")
        assert str(e.value) == "cannot parse JSON file"  # Mesh::read_json maps every serde error to this string (mesh.rs:253)
    with pytest.raises(GemlabTextError) as e:
        Mesh.read_native("/nonexistent/dir/mesh.json")
    assert str(e.value) == "cannot open file"
    m = Mesh.from_text(TWO_CUBES)
    (tmp_path / "m.json").write_text(m.to_json())
    (tmp_path / "m.msh").write_text(m.to_text())
    assert _same(m, Mesh.read_native(tmp_path / "m.json")) and _same(m, Mesh.read_native(tmp_path / "m.msh"))


@pytest.mark.skipif(not Path("/root/reference/data/meshes").exists(), reason="reference tree is only present in the authoring container")
def test_c_abi_parser_on_the_reference_mesh_files():
    ref = Path("/root/reference/data/meshes")
    n = 0
    for f in sorted(ref.glob("*.msh")):
        try:
            a = Mesh.read(f)
        except GemlabTextError as ea:
            with pytest.raises(GemlabTextError) as eb:
                Mesh.read_native(f)
            assert str(eb.value) == str(ea), f.name
            continue
        assert _same(a, Mesh.read_native(f)), f.name
        n += 1
    assert n >= 10


def test_missing_file():
    with pytest.raises(GemlabTextError) as e:
        Mesh.read("/nonexistent/dir/mesh.msh")
    assert str(e.value) == "cannot open file"


REF_MESHES = Path("/root/reference/data/meshes")


@pytest.mark.skipif(not REF_MESHES.exists(), reason="reference tree is only present in the authoring container")
def test_reference_mesh_files():
    good = {"four_hex8.msh": (3, 18, 4), "two_cubes_vertical.msh": (3, 12, 2), "five_tets_within_cube.msh": (3, 8, 5)}
    for name, (ndim, npoint, ncell) in good.items():
        m = Mesh.read(REF_MESHES / name)
        assert (m.ndim, m.npoint, m.ncell) == (ndim, npoint, ncell), name
    for name in ("mixed_shapes_2d.msh", "mixed_shapes_3d.msh", "rectangle_tris_quads.msh", "column_distorted_tris_quads.msh"):
        m = Mesh.read(REF_MESHES / name)
        assert m.ncell > 0 and int(m.conn.max()) < m.npoint, name
    bad = {"bad_empty.msh": "file is empty or header is missing", "bad_missing_points.msh": "not all points have been found",
           "bad_missing_cells.msh": "not all cells have been found", "bad_wrong_cell_kind.msh": "string representation of GeoKind is incorrect",
           "bad_extra_point_data.msh": "point data contains extra values", "bad_extra_cell_data.msh": "cell data contains extra values",
           "bad_missing_marked_edges.msh": "not all marked edges have been found",
           "bad_missing_marked_faces.msh": "not all marked faces have been found"}
    for name, msg in bad.items():
        with pytest.raises(GemlabTextError) as e:
            Mesh.read(REF_MESHES / name)
        assert str(e.value) == msg, name


@pytest.mark.gpu
def test_text_mesh_drives_the_gpu_path():
    from oracle import pyoracle as po

    ctx = g.Context(0)
    for text, two_dim in ((TWO_CUBES, False), (MIXED_2D, True)):
        mesh = Mesh.from_text(text)
        dmesh = ctx.upload(mesh)
        dd = g.lin_elasticity(480.0, 1.0 / 3.0, two_dim)
        # one modulus per cell marker, as a pmsim-style assembler would configure it
        recs = np.stack([g.mandel_matrix(dd), g.mandel_matrix(2.0 * dd)])
        got = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.per_marker([1, 2], recs))
        pos = 0
        for e in range(mesh.ncell):
            sub = mesh.permuted([e])
            ref = po.mesh_integ("mat_10_bdb", sub.ndim, sub.coords, sub.kinds, sub.conn_off, sub.conn,
                                coef=recs[int(mesh.markers[e]) - 1])
            assert np.max(np.abs(got[pos:pos + ref.size] - ref)) <= 1e-12 * np.max(np.abs(ref))
            pos += ref.size
        assert pos == got.size
    ctx.close()


@pytest.mark.gpu
def test_gl_mesh_from_text_uploads_what_it_parses():
    from oracle import pyoracle as po

    ctx = g.Context(0)
    for text, two_dim in ((TWO_CUBES, False), (MIXED_2D, True)):
        mesh = Mesh.from_text(text)
        dd = g.lin_elasticity(480.0, 1.0 / 3.0, two_dim)
        ref = po.mesh_integ("mat_10_bdb", mesh.ndim, mesh.coords, mesh.kinds, mesh.conn_off, mesh.conn, coef=dd.ravel(order="F"))
        for fmt, payload in (("msh", text), ("json", mesh.to_json())):
            dm = ctx.mesh_from_text(payload, fmt)
            got = integ.mat_10_bdb(ctx, dm, None, CommonArgs(), Coef.uniform(g.mandel_matrix(dd)))
            assert np.max(np.abs(got - ref)) <= 1e-12 * np.max(np.abs(ref)), fmt
    with pytest.raises(g.GemlabError, match="not all cells have been found"):
        ctx.mesh_from_text("2 3 1 0 0\n0 0 0 0\n1 0 1 0\n2 0 0 1\n")
    ctx.close()

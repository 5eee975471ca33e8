"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, host-side queries, Block ordering
against the reference's sample meshes, output-layout and partition logic.  No compute call is made here."""
import ctypes as C
import json
import re
from pathlib import Path

import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Block, GeoClass, GeoKind, Gauss, Mesh
from gemlab_b200._lib import SYMBOLS
from oracle import pyoracle as po

ROOT = Path(__file__).resolve().parent.parent
SAMPLES = json.loads((Path(__file__).parent / "golden" / "block_samples.json").read_text())


def test_library_exports_every_symbol_declared_in_the_header():
    header = (ROOT / "include" / "gemlab_cuda.h").read_text()
    declared = set(re.findall(r"GL_API\s+[\w\s\*]+?\b(gl_\w+)\s*\(", header))
    assert len(declared) >= 40
    handle = C.CDLL(str(g.LIB_PATH))
    for name in sorted(declared):
        assert hasattr(handle, name), f"{name} is declared in include/gemlab_cuda.h but not exported"
    # the Python binding covers the whole header
    assert declared == set(SYMBOLS), declared ^ set(SYMBOLS)


def test_no_cpu_fallback_without_a_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(g.GemlabError, match="no usable sm_100 CUDA device"):
        g.Context(0)


def test_product_package_does_not_import_the_oracle():
    for f in (ROOT / "gemlab_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cpp", ".hpp", ".h") and f.is_file():
            txt = f.read_text()
            assert "pyoracle" not in txt and "gemlab_oracle" not in txt, f


def test_geokind_queries_match_the_reference_tables():
    # GeoKind::VALUES order, nnode, ndim, class (src/shapes/enums.rs:280-360, 1467-1493)
    names = ["lin2", "lin3", "lin4", "lin5", "tri3", "tri6", "tri10", "tri15", "qua4", "qua8", "qua9", "qua12", "qua16",
             "qua17", "tet4", "tet10", "tet20", "hex8", "hex20", "hex32"]
    nnode = [2, 3, 4, 5, 3, 6, 10, 15, 4, 8, 9, 12, 16, 17, 4, 10, 20, 8, 20, 32]
    for k, (name, nn) in enumerate(zip(names, nnode)):
        kind = GeoKind(k)
        assert kind.to_string() == name and GeoKind.from_str(name) == kind
        assert kind.nnode() == nn == po.kind_nnode(k)
        assert kind.ndim() == po.kind_geo_ndim(k)
        assert int(kind.geo_class()) == po.kind_class(k)
    with pytest.raises(ValueError, match="string representation of GeoKind is incorrect"):
        GeoKind.from_str("hex9")
    assert GeoClass.Hex.ndim() == 3 and GeoClass.Tri.ndim() == 2


@pytest.mark.parametrize("kind", ["lin2", "lin3", "tri3", "tri6", "qua4", "qua8", "qua9", "tet4", "tet10", "hex8", "hex20"])
def test_device_tables_match_the_oracle_shapes(kind):
    """The library tabulates N and L from family-wise formulas; the oracle restates the reference node by node."""
    k = GeoKind.from_str(kind)
    gauss = Gauss.new(k)
    ogauss = po.Gauss.new(kind)
    np.testing.assert_array_equal(gauss.data, ogauss.data)  # rules are verbatim copies
    for p in range(gauss.npoint()):
        nn, ll = k.interp_deriv(gauss.coords(p)[:3])
        np.testing.assert_allclose(nn, po.interp(kind, gauss.coords(p)[:3]), rtol=0, atol=5e-16)
        np.testing.assert_allclose(ll, po.deriv(kind, gauss.coords(p)[:3]), rtol=0, atol=2e-15)


def test_gauss_selection_and_errors():
    assert Gauss.new(GeoKind.Hex20).npoint() == 27 and Gauss.new(GeoKind.Tet10).npoint() == 14
    assert Gauss.new_sized(GeoClass.Hex, 64).npoint() == 64
    assert Gauss.new_or_sized(GeoKind.Qua8, None).npoint() == 9
    with pytest.raises(g.GemlabError, match="requested number of integration points is not available for Tet class"):
        Gauss.new_sized(GeoClass.Tet, 3)


@pytest.mark.parametrize("name,target,corners,ndiv", [
    ("block_2d_four_qua4", GeoKind.Qua4, [[0, 0], [2, 0], [2, 2], [0, 2]], [2, 2]),
    ("block_2d_four_qua8", GeoKind.Qua8, [[0, 0], [2, 0], [2, 2], [0, 2]], [2, 2]),
    ("block_2d_four_qua9", GeoKind.Qua9, [[0, 0], [2, 0], [2, 2], [0, 2]], [2, 2]),
    ("block_3d_eight_hex8", GeoKind.Hex8, [[0, 0, 0], [2, 0, 0], [2, 2, 0], [0, 2, 0], [0, 0, 4], [2, 0, 4], [2, 2, 4], [0, 2, 4]], [2, 2, 2]),
    ("block_3d_eight_hex20", GeoKind.Hex20, [[0, 0, 0], [2, 0, 0], [2, 2, 0], [0, 2, 0], [0, 0, 4], [2, 0, 4], [2, 2, 4], [0, 2, 4]], [2, 2, 2]),
])
def test_block_subdivide_reproduces_the_reference_samples(name, target, corners, ndiv):
    """src/mesh/block.rs:1770-1782, 1866-1878, 1894-1906, 2062-2087, 2180-2194 vs Samples::* (bit-exact ordering)."""
    sample = SAMPLES[name]
    mesh = Block.new(np.array(corners, dtype=float)).set_ndiv(ndiv).subdivide(target)
    assert mesh.ndim == sample["ndim"]
    assert mesh.ncell == len(sample["cells"]) and mesh.npoint == len(sample["coords"])
    conn = mesh.conn_matrix()
    for cell in sample["cells"]:
        assert cell["kind"] == target.to_string()
        assert cell["marker"] == int(mesh.markers[cell["id"]])
        assert conn[cell["id"]].tolist() == cell["points"], (name, cell["id"])
    np.testing.assert_allclose(mesh.coords, np.array(sample["coords"]), rtol=0, atol=1e-15)


def test_block_sizes_of_the_baseline_configs():
    m = Block.new_cube(1.0).set_ndiv([20, 20, 20]).subdivide(GeoKind.Hex8)
    assert (m.ncell, m.npoint) == (8000, 21 ** 3)
    m = Block.new_cube(1.0).set_ndiv([10, 10, 10]).subdivide(GeoKind.Hex20)
    assert (m.ncell, m.npoint) == (1000, 21 ** 3 - 10 ** 3 - 3 * 10 * 10 * 11)  # lattice minus cell and face centres
    # 200^3 Hex8 -> 8,120,601 points and 100^3 Hex20 -> 4,090,601 points, by the same formulas
    assert 201 ** 3 == 8_120_601 and 201 ** 3 - 100 ** 3 - 3 * 100 * 100 * 101 == 4_090_601


def test_mesh_helpers():
    q = Block.new_square(1.0).set_ndiv([3, 2]).subdivide(GeoKind.Qua4)
    perm = np.array([5, 0, 3, 1, 4, 2])
    pm = q.permuted(perm)
    np.testing.assert_array_equal(pm.conn_matrix(), q.conn_matrix()[perm])
    np.testing.assert_array_equal(pm.cell_coords(0), q.cell_coords(5))
    cells = [(GeoKind.Tri3, [0, 1, 5]), (GeoKind.Qua4, [1, 2, 6, 5])]
    mixed = Mesh.from_cells(2, q.coords, cells)
    assert not mixed.is_homogeneous() and mixed.conn_off.tolist() == [0, 3, 7]
    t = g.split_hex_cells_into_tet10(2, 2, 2)
    assert t.ncell == 48 and t.npoint == 5 ** 3
    # every tetrahedron is positively oriented and the tets fill the cube
    vol = 0.0
    for e in range(t.ncell):
        x = t.cell_coords(e)[:4]
        v = np.linalg.det((x[1:] - x[0]).T) / 6.0
        assert v > 0
        vol += v
    assert abs(vol - 1.0) < 1e-14
    j = g.jitter_interior_points(q, 0.01)
    assert np.array_equal(j.conn, q.conn) and not np.array_equal(j.coords, q.coords)
    on_boundary = (q.coords[:, 0] == 0) | (q.coords[:, 0] == 1) | (q.coords[:, 1] == 0) | (q.coords[:, 1] == 1)
    np.testing.assert_array_equal(j.coords[on_boundary], q.coords[on_boundary])


def test_sizes_and_partition():
    from gemlab_b200 import integ

    assert integ.op_out_size("mat_10_bdb", GeoKind.Hex8, 3) == 576
    assert integ.op_out_size("mat_10_bdb", GeoKind.Hex20, 3) == 3600
    assert integ.op_out_size("mat_01_nsn", GeoKind.Tet10, 3) == 100
    assert integ.op_out_size("vec_02_nv", GeoKind.Qua8, 2) == 16
    assert integ.op_coef_size("mat_10_bdb", 3) == 36 and integ.op_coef_size("mat_03_btb", 2) == 4
    for ncell in (0, 1, 7, 8_000_000, 10 ** 7 + 3):
        for world in (1, 2, 3, 8):
            ranges = [g.partition(ncell, world, r) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == ncell
            for (a, b), (c, d) in zip(ranges[:-1], ranges[1:]):
                assert b == c and a <= b
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    dd = g.lin_elasticity(480.0, 1 / 3, False)
    np.testing.assert_array_equal(dd, po.lin_elasticity(480.0, 1 / 3, False, False))
    np.testing.assert_array_equal(g.lin_elasticity(10_000.0, 0.2, True, True), po.lin_elasticity(10_000.0, 0.2, True, True))


def test_rust_ffi_declares_every_abi_symbol():
    """gemlab-cuda/src/ffi.rs (source only: no cargo in this image) must declare every GL_API symbol of the header."""
    import re
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    hdr = (root / "include" / "gemlab_cuda.h").read_text()
    names = set(re.findall(r"GL_API\s+[\w\s\*]+?\b(gl_\w+)\s*\(", hdr))
    ffi = set(re.findall(r"pub fn (gl_\w+)", (root / "gemlab-cuda" / "src" / "ffi.rs").read_text()))
    assert names and not (names - ffi), sorted(names - ffi)


def test_cpp_host_mirror_compiles_and_links():
    """cpp/gemlab.hpp and its two test programs build against the in-tree library (they RUN in the -m gpu suite)."""
    import subprocess
    import tempfile
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    with tempfile.TemporaryDirectory() as tmp:
        for name in ("test_batch", "test_device_paths"):
            subprocess.run(["/usr/bin/g++", "-std=c++17", "-O1", "-Wall", "-o", str(Path(tmp) / name), str(root / "tests" / "cpp" / f"{name}.cpp"),
                            "-L", str(root / "gemlab_b200"), "-lgemlab_cuda", f"-Wl,-rpath,{root / 'gemlab_b200'}"], check=True)


def test_rust_crate_uses_only_declared_ffi_items_and_matches_the_header_enums():
    """gemlab-cuda cannot be compiled here (no cargo): at least every `ffi::name` lib.rs uses must exist in ffi.rs, the op / format
    constants must equal the header's enums, and the repr(C) structs must list the header's fields in order."""
    import re
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    lib_rs = (root / "gemlab-cuda" / "src" / "lib.rs").read_text()
    ffi_rs = (root / "gemlab-cuda" / "src" / "ffi.rs").read_text()
    hdr = (root / "include" / "gemlab_cuda.h").read_text()
    declared = set(re.findall(r"pub (?:fn|const|struct) (\w+)", ffi_rs))
    used = set(re.findall(r"ffi::(\w+)", lib_rs))
    assert not (used - declared), sorted(used - declared)
    for name, val in re.findall(r"pub const (GL_OP_\w+): c_int = (\d+);", ffi_rs):
        m = re.search(rf"\b{name} = (\d+)", hdr)
        assert m and int(m.group(1)) == int(val), name
    for name, val in re.findall(r"pub const (GL_MESH_FORMAT_\w+): c_int = (\d+);", ffi_rs):
        m = re.search(rf"\b{name} = (\d+)", hdr)
        assert m and int(m.group(1)) == int(val), name
    # struct field order: gl_slab, gl_args, gl_coef, gl_out
    for st in ("gl_slab", "gl_args", "gl_coef", "gl_out", "gl_fused_item"):
        c_body = re.search(rf"typedef struct {st} \{{(.*?)\}} {st};", hdr, re.S).group(1)
        c_body = re.sub(r"/\*.*?\*/", "", c_body, flags=re.S)
        c_fields = [re.sub(r"\[.*", "", f.strip().split()[-1]).lstrip("*") for f in c_body.split(";") if f.strip()]
        c_fields = [f for decl in c_fields for f in decl.split(",")]
        r_body = re.search(rf"pub struct {st} \{{(.*?)\n\}}", ffi_rs, re.S).group(1)
        r_fields = re.findall(r"pub (\w+):", r_body)
        flat_c = []
        for f in c_body.split(";"):
            f = f.strip()
            if not f:
                continue
            names = f.split(None, 1)[1] if " " in f else f
            # "const double *data" / "uint32_t ii0, jj0" / "double ksi[3]"
            names = re.sub(r"^(?:const\s+)?\w+\s+", "", f)
            for n in names.split(","):
                flat_c.append(re.sub(r"\[.*", "", n.strip().lstrip("*").strip()))
        assert flat_c == r_fields, (st, flat_c, r_fields)

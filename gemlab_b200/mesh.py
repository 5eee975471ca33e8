"""Host-side Mesh and Block -- mirror of gemlab::mesh::{Mesh, Point, Cell, Block} for the integration path.

Reference: `Mesh{ndim, points: Vec<Point{id, marker, coords}>, cells: Vec<Cell{id, marker, kind, points}>}`
(src/mesh/mesh.rs:27-54, 110-135).  The reference stores one heap allocation per point and per cell (array of
structures); here the same information is held as flat numpy arrays that go to the device in one copy each:

    coords   (npoint, ndim) float64     Point.coords        (point id = row)
    kinds    (ncell,)       uint8       Cell.kind           (GeoKind ordinal, GeoKind::VALUES order)
    markers  (ncell,)       int32       Cell.marker
    conn_off (ncell+1,)     uint64      offsets into conn
    conn     (...,)         uint64      Cell.points         (usize in the reference)

`Block` reproduces the cell ordering (x fastest, then y, then z) and the FIRST-TOUCH point numbering of
`Block::subdivide` (src/mesh/block.rs:661-1065; loops :749-761, dedup :782-810) with closed-form lattice indexing
instead of the reference's GridSearch hash; fixtures: Samples::block_2d_four_qua4 / qua8,
block_3d_eight_hex8 / hex20 (src/mesh/samples.rs:1407, 1466, 1901, 2043).
"""
import numpy as np

from .geo import GeoClass, GeoKind

# NODE_REFERENCE_COORDS of the tensor-product kinds (src/shapes/qua4.rs:67, qua8.rs:68, qua9.rs, hex8.rs:122, hex20.rs:123)
_QUA_REF = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1], [0, -1], [1, 0], [0, 1], [-1, 0], [0, 0]], dtype=np.int64)
_HEX_REF = np.array([[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1],
                     [0, -1, -1], [1, 0, -1], [0, 1, -1], [-1, 0, -1], [0, -1, 1], [1, 0, 1], [0, 1, 1], [-1, 0, 1],
                     [-1, -1, 0], [1, -1, 0], [1, 1, 0], [-1, 1, 0]], dtype=np.int64)


class GemlabTextError(ValueError):
    """A StrError of the MSH reader; str(e) is the reference's message (src/mesh/read_text_mesh.rs)."""


class Mesh:
    """Flat-array mesh (see module docstring)."""

    def __init__(self, ndim, coords, kinds, conn_off, conn, markers=None):
        self.ndim = int(ndim)
        self.coords = np.ascontiguousarray(coords, dtype=np.float64).reshape(-1, self.ndim)
        self.kinds = np.ascontiguousarray(kinds, dtype=np.uint8)
        self.conn_off = np.ascontiguousarray(conn_off, dtype=np.uint64)
        self.conn = np.ascontiguousarray(conn, dtype=np.uint64)
        ncell = self.kinds.shape[0]
        self.markers = (np.ones(ncell, dtype=np.int32) if markers is None
                        else np.ascontiguousarray(markers, dtype=np.int32))
        if self.conn_off.shape[0] != ncell + 1 or self.markers.shape[0] != ncell:
            raise ValueError("inconsistent mesh arrays")

    @property
    def npoint(self):
        return self.coords.shape[0]

    @property
    def ncell(self):
        return self.kinds.shape[0]

    @staticmethod
    def homogeneous(ndim, coords, kind, conn, markers=None):
        """All cells of one GeoKind; conn is (ncell, nnode)."""
        conn = np.ascontiguousarray(conn, dtype=np.uint64)
        ncell, nnode = conn.shape
        if nnode != GeoKind(kind).nnode():
            raise ValueError("connectivity width != nnode")
        kinds = np.full(ncell, int(kind), dtype=np.uint8)
        conn_off = np.arange(ncell + 1, dtype=np.uint64) * np.uint64(nnode)
        return Mesh(ndim, coords, kinds, conn_off, conn.reshape(-1), markers)

    @staticmethod
    def from_cells(ndim, coords, cells, markers=None):
        """cells: list of (GeoKind, [point ids]) -- the reference's Vec<Cell>."""
        kinds = np.array([int(k) for k, _ in cells], dtype=np.uint8)
        lens = np.array([len(p) for _, p in cells], dtype=np.uint64)
        conn_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        conn = np.concatenate([np.asarray(p, dtype=np.uint64) for _, p in cells]) if cells else np.zeros(0, np.uint64)
        return Mesh(ndim, coords, kinds, conn_off, conn, markers)

    # ---- MSH text format (README.md "MSH file format"; reader: src/mesh/read_text_mesh.rs:36-256, 312-420) --------------

    @staticmethod
    def from_text(text):
        """Parses a mesh in gemlab's MSH text format (the step BEFORE the integration path: it lets the reference's own
        data/meshes/*.msh drive the GPU path unmodified).  Same grammar and the same error strings as Mesh::read /
        Mesh::from_text: `#` comments, header `ndim npoint ncell [nmarked_edge nmarked_face]`, points `id marker x y [z]`
        with id == position, cells `id marker kind p0 p1 ..` with lower-case GeoKind keys (GeoKind::from,
        src/shapes/enums.rs:217-244), then optional marked edges `marker p1 p2` and faces `marker p1 p2 p3 [p4]`."""
        def fields(line):
            t = line.strip()
            return None if (not t or t.startswith("#")) else t.split()

        def num(tok, conv, msg):
            try:
                return conv(tok)
            except ValueError:
                raise GemlabTextError(msg) from None

        def uint(tok, msg):
            v = num(tok, int, msg)
            if v < 0 or not tok.lstrip("+").isdigit():
                raise GemlabTextError(msg)
            return v

        lines = iter(text.splitlines())
        header = None
        for line in lines:
            f = fields(line)
            if f is None:
                continue
            names = ["ndim", "npoint", "ncell", "nmarked_edge", "nmarked_face"]
            header = []
            for i, name in enumerate(names):
                if i >= len(f):
                    raise GemlabTextError(f"cannot read {name}")
                header.append(uint(f[i], f"cannot parse {name}"))
            break
        if header is None:
            raise GemlabTextError("file is empty or header is missing")
        ndim, npoint, ncell, nedge, nface = header

        coords, pmarkers = [], []
        for line in lines:
            if len(coords) == npoint:
                break
            f = fields(line)
            if f is None:
                continue
            if uint(f[0], "cannot parse point id") != len(coords):
                raise GemlabTextError("the id and index of points must equal each other")
            if len(f) < 2:
                raise GemlabTextError("cannot read point marker")
            pmarkers.append(num(f[1], int, "cannot parse point marker"))
            xyz = []
            for j, axis in enumerate("xyz"[:ndim]):
                if len(f) < 3 + j:
                    raise GemlabTextError(f"cannot read point {axis} coordinate")
                xyz.append(num(f[2 + j], float, f"cannot parse point {axis} coordinate"))
            if len(f) > 2 + ndim:
                raise GemlabTextError("point data contains extra values")
            coords.append(xyz)
            if len(coords) == npoint:
                break
        if len(coords) != npoint:
            raise GemlabTextError("not all points have been found")

        cells, cmarkers = [], []
        for line in lines:
            f = fields(line)
            if f is None:
                continue
            if uint(f[0], "cannot parse cell id") != len(cells):
                raise GemlabTextError("the id and index of cells must equal each other")
            if len(f) < 2:
                raise GemlabTextError("cannot read cell marker")
            cmarkers.append(num(f[1], int, "cannot parse cell marker"))
            if len(f) < 3:
                raise GemlabTextError("cannot read cell kind")
            try:
                kind = GeoKind.from_str(f[2])
            except Exception:
                raise GemlabTextError("string representation of GeoKind is incorrect") from None
            nn = kind.nnode()
            if len(f) < 3 + nn:
                raise GemlabTextError("cannot read cell point id")
            pts = [uint(t, "cannot parse cell point id") for t in f[3:3 + nn]]
            if len(f) > 3 + nn:
                raise GemlabTextError("cell data contains extra values")
            cells.append((kind, pts))
            if len(cells) == ncell:
                break
        if len(cells) != ncell:
            raise GemlabTextError("not all cells have been found")

        def marked(count, what, npts_min, npts_max):
            out = []
            if count == 0:
                return out
            for line in lines:
                f = fields(line)
                if f is None:
                    continue
                marker = num(f[0], int, f"cannot parse marked {what}")
                pts = []
                for j in range(npts_max):
                    if len(f) < 2 + j:
                        if j < npts_min:
                            raise GemlabTextError(f"cannot read p{j + 1} of marked {what}")
                        break
                    pts.append(uint(f[1 + j], f"cannot parse p{j + 1} of marked {what}"))
                if len(f) > 1 + npts_max:
                    raise GemlabTextError(f"marked {what} data contains extra values")
                out.append((marker, *pts))
                if len(out) == count:
                    break
            if len(out) != count:
                raise GemlabTextError(f"not all marked {what}s have been found")
            return out

        edges = marked(nedge, "edge", 2, 2)
        faces = marked(nface, "face", 3, 4)
        mesh = Mesh.from_cells(ndim, np.array(coords, dtype=np.float64).reshape(npoint, ndim), cells, cmarkers)
        mesh.point_markers = np.array(pmarkers, dtype=np.int32)
        mesh.marked_edges = edges
        mesh.marked_faces = faces
        return mesh

    # ---- the same formats through the C ABI (gl_host_mesh_parse: the parser a Rust / C++ host links) ---------------------------

    @staticmethod
    def _from_host_handle(h):
        import ctypes as C

        from ._lib import lib

        L = lib()
        try:
            ndim, npoint, ncell = L.gl_host_mesh_ndim(h), L.gl_host_mesh_npoint(h), L.gl_host_mesh_ncell(h)
            ne, nf = L.gl_host_mesh_nmarked_edge(h), L.gl_host_mesh_nmarked_face(h)

            def arr(ptr, n, dtype):
                return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dtype, copy=True) if n else np.zeros(0, dtype)

            conn_off = arr(L.gl_host_mesh_conn_off(h), ncell + 1, np.uint64)
            mesh = Mesh(ndim, arr(L.gl_host_mesh_coords(h), npoint * ndim, np.float64).reshape(npoint, ndim),
                        arr(L.gl_host_mesh_kinds(h), ncell, np.uint8), conn_off,
                        arr(L.gl_host_mesh_conn(h), int(conn_off[-1]) if ncell else 0, np.uint64), arr(L.gl_host_mesh_markers(h), ncell, np.int32))
            mesh.point_markers = arr(L.gl_host_mesh_point_markers(h), npoint, np.int32)
            ed = arr(L.gl_host_mesh_marked_edges(h), ne * 3, np.int64).reshape(ne, 3)
            fa = arr(L.gl_host_mesh_marked_faces(h), nf * 5, np.int64).reshape(nf, 5)
            mesh.marked_edges = [tuple(int(v) for v in r) for r in ed]
            mesh.marked_faces = [tuple(int(v) for v in r if v >= 0 or v == r[0]) if r[4] >= 0 else (int(r[0]), int(r[1]), int(r[2]), int(r[3]))
                                 for r in fa]
            return mesh
        finally:
            L.gl_host_mesh_free(h)

    @staticmethod
    def from_text_native(text, fmt="msh"):
        """Parses MSH text (fmt="msh") or the JSON image of Mesh (fmt="json", Mesh::read_json, src/mesh/mesh.rs:245) with the C++
        parser behind the C ABI (gl_host_mesh_parse) -- the reference's grammar and error strings."""
        import ctypes as C

        from ._lib import lib

        data = text.encode("utf-8") if isinstance(text, str) else bytes(text)
        h = C.c_void_p()
        st = lib().gl_host_mesh_parse(data, len(data), {"msh": 0, "json": 1}[fmt], C.byref(h))
        if st:
            raise GemlabTextError(lib().gl_mesh_text_last_error().decode("utf-8"))
        return Mesh._from_host_handle(h)

    @staticmethod
    def read_native(path, fmt=None):
        """Mesh::read / Mesh::read_json through the C ABI (gl_host_mesh_read); fmt defaults to the file extension."""
        import ctypes as C

        from ._lib import lib

        fmt = fmt or ("json" if str(path).lower().endswith(".json") else "msh")
        h = C.c_void_p()
        st = lib().gl_host_mesh_read(str(path).encode("utf-8"), {"msh": 0, "json": 1}[fmt], C.byref(h))
        if st:
            raise GemlabTextError(lib().gl_mesh_text_last_error().decode("utf-8"))
        return Mesh._from_host_handle(h)

    def to_json(self):
        """serde_json's image of the reference's Mesh (Mesh::write_json, src/mesh/mesh.rs:263): GeoKind as its variant name, the
        missing fourth point of a triangular marked face as usize::MAX."""
        import json

        pm = getattr(self, "point_markers", np.zeros(self.npoint, dtype=np.int32))
        variant = lambda k: (lambda s: s[0].upper() + s[1:])(GeoKind(int(k)).to_string())
        return json.dumps({
            "ndim": self.ndim,
            "points": [{"id": i, "marker": int(pm[i]), "coords": [float(v) for v in self.coords[i]]} for i in range(self.npoint)],
            "cells": [{"id": e, "marker": int(self.markers[e]), "kind": variant(self.kinds[e]), "points": [int(p) for p in self.cell_points(e)]}
                      for e in range(self.ncell)],
            "marked_edges": [[int(v) for v in ed] for ed in getattr(self, "marked_edges", [])],
            "marked_faces": [[int(v) for v in fa] + [18446744073709551615] * (5 - len(fa)) for fa in getattr(self, "marked_faces", [])],
        })

    @staticmethod
    def read(path):
        """Mesh::read (src/mesh/read_text_mesh.rs:312)."""
        try:
            with open(path, "r") as fh:
                text = fh.read()
        except OSError:
            raise GemlabTextError("cannot open file") from None
        return Mesh.from_text(text)

    def to_text(self):
        """The inverse of from_text (Mesh::write, src/mesh/write_text_mesh.rs): ids equal positions, lower-case kinds."""
        pm = getattr(self, "point_markers", np.zeros(self.npoint, dtype=np.int32))
        edges = getattr(self, "marked_edges", [])
        faces = getattr(self, "marked_faces", [])
        out = ["# header", "# ndim npoint ncell nmarked_edge nmarked_face",
               f"{self.ndim} {self.npoint} {self.ncell} {len(edges)} {len(faces)}", "", "# points", "# id marker x y {z}"]
        for i in range(self.npoint):
            out.append(f"{i} {int(pm[i])} " + " ".join(repr(float(v)) for v in self.coords[i]))
        out += ["", "# cells", "# id marker kind points"]
        for e in range(self.ncell):
            out.append(f"{e} {int(self.markers[e])} {GeoKind(int(self.kinds[e])).to_string()} "
                       + " ".join(str(int(p)) for p in self.cell_points(e)))
        if edges:
            out += ["", "# marked edges", "# marker p1 p2"] + [" ".join(str(int(v)) for v in ed) for ed in edges]
        if faces:
            out += ["", "# marked faces", "# marker p1 p2 p3 {p4}"] + [" ".join(str(int(v)) for v in fa) for fa in faces]
        return "\n".join(out) + "\n"

    def is_homogeneous(self):
        return self.ncell == 0 or bool(np.all(self.kinds == self.kinds[0]))

    def cell_points(self, cell_id):
        return self.conn[int(self.conn_off[cell_id]): int(self.conn_off[cell_id + 1])]

    def cell_coords(self, cell_id):
        """What Mesh::set_pad gathers into pad.xxt (src/mesh/mesh.rs:448-456), as (nnode, ndim)."""
        return self.coords[self.cell_points(cell_id).astype(np.int64)]

    def conn_matrix(self):
        """(ncell, nnode) connectivity of a homogeneous mesh."""
        if not self.is_homogeneous():
            raise ValueError("mesh is not homogeneous")
        nnode = GeoKind(int(self.kinds[0])).nnode() if self.ncell else 0
        return self.conn.reshape(self.ncell, nnode)

    def permuted(self, perm):
        """A new mesh whose cell i is this mesh's cell perm[i] (used to build irregular-gather test meshes)."""
        perm = np.asarray(perm, dtype=np.int64)
        lens = (self.conn_off[1:] - self.conn_off[:-1]).astype(np.int64)[perm]
        conn_off = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64)
        starts = self.conn_off[:-1].astype(np.int64)[perm]
        idx = np.repeat(starts - conn_off[:-1].astype(np.int64), lens) + np.arange(int(conn_off[-1]))
        return Mesh(self.ndim, self.coords, self.kinds[perm], conn_off, self.conn[idx], self.markers[perm])


def _serendipity_interp(ref, ksi):
    """N (npoint, nnode) of Qua8 / Hex20 at ksi (npoint, ndim) -- the Block's own geometry interpolation
    (block_pad.calc_coords, src/mesh/block.rs:794)."""
    ndim = ref.shape[1]
    a = ref.astype(np.float64)
    f = [1.0 + ksi[:, [d]] * a[None, :, d] for d in range(ndim)]  # (npoint, nnode)
    nn = np.empty((ksi.shape[0], ref.shape[0]))
    ncorner = 4 if ndim == 2 else 8
    g = sum(ksi[:, [d]] * a[None, :, d] for d in range(ndim)) - (ndim - 1)
    prod = np.ones_like(f[0])
    for d in range(ndim):
        prod = prod * f[d]
    nn[:, :ncorner] = (prod * g)[:, :ncorner] / (4.0 if ndim == 2 else 8.0)
    for m in range(ncorner, ref.shape[0]):
        val = np.ones(ksi.shape[0])
        for d in range(ndim):
            val = val * ((1.0 - ksi[:, d] ** 2) if ref[m, d] == 0 else f[d][:, m])
        nn[:, m] = val / (2.0 if ndim == 2 else 4.0)
    return nn


class Block:
    """gemlab::mesh::Block (src/mesh/block.rs:207-408, 661-1065): a Qua8 (2D) / Hex20 (3D) block subdivided into
    nx x ny (x nz) Qua / Hex cells.  Constraints, ring transform and edge/face markers are out of scope."""

    def __init__(self, coords):
        coords = np.asarray(coords, dtype=np.float64)
        ndim = coords.shape[1]
        if ndim not in (2, 3):
            raise ValueError("ndim must be 2 or 3")
        ref = (_QUA_REF[:8] if ndim == 2 else _HEX_REF).astype(np.float64)
        ncorner, nfull = (4, 8) if ndim == 2 else (8, 20)
        self._corners = coords.copy() if coords.shape[0] == ncorner else None
        if coords.shape[0] == ncorner:
            # Block::new accepts corners only and generates the mid nodes on straight edges (block.rs:225-262)
            lin = np.ones((nfull, ncorner))
            for d in range(ndim):
                lin = lin * (1.0 + ref[:, [d]] * ref[None, :ncorner, d])
            coords = (lin / (4.0 if ndim == 2 else 8.0)) @ coords
        elif coords.shape[0] != nfull:
            raise ValueError("in 2D, the number of points must be 4 or 8; in 3D, 8 or 20")
        self.ndim = ndim
        self.coords = coords
        self.ndiv = [2] * ndim  # Block::new default
        self.cell_marker = 1

    @staticmethod
    def new(coords):
        return Block(coords)

    @staticmethod
    def new_square(length):
        return Block([[0.0, 0.0], [length, 0.0], [length, length], [0.0, length]])

    @staticmethod
    def new_cube(length):
        l = length
        return Block([[0, 0, 0], [l, 0, 0], [l, l, 0], [0, l, 0], [0, 0, l], [l, 0, l], [l, l, l], [0, l, l]])

    def set_ndiv(self, ndiv):
        if len(ndiv) != self.ndim or any(n < 1 for n in ndiv):
            raise ValueError("ndiv.len() must be equal to ndim and ndiv must be >= 1")
        self.ndiv = [int(n) for n in ndiv]
        return self

    def subdivide(self, target):
        target = GeoKind(target)
        if self.ndim == 2 and target.geo_class() != GeoClass.Qua:
            raise ValueError("in 2D, the GeoClass of target must be Qua")
        if self.ndim == 3 and target.geo_class() != GeoClass.Hex:
            raise ValueError("in 3D, the GeoClass of target must be Hex")
        if target not in (GeoKind.Qua4, GeoKind.Qua8, GeoKind.Qua9, GeoKind.Hex8, GeoKind.Hex20):
            raise ValueError("target kind is not supported by the lattice generator")
        ndim = self.ndim
        nnode = target.nnode()
        ref = (_QUA_REF if ndim == 2 else _HEX_REF)[:nnode]
        f = 1 if target in (GeoKind.Qua4, GeoKind.Hex8) else 2  # lattice steps per cell
        off = (ref + 1) * f // 2                                 # (nnode, ndim) lattice offsets in {0..f}
        n = self.ndiv + ([1] if ndim == 2 else [])
        nx, ny, nz = n
        lat = [f * nx + 1, f * ny + 1, (f * nz + 1) if ndim == 3 else 1]
        ncell = nx * ny * nz
        cid = np.arange(ncell, dtype=np.int64)
        ci, cj, ck = cid % nx, (cid // nx) % ny, cid // (nx * ny)  # x fastest (block.rs:749-761)

        def lattice_index(m):
            a = f * ci + off[m, 0]
            b = f * cj + off[m, 1]
            c = (f * ck + off[m, 2]) if ndim == 3 else 0
            return a + lat[0] * (b + lat[1] * c)

        # first touch: smallest (cell id, local node) that reaches each lattice point (block.rs:770-812)
        big = np.iinfo(np.int64).max
        first = np.full(lat[0] * lat[1] * lat[2], big, dtype=np.int64)
        lat_of = []
        for m in range(nnode):
            li = lattice_index(m)
            lat_of.append(li)
            first[li] = np.minimum(first[li], cid * nnode + m)
        used = np.nonzero(first != big)[0]
        order = np.argsort(first[used], kind="stable")
        point_of_lattice = np.full(first.shape[0], -1, dtype=np.int64)
        point_of_lattice[used[order]] = np.arange(used.shape[0])
        conn = np.stack([point_of_lattice[li] for li in lat_of], axis=1)

        # coordinates: block interpolation at ksi = lattice position mapped to [-1, 1]
        lp = used[order]
        a = lp % lat[0]
        b = (lp // lat[0]) % lat[1]
        c = lp // (lat[0] * lat[1])
        ksi = np.empty((lp.shape[0], ndim))
        ksi[:, 0] = -1.0 + 2.0 * a / (f * nx)
        ksi[:, 1] = -1.0 + 2.0 * b / (f * ny)
        if ndim == 3:
            ksi[:, 2] = -1.0 + 2.0 * c / (f * nz)
        bref = _QUA_REF[:8] if ndim == 2 else _HEX_REF
        coords = np.empty((lp.shape[0], ndim))
        step = 1 << 20
        for s in range(0, lp.shape[0], step):
            k = ksi[s:s + step]
            if self._corners is not None:
                # straight-edged block: the Qua8/Hex20 serendipity map reduces to the bi/tri-linear map of the corners
                ncorner = self._corners.shape[0]
                lin = np.full((k.shape[0], ncorner), 0.25 if ndim == 2 else 0.125)
                for d in range(ndim):
                    lin *= 1.0 + k[:, [d]] * bref[None, :ncorner, d]
                coords[s:s + step] = lin @ self._corners
            else:
                coords[s:s + step] = _serendipity_interp(bref, k) @ self.coords
        markers = np.full(ncell, self.cell_marker, dtype=np.int32)
        return Mesh.homogeneous(ndim, coords, target, conn.astype(np.uint64), markers)


def split_hex_cells_into_tet10(nx, ny, nz, length=1.0, seed=None):
    """Config 4 of BASELINE.json: an "unstructured-style" Tet10 mesh -- a cube lattice whose cells are each split into 6
    tetrahedra (Kuhn split), mid-edge nodes added in the Tet10 local order (src/shapes/tet10.rs:150-161: edges 0-1, 1-2,
    2-0, 0-3, 1-3, 2-3), and the cell order randomly permuted (seed) so the coordinate gather is irregular."""
    f = 2
    lat = [f * nx + 1, f * ny + 1, f * nz + 1]
    cid = np.arange(nx * ny * nz, dtype=np.int64)
    ci, cj, ck = cid % nx, (cid // nx) % ny, cid // (nx * ny)
    # Kuhn triangulation: the 6 permutations of the axes; every tet runs corner 000 -> +e_a -> +e_b -> 111
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    cells = []
    for perm in perms:
        v = [np.zeros(3, dtype=np.int64)]
        for ax in perm:
            nxt = v[-1].copy()
            nxt[ax] = 1
            v.append(nxt)
        v = np.array(v)  # 4 corner offsets in {0,1}^3
        # orientation: make det J > 0 (swap two vertices for odd permutations)
        d = np.linalg.det((v[1:] - v[0]).T.astype(float))
        if d < 0:
            v[[1, 2]] = v[[2, 1]]
        edges = [(0, 1), (1, 2), (2, 0), (0, 3), (1, 3), (2, 3)]
        offs = [2 * v[m] for m in range(4)] + [v[a] + v[b] for a, b in edges]
        cols = []
        for o in offs:
            cols.append((f * ci + o[0]) + lat[0] * ((f * cj + o[1]) + lat[1] * (f * ck + o[2])))
        cells.append(np.stack(cols, axis=1))
    lat_conn = np.stack(cells, axis=1).reshape(-1, 10)  # cell-major: 6 tets per hex
    used, inv = np.unique(lat_conn.reshape(-1), return_inverse=True)
    conn = inv.reshape(-1, 10)
    a, b, c = used % lat[0], (used // lat[0]) % lat[1], used // (lat[0] * lat[1])
    coords = np.stack([a / (f * nx), b / (f * ny), c / (f * nz)], axis=1) * length
    if seed is not None:
        rng = np.random.default_rng(seed)
        conn = conn[rng.permutation(conn.shape[0])]
    return Mesh.homogeneous(3, coords, GeoKind.Tet10, conn.astype(np.uint64))


def jitter_interior_points(mesh, amplitude, seed=12345):
    """Distorted variant of a Block mesh: moves every point that is not on the bounding box by
    U(-amplitude, amplitude) per coordinate; returns a new Mesh sharing the connectivity."""
    rng = np.random.default_rng(seed)
    lo, hi = mesh.coords.min(axis=0), mesh.coords.max(axis=0)
    tol = 1e-12 * max(1.0, float(np.max(hi - lo)))
    interior = np.all((mesh.coords > lo + tol) & (mesh.coords < hi - tol), axis=1)
    coords = mesh.coords.copy()
    coords[interior] += rng.uniform(-amplitude, amplitude, size=(int(interior.sum()), mesh.ndim))
    return Mesh(mesh.ndim, coords, mesh.kinds, mesh.conn_off, mesh.conn, mesh.markers)

"""Which kernel ran?  The dispatch (gl_api.cu: launch_range) tries the tuned kernels in turn and ends at the generic one, so
a tuned kernel that starts declining a configuration would leave every parity test green at a fraction of the speed.
These tests pin the kernel FAMILY per configuration through gl_ctx_last_kernel, and exercise the FP64 tensor-core
stiffness kernel of the quadratic solids (gl_kernels_syrk.cu) on everything its launcher accepts."""
import numpy as np
import pytest

import gemlab_b200 as g
from gemlab_b200 import Block, Coef, CommonArgs, GeoKind, Mesh, integ
from oracle import pyoracle as po

from test_gpu_parity import make_mesh, oracle_integ, random_coef, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def ctx():
    c = g.Context(0)
    yield c
    c.close()


def iso(ndim):
    return g.mandel_matrix(g.lin_elasticity(480.0, 1.0 / 3.0, ndim == 2))


# (kind, op, CommonArgs kwargs, expected family prefix)
FAMILIES = [
    ("hex8", "mat_10_bdb", {}, "hex8_bdb<pat=2"),
    ("hex8", "mat_10_bdb", {"ngauss": 27}, "bdb<nn=8,sd=3"),
    ("hex20", "mat_10_bdb", {}, "syrk_dmma<nn=20,pat=2>"),
    ("hex20", "mat_10_bdb", {"ngauss": 8}, "syrk_dmma<nn=20,pat=2>"),
    ("hex20", "mat_10_bdb", {"ngauss": 64}, "bdb<nn=20,sd=3"),
    ("hex20", "mat_10_bdb", {"clear": False}, "syrk_dmma<nn=20,pat=2,gen>"),
    ("hex20", "mat_10_bdb", {"ii0": 1, "jj0": 2, "nrow": 64, "ncol": 63}, "syrk_dmma<nn=20,pat=2,gen>"),
    ("tet10", "mat_10_bdb", {}, "syrk_dmma<nn=10,pat=2>"),
    ("tet10", "mat_10_bdb", {"ngauss": 24}, "bdb<nn=10,sd=3"),
    ("tet10", "mat_10_bdb", {"clear": False}, "syrk_dmma<nn=10,pat=2,gen>"),
    ("tet4", "mat_10_bdb", {}, "bdb<nn=4,sd=3"),
    ("qua4", "mat_10_bdb", {}, "small2d<nn=4,axi=0"),
    ("tri3", "mat_10_bdb", {}, "small2d<nn=3,axi=0"),
    ("qua8", "mat_10_bdb", {}, "qua8_dmma<axi=0,pat=2>"),
    ("qua8", "mat_10_bdb", {"axisymmetric": True}, "qua8_dmma<axi=1,pat=2>"),
    ("qua8", "mat_10_bdb", {"ngauss": 16}, "bdb<nn=8,sd=2,axi=0"),
    ("qua8", "mat_10_bdb", {"axisymmetric": True, "clear": False}, "qua8_dmma<axi=1,pat=2,gen>"),
    ("tri6", "mat_10_bdb", {"ii0": 1, "jj0": 2, "nrow": 15, "ncol": 16}, "tri6_dmma<axi=0,pat=2,gen>"),
    ("qua8", "mat_10_bdb", {"ngauss": 16, "clear": False}, "bdb<nn=8,sd=2,axi=0,pat=2,gen>"),
    ("qua9", "mat_10_bdb", {}, "bdb<nn=9,sd=2"),
    ("tri6", "mat_10_bdb", {}, "tri6_dmma<axi=0,pat=2>"),
    ("tri6", "mat_10_bdb", {"axisymmetric": True}, "tri6_dmma<axi=1,pat=2>"),
    ("tri6", "mat_10_bdb", {"ngauss": 12}, "bdb<nn=6,sd=2"),
    ("tet10", "mat_01_nsn", {}, "mass_dmma<nn=10,sd=3"),
    ("tet10", "vec_01_ns", {}, "mass_dmma<nn=10,sd=3"),
    ("hex8", "mat_01_nsn", {}, "mass_dmma<nn=8,sd=3"),
    ("qua8", "mat_03_btb", {}, "scalar_tpc<nn=8,sd=2"),
    ("tri3", "vec_03_bv", {"axisymmetric": True}, "scalar_tpc<nn=3,sd=2"),
    ("hex8", "vec_04_bt", {}, "vec_tpc<nn=8,sd=3"),
    ("qua4", "vec_02_nv", {}, "vec_tpc<nn=4,sd=2"),
    ("hex8", "mat_08_ntn", {}, "generic<sd=3,gd=3>"),
]


@pytest.mark.parametrize("kind,op,kw,family", FAMILIES, ids=[f"{k}-{o}-{'-'.join(map(str, a)) or 'default'}" for k, o, a, _ in FAMILIES])
def test_kernel_family(ctx, kind, op, kw, family):
    mesh = make_mesh(kind, n=3)
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(3)
    coef = iso(mesh.ndim) if op == "mat_10_bdb" else random_coef(op, mesh.ndim, 1, rng)[0]
    args = CommonArgs(**kw)
    total = integ.out_total(dmesh, op, None, args)
    out = np.zeros(total)
    getattr(integ, op)(ctx, dmesh, None, args, Coef.uniform(coef), out=out)
    assert ctx.last_kernel().startswith(family), f"{kind} {op} {kw}: ran {ctx.last_kernel()!r}, expected {family!r}"


def bulged_hex20(n=(3, 2, 4), bulge=0.05, seed=5):
    """SURVEY.md 8(d) config 3 variant: Hex20 Block with the mid-edge nodes pushed off their edges by `bulge` x cell size
    (curved edges: J varies inside every cell) and the corner nodes jittered."""
    m = Block.new_cube(1.0).set_ndiv(list(n)).subdivide(GeoKind.Hex20)
    m = g.jitter_interior_points(m, 0.03 / max(n), seed)
    rng = np.random.default_rng(seed)
    conn = m.conn_matrix()
    mid = np.unique(conn[:, 8:])
    h = 1.0 / max(n)
    m.coords[mid.astype(np.int64)] += bulge * h * rng.uniform(-1.0, 1.0, (mid.size, 3))
    return m


@pytest.mark.parametrize("rule", [6, 8, 14, 27])
@pytest.mark.parametrize("structure", ["isotropic", "symmetric", "general"])
def test_hex20_tensor_core_stiffness(ctx, rule, structure):
    """mat_10_bdb on bulged Hex20 cells through the DMMA kernel vs the oracle (src/integ/mat_10_bdb.rs:179-208), every rule it
    accepts, the three modulus patterns; cell counts that do not fill the 7 cell slots of a CTA."""
    mesh = bulged_hex20()
    rng = np.random.default_rng(31)
    dd = iso(3) if structure == "isotropic" else random_coef("mat_10_bdb", 3, 1, rng, spd=(structure == "symmetric"))[0]
    dmesh = ctx.upload(mesh)
    ref = oracle_integ("mat_10_bdb", mesh, dd, ngauss=rule)
    for b, e in [(0, mesh.ncell), (0, 1), (3, 9), (5, 19), (1, 15)]:
        got = integ.mat_10_bdb(ctx, dmesh, (b, e), CommonArgs(ngauss=rule), Coef.uniform(dd))
        assert ctx.last_kernel() == f"syrk_dmma<nn=20,pat={ {'isotropic': 2, 'symmetric': 1, 'general': 0}[structure] }>"
        assert rel_err(got, ref[b * 3600:e * 3600], np.arange(e - b + 1) * 3600) <= TOL, f"rule {rule} range {b}:{e}"


@pytest.mark.parametrize("axisymmetric", [False, True])
@pytest.mark.parametrize("kind,rule", [("qua8", 1), ("qua8", 4), ("qua8", 9), ("tri6", 1), ("tri6", 3), ("tri6", 4), ("tri6", 6), ("tri6", 7)])
@pytest.mark.parametrize("structure", ["isotropic", "symmetric", "general"])
def test_qua8_tensor_core_stiffness(ctx, axisymmetric, kind, rule, structure):
    """mat_10_bdb on jittered Qua8 / Tri6 cells through the DMMA kernel of gl_kernels_qua8.cu vs the oracle
    (src/integ/mat_10_bdb.rs:179-233, plane and axisymmetric), every rule it accepts (four Gauss points per k-step, a single leftover
    point as a rank-1 update, other remainders as a padded step), isotropic / symmetric / general moduli; ranges that do not fill the
    three cells of a warp, and more triples than one CTA holds warps."""
    mesh = make_mesh(kind, n=7 if kind == "qua8" else 5, seed=11)  # 56 Qua8 / 50 Tri6 cells
    rng = np.random.default_rng(37)
    dd = iso(2) if structure == "isotropic" else random_coef("mat_10_bdb", 2, 1, rng, spd=(structure == "symmetric"))[0]
    dmesh = ctx.upload(mesh)
    bs = 256 if kind == "qua8" else 144
    ref = oracle_integ("mat_10_bdb", mesh, dd, ngauss=rule, axisymmetric=axisymmetric)
    for b, e in [(0, mesh.ncell), (0, 1), (3, 5), (5, 19), (1, mesh.ncell), (2, 27)]:
        got = integ.mat_10_bdb(ctx, dmesh, (b, e), CommonArgs(ngauss=rule, axisymmetric=axisymmetric), Coef.uniform(dd))
        assert ctx.last_kernel() == f"{kind}_dmma<axi={int(axisymmetric)},pat={2 if structure == 'isotropic' else 0}>"
        assert rel_err(got, ref[b * bs:e * bs], np.arange(e - b + 1) * bs) <= TOL, f"rule {rule} range {b}:{e}"


@pytest.mark.parametrize("axisymmetric", [False, True])
@pytest.mark.parametrize("kind", ["qua8", "tri6"])
def test_qua8_modulus_records_on_the_tensor_core_kernel(ctx, axisymmetric, kind):
    """One modulus per marker / per cell (multi-material meshes) on the Qua8 DMMA kernel: records fetched one triple ahead and staged
    per cell; ragged ranges, host- and device-resident records (the latter through the structure probe: both pattern variants are
    enqueued, one runs)."""
    import torch
    mesh = make_mesh(kind, n=6 if kind == "qua8" else 4, seed=8)
    bs = 256 if kind == "qua8" else 144
    rng = np.random.default_rng(29)
    mesh.markers[:] = rng.integers(1, 4, mesh.ncell)
    dmesh = ctx.upload(mesh)
    args = CommonArgs(axisymmetric=axisymmetric)
    for structure in ("isotropic", "general"):
        if structure == "isotropic":
            recs = np.stack([g.mandel_matrix(g.lin_elasticity(100.0 * (k + 1), 0.1 * (k + 1), True)) for k in range(3)])
        else:
            recs = random_coef("mat_10_bdb", 2, 3, rng, spd=False)
        per_cell = np.ascontiguousarray(recs[mesh.markers - 1])
        ref = oracle_integ("mat_10_bdb", mesh, per_cell, cell_stride=16, axisymmetric=axisymmetric)
        pat = 2 if structure == "isotropic" else 0
        for b, e in [(0, mesh.ncell), (1, 2), (4, 21), (7, mesh.ncell)]:
            sizes = np.arange(e - b + 1) * bs
            got = integ.mat_10_bdb(ctx, dmesh, (b, e), args, Coef.per_marker([1, 2, 3], recs))
            assert ctx.last_kernel() == f"{kind}_dmma<axi={int(axisymmetric)},pat={pat},rec>"
            assert rel_err(got, ref[b * bs:e * bs], sizes) <= TOL, f"per_marker {structure} {b}:{e}"
            got = integ.mat_10_bdb(ctx, dmesh, (b, e), args, Coef.per_cell(per_cell[b:e]))
            assert ctx.last_kernel() == f"{kind}_dmma<axi={int(axisymmetric)},pat={pat},rec>"
            assert rel_err(got, ref[b * bs:e * bs], sizes) <= TOL, f"per_cell {structure} {b}:{e}"
            out = torch.full(((e - b) * bs + 8,), -1.5, dtype=torch.float64, device="cuda")
            integ.mat_10_bdb(ctx, dmesh, (b, e), args, Coef.per_cell(torch.from_numpy(per_cell[b:e].copy()).cuda()), out=out[4:-4])
            assert ctx.last_kernel().startswith(f"{kind}_dmma<axi={int(axisymmetric)},pat=") and ctx.last_kernel().endswith(",rec>")
            res = out.cpu().numpy()
            assert np.all(res[:4] == -1.5) and np.all(res[-4:] == -1.5)
            assert rel_err(res[4:-4], ref[b * bs:e * bs], sizes) <= TOL, f"per_cell device {structure} {b}:{e}"


def test_qua8_tensor_core_on_mixed_mesh_and_guard_band(ctx):
    """The Qua8 stripes of a mixed Tri3 / Qua8 mesh (BASELINE.json configs[4]) leave as runs of consecutive blocks; nothing
    outside the integrated range is written; zero determinants are reported with the smallest cell id."""
    from test_gpu_parity import mixed_mesh_2d
    mesh = mixed_mesh_2d(n=7)
    dmesh = ctx.upload(mesh)
    dd = iso(2)
    args = CommonArgs(axisymmetric=True)
    off = integ.out_offsets(dmesh, "mat_10_bdb", None, args)
    ref = oracle_integ("mat_10_bdb", mesh, dd, axisymmetric=True)
    for b, e in [(0, mesh.ncell), (2, mesh.ncell - 3), (8, 30)]:
        import torch
        n = int(off[e] - off[b])
        buf = torch.full((n + 64,), 7.25, dtype=torch.float64, device="cuda")
        integ.mat_10_bdb(ctx, dmesh, (b, e), args, Coef.uniform(dd), out=buf[32:32 + n])
        got = buf.cpu().numpy()
        assert np.all(got[:32] == 7.25) and np.all(got[32 + n:] == 7.25)
        assert rel_err(got[32:32 + n], ref[int(off[b]):int(off[e])], off[b:e + 1] - off[b]) <= TOL
    # homogeneous Qua8 mesh with a collapsed cell: the reference's error string and the smallest failing cell
    q8 = make_mesh("qua8", n=4)
    conn = q8.conn_matrix()
    bad = 7
    q8.coords[conn[bad].astype(np.int64)] = q8.coords[int(conn[bad][0])]
    dq = ctx.upload(q8)
    with pytest.raises(g.GemlabError, match="cannot compute inverse due to zero determinant") as ei:
        integ.mat_10_bdb(ctx, dq, (2, q8.ncell), CommonArgs(), Coef.uniform(dd))
    assert ctx.last_kernel().startswith("qua8_dmma")
    assert ei.value.cell == bad


@pytest.mark.parametrize("axisymmetric", [False, True])
@pytest.mark.parametrize("which", ["both", "mat_03_btb", "vec_03_bv"])
@pytest.mark.parametrize("layout", ["mixed", "qua8", "tri6", "tri3", "qua4"])
def test_stiffness_conductivity_flux_from_one_pass(ctx, axisymmetric, which, layout):
    """gl_integ_fused with mat_10_bdb + mat_03_btb and / or vec_03_bv (BASELINE.json configs[4]): the Qua8 cells take the scalar
    cases from the geometric tensors of the tensor-core stiffness kernel, the other kinds run their stiffness kernel and ONE fused
    scalar kernel; results equal the oracle's for every case (src/integ/mat_10_bdb.rs:179-233, mat_03_btb.rs:50-131,
    vec_03_bv.rs), ragged ranges included."""
    import torch
    from test_gpu_parity import mixed_mesh_2d
    mesh = mixed_mesh_2d(n=7) if layout == "mixed" else make_mesh(layout, n=5, seed=4)
    dmesh = ctx.upload(mesh)
    args = CommonArgs(axisymmetric=axisymmetric)
    coefs = {"mat_10_bdb": iso(2), "mat_03_btb": np.array([2.0, 3.5, 0.0, 0.75]), "vec_03_bv": np.array([1.0, -2.0])}
    ops = ["mat_10_bdb"] + (["mat_03_btb", "vec_03_bv"] if which == "both" else [which])
    kw = {"axisymmetric": True} if axisymmetric else {}
    for b, e in [(0, mesh.ncell), (3, mesh.ncell - 2), (1, 2)]:
        outs, offs = [], []
        for op in ops:
            off = integ.out_offsets(dmesh, op, (b, e), args)
            offs.append(off)
            outs.append(torch.full((int(off[-1]) + 16,), 3.25, dtype=torch.float64, device="cuda"))
        integ.fused(ctx, dmesh, (b, e), args, [(op, Coef.uniform(coefs[op]), o[8:-8]) for op, o in zip(ops, outs)])
        ran = ctx.last_kernel().split(" + ")
        mask = {"both": 10, "mat_03_btb": 2, "vec_03_bv": 8}[which]
        if layout in ("mixed", "qua8"):
            assert f"qua8_dmma<axi={int(axisymmetric)},pat=2,fused={mask}>" in ran, ran
            assert not any(k.startswith("scalar_tpc<nn=8") for k in ran), ran
        if layout == "tri6":
            assert ran == [f"tri6_dmma<axi={int(axisymmetric)},pat=2,fused={mask}>"], ran
        if layout == "tri3" or (layout == "mixed" and e - b > 1):
            assert f"small2d<nn=3,axi={int(axisymmetric)},pat=2>" in ran and f"scalar_tpc<nn=3,sd=2,mask={mask}>" in ran, ran
        if layout == "qua4":
            assert f"scalar_tpc<nn=4,sd=2,mask={mask}>" in ran, ran
        sub = mesh.permuted(np.arange(b, e))
        for op, o, off in zip(ops, outs, offs):
            got = o.cpu().numpy()
            assert np.all(got[:8] == 3.25) and np.all(got[-8:] == 3.25), op
            ref = oracle_integ(op, sub, coefs[op], **kw)
            assert rel_err(got[8:-8], ref, off) <= TOL, f"{op} range {b}:{e}"
    # rules the tensor-core kernel does not take (16 points): the same call still answers, case by case
    a16 = CommonArgs(axisymmetric=axisymmetric, ngauss=16) if layout == "qua8" else None
    if a16 is not None:
        res = integ.fused(ctx, dmesh, None, a16, [(op, Coef.uniform(coefs[op]), None) for op in ops])
        for op, got in zip(ops, res):
            ref = oracle_integ(op, mesh, coefs[op], ngauss=16, **kw)
            bs = integ.op_out_size(op, GeoKind.Qua8, 2)  # (layout == "qua8" here)
            assert rel_err(np.asarray(got), ref, np.arange(mesh.ncell + 1) * bs) <= TOL, op


@pytest.mark.parametrize("kind", ["hex8", "tet10"])
def test_stiffness_plus_scalar_cases_in_one_call_3d(ctx, kind):
    """gl_integ_fused with the stiffness next to scalar-dof cases on a 3D mesh: the stiffness runs on its own kernel, the scalar cases
    share one; same results as the separate calls (the fused call is defined as equivalent to them, include/gemlab_cuda.h)."""
    import torch
    mesh = make_mesh(kind, n=3)
    dmesh = ctx.upload(mesh)
    args = CommonArgs()
    coefs = {"mat_10_bdb": iso(3), "mat_03_btb": np.array([2.0, 3.5, 1.5, 0.75, -0.25, 0.5]), "vec_03_bv": np.array([1.0, -2.0, 0.5]),
             "mat_01_nsn": np.array([2.7])}
    for ops in (["mat_10_bdb", "mat_03_btb", "vec_03_bv"], ["mat_01_nsn", "mat_10_bdb"]):
        outs = [torch.empty(integ.out_total(dmesh, op, None, args), dtype=torch.float64, device="cuda") for op in ops]
        integ.fused(ctx, dmesh, None, args, [(op, Coef.uniform(coefs[op]), o) for op, o in zip(ops, outs)])
        ran = ctx.last_kernel().split(" + ")
        assert len(ran) == 2 and not any(k.startswith("generic") for k in ran), ran
        for op, o in zip(ops, outs):
            ref = oracle_integ(op, mesh, coefs[op])
            bs = integ.op_out_size(op, GeoKind.from_str(kind), 3)
            assert rel_err(o.cpu().numpy(), ref, np.arange(mesh.ncell + 1) * bs) <= TOL, op


def mixed_tri6_qua8(n=6, seed=6):
    """Stripes of Qua8 cells and of Tri6 pairs over one Qua9 lattice (a Qua8 is the Qua9's first eight nodes; the Tri6 take its centre)."""
    q9 = Block.new_square(2.0).set_ndiv([n, n]).subdivide(GeoKind.Qua9)
    q9.coords[:, 0] += 1.0
    c = q9.conn_matrix()
    cells = []
    for e in range(q9.ncell):
        if (e // n) % 2 == 0:
            cells.append((GeoKind.Qua8, c[e, :8]))
        else:
            cells.append((GeoKind.Tri6, c[e, [0, 1, 2, 4, 5, 8]]))
            cells.append((GeoKind.Tri6, c[e, [0, 2, 3, 8, 6, 7]]))
    return g.jitter_interior_points(Mesh.from_cells(2, q9.coords, cells), 0.01, seed)


@pytest.mark.parametrize("axisymmetric", [False, True])
def test_tensor_core_kernels_on_a_mixed_tri6_qua8_mesh(ctx, axisymmetric):
    """Both kinds of a mixed quadratic mesh on the DMMA kernel: blocks of 144 and 256 doubles interleave in the output, so the warps
    (three Qua8 / four Tri6 cells each) ship runs that start and end inside stripes; stiffness alone and with the scalar by-products."""
    import torch
    mesh = mixed_tri6_qua8()
    dmesh = ctx.upload(mesh)
    args = CommonArgs(axisymmetric=axisymmetric)
    kw = {"axisymmetric": True} if axisymmetric else {}
    coefs = {"mat_10_bdb": iso(2), "mat_03_btb": np.array([2.0, 3.5, 0.0, 0.75]), "vec_03_bv": np.array([1.0, -2.0])}
    for b, e in [(0, mesh.ncell), (2, mesh.ncell - 3), (5, 17), (7, 8)]:
        sub = mesh.permuted(np.arange(b, e))
        off = integ.out_offsets(dmesh, "mat_10_bdb", (b, e), args)
        buf = torch.full((int(off[-1]) + 32,), 4.5, dtype=torch.float64, device="cuda")
        integ.mat_10_bdb(ctx, dmesh, (b, e), args, Coef.uniform(coefs["mat_10_bdb"]), out=buf[16:-16])
        got = buf.cpu().numpy()
        assert np.all(got[:16] == 4.5) and np.all(got[-16:] == 4.5)
        assert rel_err(got[16:-16], oracle_integ("mat_10_bdb", sub, coefs["mat_10_bdb"], **kw), off) <= TOL, f"{b}:{e}"
        ops = ["mat_10_bdb", "mat_03_btb", "vec_03_bv"]
        offs = [integ.out_offsets(dmesh, op, (b, e), args) for op in ops]
        outs = [torch.full((int(o[-1]) + 16,), 3.25, dtype=torch.float64, device="cuda") for o in offs]
        integ.fused(ctx, dmesh, (b, e), args, [(op, Coef.uniform(coefs[op]), o[8:-8]) for op, o in zip(ops, outs)])
        ran = ctx.last_kernel().split(" + ")
        assert all(k.startswith(("qua8_dmma<", "tri6_dmma<")) and k.endswith(",fused=10>") for k in ran), ran
        for op, o, of in zip(ops, outs, offs):
            res = o.cpu().numpy()
            assert np.all(res[:8] == 3.25) and np.all(res[-8:] == 3.25), op
            assert rel_err(res[8:-8], oracle_integ(op, sub, coefs[op], **kw), of) <= TOL, f"{op} {b}:{e}"


def test_hex20_general_blocks_and_accumulation(ctx):
    """nrow / ncol / ii0 / jj0 and clear = false in the store epilogue of the tensor-core kernel (CommonArgs,
    src/integ/common_args.rs:5-43; kk.fill(0.0) clears the WHOLE block, mat_10_bdb.rs:141-143)."""
    mesh = bulged_hex20((2, 2, 3))
    dmesh = ctx.upload(mesh)
    dd = iso(3)
    base = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd)).reshape(mesh.ncell, 60, 60)
    out = np.full(base.size, 12.5)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(clear=False), Coef.uniform(dd), out=out)
    assert ctx.last_kernel() == "syrk_dmma<nn=20,pat=2,gen>"
    np.testing.assert_allclose(out.reshape(base.shape), base + 12.5, rtol=1e-13, atol=0)
    args = CommonArgs(ii0=3, jj0=1, nrow=65, ncol=62)
    out = np.full(mesh.ncell * 65 * 62, 99.0)
    integ.mat_10_bdb(ctx, dmesh, None, args, Coef.uniform(dd), out=out)
    assert ctx.last_kernel() == "syrk_dmma<nn=20,pat=2,gen>"
    blk = out.reshape(mesh.ncell, 62, 65)  # [cell][col][row]
    np.testing.assert_array_equal(blk[:, 1:61, 3:63], base)  # same arithmetic, another store path
    outside = blk.copy()
    outside[:, 1:61, 3:63] = 0.0
    assert np.count_nonzero(outside) == 0
    # accumulate into an offset window: everything outside stays as it was
    out2 = np.full(out.size, 2.0)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(ii0=3, jj0=1, nrow=65, ncol=62, clear=False), Coef.uniform(dd), out=out2)
    blk2 = out2.reshape(mesh.ncell, 62, 65)
    np.testing.assert_allclose(blk2[:, 1:61, 3:63], base + 2.0, rtol=1e-13, atol=0)
    rest = blk2.copy()
    rest[:, 1:61, 3:63] = 2.0
    assert np.all(rest == 2.0)


def test_hex20_moduli_per_marker_and_per_cell(ctx):
    import torch

    mesh = bulged_hex20((3, 3, 2))
    rng = np.random.default_rng(41)
    mesh.markers[:] = rng.integers(1, 4, mesh.ncell)
    dmesh = ctx.upload(mesh)
    for structure in ("isotropic", "symmetric", "general"):
        if structure == "isotropic":
            recs = np.stack([g.mandel_matrix(g.lin_elasticity(100.0 * (k + 1), 0.1 * (k + 1), False)) for k in range(3)])
        else:
            recs = random_coef("mat_10_bdb", 3, 3, rng, spd=(structure == "symmetric"))
        per_cell = recs[mesh.markers - 1]
        ref = oracle_integ("mat_10_bdb", mesh, per_cell, cell_stride=36)
        sizes = np.arange(mesh.ncell + 1) * 3600
        got = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.per_marker([1, 2, 3], recs))
        assert ctx.last_kernel().startswith("syrk_dmma<nn=20") and ",rec" in ctx.last_kernel()
        assert rel_err(got, ref, sizes) <= TOL, structure
        out = torch.empty(mesh.ncell * 3600, dtype=torch.float64, device="cuda")
        integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.per_cell(torch.from_numpy(np.ascontiguousarray(per_cell)).cuda()), out=out)
        ctx.sync()
        assert rel_err(out.cpu().numpy(), ref, sizes) <= TOL, structure + " (device records)"


def test_mixed_3d_mesh_with_hex20(ctx):
    """A 3D mesh of Hex8 and Hex20 cells: the tensor-core kernel writes its cells at their offsets in the mixed layout."""
    h20 = bulged_hex20((2, 2, 2))
    c20 = h20.conn_matrix()
    cells = []
    for e in range(h20.ncell):
        cells.append((GeoKind.Hex20, c20[e]) if e % 3 != 1 else (GeoKind.Hex8, c20[e, :8]))
    mesh = Mesh.from_cells(3, h20.coords, cells)
    dmesh = ctx.upload(mesh)
    dd = iso(3)
    off = integ.out_offsets(dmesh, "mat_10_bdb", None, CommonArgs())
    got = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd))
    ref = oracle_integ("mat_10_bdb", mesh, dd)
    assert rel_err(got, ref, off) <= TOL
    sub = integ.mat_10_bdb(ctx, dmesh, (2, 7), CommonArgs(), Coef.uniform(dd))
    np.testing.assert_array_equal(sub, got[int(off[2]):int(off[7])])


def test_nan_coordinates_do_not_leak_into_other_cells(ctx):
    """A cell slot reuses one shared-memory region for the gradients, the geometric tensors and the staged block of
    successive cells: a non-finite cell must not contaminate the cells that follow it in the slot."""
    mesh = bulged_hex20((2, 2, 4))
    bad = 3
    pts = mesh.conn_matrix()[bad]
    own = [int(q) for q in pts if np.count_nonzero(mesh.conn_matrix() == q) == 1]
    mesh.coords[own[0], 0] = np.nan
    dmesh = ctx.upload(mesh)
    dd = iso(3)
    got = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd)).reshape(mesh.ncell, -1)
    assert np.all(np.isnan(got[bad]).any())
    ok = [e for e in range(mesh.ncell) if e != bad]
    assert np.all(np.isfinite(got[ok]))


@pytest.mark.parametrize("kind,op", [("hex8", "mat_10_bdb"), ("hex20", "mat_10_bdb"), ("qua4", "mat_10_bdb"), ("tet10", "mat_01_nsn"),
                                     ("qua8", "mat_03_btb")])
def test_packed_upper_triangles(ctx, kind, op):
    """gl_args.packed: symmetric element matrices leave as packed upper triangles (n(n+1)/2 doubles per cell) -- same values, bit
    for bit, as the full blocks; host (chunked, pipelined) and device outputs."""
    import torch

    mesh = make_mesh(kind, n=4)
    dmesh = ctx.upload(mesh)
    rng = np.random.default_rng(9)
    coef = iso(mesh.ndim) if op == "mat_10_bdb" else random_coef(op, mesh.ndim, 1, rng)[0]
    n = int(round(np.sqrt(integ.op_out_size(op, GeoKind.from_str(kind), mesh.ndim))))
    tri = n * (n + 1) // 2
    full = getattr(integ, op)(ctx, dmesh, None, CommonArgs(), Coef.uniform(coef)).reshape(mesh.ncell, n, n)  # [cell][col][row]
    args = CommonArgs(packed=True)
    assert integ.out_total(dmesh, op, None, args) == mesh.ncell * tri
    np.testing.assert_array_equal(integ.out_offsets(dmesh, op, (1, 4), args), np.arange(4) * tri)
    host = getattr(integ, op)(ctx, dmesh, None, args, Coef.uniform(coef))
    dev = torch.full((mesh.ncell * tri,), -1.0, dtype=torch.float64, device="cuda")
    getattr(integ, op)(ctx, dmesh, (2, mesh.ncell - 1), args, Coef.uniform(coef), out=dev)
    ctx.sync()
    jj, ii = np.nonzero(np.triu(np.ones((n, n))).T)  # column by column: (col jj, row ii <= jj)
    expect = full[:, jj, ii]
    np.testing.assert_array_equal(host.reshape(mesh.ncell, tri), expect)
    got = dev.cpu().numpy()
    np.testing.assert_array_equal(got[:(mesh.ncell - 3) * tri].reshape(-1, tri), expect[2:mesh.ncell - 1])
    assert np.all(got[(mesh.ncell - 3) * tri:] == -1.0)
    sym = integ.unpack_symmetric(host, n).reshape(mesh.ncell, n, n)
    np.testing.assert_allclose(sym, full, rtol=0, atol=1e-12 * np.max(np.abs(full)))
    with pytest.raises(g.GemlabError):
        getattr(integ, op)(ctx, dmesh, None, CommonArgs(packed=True, clear=False), Coef.uniform(coef), out=np.zeros(mesh.ncell * tri))


# ---------------------------------------------------------------------------------------------------------------
# one modulus per Gauss point (the reference closure's natural use, mat_10_bdb.rs:156-159) on the tuned kernels
# ---------------------------------------------------------------------------------------------------------------

PG_KINDS = ["tri3", "tri6", "qua4", "qua8", "qua9", "tet4", "tet10", "hex8", "hex20"]


@pytest.mark.parametrize("kind", PG_KINDS)
@pytest.mark.parametrize("structure", ["symmetric", "general", "one_general_record"])
def test_stiffness_with_one_modulus_per_gauss_point(ctx, kind, structure):
    """D_p differs at every Gauss point: K = sum_p c_p M_p^T D_p M_p (add_to_kk, mat_10_bdb.rs:179-233).  Symmetric records run
    the mirrored variant, a single non-symmetric record anywhere in the range makes the general variant redo the call (the
    check happens on the device); both must match the oracle and stay off the generic kernel."""
    import torch

    mesh = make_mesh(kind, n=4)
    ng = g.Gauss.new(GeoKind.from_str(kind)).npoint()
    rng = np.random.default_rng(23)
    ncell, ns = mesh.ncell, 2 * mesh.ndim
    recs = random_coef("mat_10_bdb", mesh.ndim, ncell * ng, rng, spd=(structure != "general"))
    if structure == "one_general_record":
        recs[(ncell * ng * 2) // 3, 1] += 0.25  # D[1,0] != D[0,1] in one record only
    dmesh = ctx.upload(mesh)
    bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), mesh.ndim)
    sizes = np.arange(ncell + 1) * bs
    for axisym in ([False, True] if mesh.ndim == 2 else [False]):
        args = CommonArgs(axisymmetric=axisym, alpha=1.3)
        ref = oracle_integ("mat_10_bdb", mesh, recs, cell_stride=ng * ns * ns, gp_stride=ns * ns, axisymmetric=axisym, alpha=1.3)
        got = integ.mat_10_bdb(ctx, dmesh, None, args, Coef.per_gauss(recs))
        fam = ctx.last_kernel()
        assert fam.startswith("hex8_gauss") or (fam.startswith("bdb<") and ",gauss" in fam), f"{kind}: ran {fam!r}"
        assert rel_err(got, ref, sizes) <= TOL, f"{kind} {structure} axisym={axisym}"
        # device-resident records, device output, a sub-range that does not start at a tile boundary
        b, e = 3, ncell - 2
        out = torch.full(((e - b) * bs,), 7.0, dtype=torch.float64, device="cuda")
        integ.mat_10_bdb(ctx, dmesh, (b, e), args, Coef.per_gauss(torch.from_numpy(recs[b * ng:e * ng]).cuda()), out=out)
        ctx.sync()
        assert rel_err(out.cpu().numpy(), ref[b * bs:e * bs], sizes[: e - b + 1]) <= TOL, f"{kind} {structure} axisym={axisym} range"


def test_per_gauss_moduli_constant_records_equal_uniform_path(ctx):
    """The same D at every Gauss point must reproduce the uniform-D kernels (another algebra: geometric tensors) to rounding."""
    for kind in ["hex8", "qua8", "tet10"]:
        mesh = make_mesh(kind, n=3)
        ng = g.Gauss.new(GeoKind.from_str(kind)).npoint()
        dd = iso(mesh.ndim)
        dmesh = ctx.upload(mesh)
        uni = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.uniform(dd))
        pg = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), Coef.per_gauss(np.tile(np.asarray(dd).ravel(), (mesh.ncell * ng, 1))))
        bs = integ.op_out_size("mat_10_bdb", GeoKind.from_str(kind), mesh.ndim)
        assert rel_err(pg, uni, np.arange(mesh.ncell + 1) * bs) <= TOL


@pytest.mark.parametrize("kind", ["hex8", "tet10", "qua8", "tri6", "tet4", "qua9"])
@pytest.mark.parametrize("mode", ["uniform", "per_gauss"])
def test_general_blocks_and_accumulation_on_the_register_blocked_kernel(ctx, kind, mode):
    """CommonArgs nrow / ncol / ii0 / jj0 and clear = false (src/integ/common_args.rs:5-43; kk.fill(0.0) clears the WHOLE block,
    mat_10_bdb.rs:141-143) are served by the store epilogue of the tuned kernel, not by the generic one."""
    mesh = make_mesh(kind, n=3)
    dmesh = ctx.upload(mesh)
    gk = GeoKind.from_str(kind)
    nd = gk.nnode() * mesh.ndim
    ng = g.Gauss.new(gk).npoint()
    rng = np.random.default_rng(4)
    ns = 2 * mesh.ndim
    if mode == "uniform":
        coef = Coef.uniform(iso(mesh.ndim))
    else:
        coef = Coef.per_gauss(random_coef("mat_10_bdb", mesh.ndim, mesh.ncell * ng, rng))
    base = integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(), coef).reshape(mesh.ncell, nd, nd)
    # the tight call may run on another kernel family (Qua8: the tensor-core kernel) whose sums round differently
    tol = dict(rtol=1e-12, atol=1e-13 * float(np.abs(base).max()))
    out = np.full(base.size, 12.5)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(clear=False), coef, out=out)
    fam = ctx.last_kernel()
    assert not fam.startswith("generic"), fam
    assert fam.endswith(",gen>"), fam
    np.testing.assert_allclose(out.reshape(base.shape), base + 12.5, **tol)
    nrow, ncol = nd + 5, nd + 2
    args = CommonArgs(ii0=3, jj0=1, nrow=nrow, ncol=ncol)
    out = np.full(mesh.ncell * nrow * ncol, 99.0)
    integ.mat_10_bdb(ctx, dmesh, None, args, coef, out=out)
    assert ctx.last_kernel().endswith(",gen>"), ctx.last_kernel()
    blk = out.reshape(mesh.ncell, ncol, nrow)  # [cell][col][row]
    np.testing.assert_allclose(blk[:, 1:1 + nd, 3:3 + nd], base, **tol)
    outside = blk.copy()
    outside[:, 1:1 + nd, 3:3 + nd] = 0.0
    assert np.count_nonzero(outside) == 0
    out2 = np.full(out.size, 2.0)
    integ.mat_10_bdb(ctx, dmesh, None, CommonArgs(ii0=3, jj0=1, nrow=nrow, ncol=ncol, clear=False), coef, out=out2)
    blk2 = out2.reshape(mesh.ncell, ncol, nrow)
    np.testing.assert_allclose(blk2[:, 1:1 + nd, 3:3 + nd], base + 2.0, **tol)
    rest = blk2.copy()
    rest[:, 1:1 + nd, 3:3 + nd] = 2.0
    assert np.all(rest == 2.0)


@pytest.mark.parametrize("rule", [1, 4, 5, 8, 14, 15])
@pytest.mark.parametrize("structure", ["isotropic", "symmetric", "general"])
def test_tet10_tensor_core_stiffness(ctx, rule, structure):
    """Tet10 mat_10_bdb through the DMMA rank-ng kernel (30 dofs -> 4 x 4 dof tiles, 14 cell slots per CTA) vs the oracle, every rule it
    accepts, the three modulus patterns, ranges that do not fill the slots."""
    mesh = make_mesh("tet10", n=3)
    rng = np.random.default_rng(77)
    dd = iso(3) if structure == "isotropic" else random_coef("mat_10_bdb", 3, 1, rng, spd=(structure == "symmetric"))[0]
    dmesh = ctx.upload(mesh)
    ref = oracle_integ("mat_10_bdb", mesh, dd, ngauss=rule)
    pat = {"isotropic": 2, "symmetric": 1, "general": 0}[structure]
    for b, e in [(0, mesh.ncell), (0, 1), (3, 9), (5, 40), (1, mesh.ncell - 1)]:
        got = integ.mat_10_bdb(ctx, dmesh, (b, e), CommonArgs(ngauss=rule), Coef.uniform(dd))
        assert ctx.last_kernel() == f"syrk_dmma<nn=10,pat={pat}>"
        assert rel_err(got, ref[b * 900:e * 900], np.arange(e - b + 1) * 900) <= TOL, f"rule {rule} range {b}:{e}"

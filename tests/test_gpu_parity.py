"""GPU parity tests proper: every entry of the operator surface, called through the C ABI on a B200, against
the CPU oracle on the same seeded inputs.  FP64, -fmad=false and the Fortran evaluation order on both sides =>
the bar is BIT-EXACT (np.array_equal) for every kernel except order-dependent sums (p-norms, dot), where the
tolerance is 1e-13 relative.  The north-star tolerance (1e-10 relative L2 after fixed V-cycles) is asserted too."""
import numpy as np
import pytest

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests.problem import GpuSide, OracleSide, fields_equal, rel_l2

pytestmark = pytest.mark.gpu

CASES = [  # (config, scale, max_box override)
    ("C1", 1, None), ("C1", 4, None), ("C2", 2, None), ("C3", 1, None), ("C4", 1, None), ("C5", 1, None),
]


def make(ctx, name, scale=1, max_box=None, **kw):
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, max_box or cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes, **kw)
    orc.init_bcoef()
    gpu = GpuSide(ctx, orc)
    return cfg, orc, gpu


def assert_same(gpu_ld, orc_f, what):
    d, eq = fields_equal(gpu_ld, orc_f)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


@pytest.mark.parametrize("name,scale,mb", CASES)
def test_upload_download_roundtrip(gpu_ctx, name, scale, mb):
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    for k in ("head", "B", "rhs", "bX", "bY"):
        assert_same(gpu.F[k], orc.F[k], f"roundtrip {k}")
        # per-box path, ghosts included where they lie outside the level
        b = len(orc.boxes) - 1
        got = gpu.F[k].download_box(b)
        exp = orc.F[k].fab(b)[0]
        ng = gpu.F[k].ng
        core = (slice(None), slice(ng, exp.shape[1] - ng), slice(ng, exp.shape[2] - ng))
        assert np.array_equal(got[core], exp[core])


def set_mode(ctx, mode):
    """relax mode; "1t" = the default mode with its two-iterations-per-launch smoother on every level (tune key 19), not only above 2 M cells"""
    ctx.set_relax_mode(1 if mode == "1t" else mode)
    ctx.set_tuning(19, 1 if mode == "1t" else 0)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5, "1t"])
@pytest.mark.parametrize("name,scale,mb", CASES)
def test_relax_bit_exact(gpu_ctx, name, scale, mb, mode):
    """a1/a2/a8: levelGSRB with fused NL + lambda; generic colour passes (mode 0), cp.async-staged streaming sweep (mode 1), register-only fused sweep (mode 2), two iterations per sweep (modes 3, 4; mode 5 = the lean k_gsrb_twin with the deferred division slow path)."""
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    set_mode(gpu_ctx, mode)
    try:
        oop = orc.op()
        gop = gpu.factory.AMRnewOp(0)
        for n in (1, 3, 4):
            oop.relax(orc.F["head"], orc.F["rhs"], n)
            gop.relax(gpu.F["head"], gpu.F["rhs"], n)
            assert_same(gpu.F["head"], orc.F["head"], f"relax x{n} mode {mode}")
    finally:
        set_mode(gpu_ctx, 1)


def test_relax_inhomogeneous_bc_values(gpu_ctx):
    """non-zero Dirichlet / Neumann values exercise the on-the-fly BC of the fused sweep"""
    cfg = syn.config("C5", 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, 64, 2)
    orc = OracleSide(cfg, boxes, bc_vals=((1500.0, 1e-3), (-2e-3, 900.0)))
    orc.init_bcoef()
    gpu = GpuSide(gpu_ctx, orc)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    for mode in (0, 1, 2, 3, 4, 5):
        gpu_ctx.set_relax_mode(mode)
        oop.relax(orc.F["head"], orc.F["rhs"], 2)
        gop.relax(gpu.F["head"], gpu.F["rhs"], 2)
        assert_same(gpu.F["head"], orc.F["head"], f"relax inhomogeneous BC mode {mode}")
    gpu_ctx.set_relax_mode(1)


@pytest.mark.parametrize("name,scale,mb", CASES)
def test_residual_apply_bit_exact(gpu_ctx, name, scale, mb):
    """a3/a4: residual, applyOp (inhomogeneous and homogeneous via applyOpMg), max-norm"""
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    ores, gres = ob.Field(orc.layout, 1, 0), gpu.new_like("rhs")
    oop.residual(ores, orc.F["head"], orc.F["rhs"])
    gop.residual(gres, gpu.F["head"], gpu.F["rhs"])
    assert_same(gres, ores, "residual")
    assert gop.norm(gres, 0) == ores.norm(0)
    assert gop.localMaxNorm(gres) == ores.norm(0)
    for p in (1, 2):
        assert gop.norm(gres, p) == pytest.approx(ores.norm(p), rel=1e-13)
    oop.apply(ores, orc.F["head"], False)
    gop.applyOp(gres, gpu.F["head"], False)
    assert_same(gres, ores, "applyOp")
    oop.apply(ores, orc.F["head"], True)
    gop.applyOpMg(gres, gpu.F["head"], None, True)
    assert_same(gres, ores, "applyOpMg homogeneous")
    # the reference aborts on a homogeneous residualI / restrictResidual / UpdateOperator
    from suhmo_b200.capi import SuhmoGpuError, ERR_ABORT
    with pytest.raises(SuhmoGpuError) as e:
        gop.residualNF(gres, gpu.F["head"], None, gpu.F["rhs"], True)
    assert e.value.code == ERR_ABORT


@pytest.mark.parametrize("name,scale,mb", CASES)
def test_restrict_prolong_bit_exact(gpu_ctx, name, scale, mb):
    """a5/a6/a7: restrictResidual, restrictR, prolongIncrement"""
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    olc = orc.layout.coarsen(2)
    oresc, ophic = ob.Field(olc, 1, 0), ob.Field(olc, 1, 1)
    gphic = gop.createCoarser(gpu.F["head"])
    gresc = gop.createCoarser(gpu.F["rhs"])
    oop.restrict_residual(oresc, orc.F["head"], orc.F["rhs"])
    gop.restrictResidual(gresc, gpu.F["head"], None, gpu.F["rhs"], False)
    assert_same(gresc, oresc, "restrictResidual")
    oop.restrict_r(ophic, orc.F["head"])
    gop.restrictR(gphic, gpu.F["head"])
    assert_same(gphic, ophic, "restrictR")
    oop.prolong_increment(orc.F["head"], ophic)
    gop.prolongIncrement(gpu.F["head"], gphic)
    assert_same(gpu.F["head"], orc.F["head"], "prolongIncrement")


@pytest.mark.parametrize("name,scale,mb", CASES)
def test_update_and_average_operator_bit_exact(gpu_ctx, name, scale, mb):
    """a9/a10/a11: UpdateOperator (B(h): gradient, ExtrapGhostCells, Re, CellToEdge, ice mask, bcoef), the factory's
    MG operators (coefficient averaging) and AverageOperator"""
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    # perturb head so the update differs from the initial coefficients
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    oop.relax(orc.F["head"], orc.F["rhs"], 1)
    gop.relax(gpu.F["head"], gpu.F["rhs"], 1)
    oop.update_operator(orc.F["head"])
    gop.UpdateOperator(gpu.F["head"], None, 0, 0, False)
    assert_same(gpu.F["bX"], orc.F["bX"], "UpdateOperator bX")
    assert_same(gpu.F["bY"], orc.F["bY"], "UpdateOperator bY")
    glam = gpu.new_like("rhs")
    gop.lambda_(glam)
    assert_same(glam, oop.lambda_field(), "lambda")
    osol = orc.solver()
    for depth in range(1, osol.depth):
        gd = gpu.factory.MGnewOp(0, depth)
        assert gd is not None
        from oracle.binding import lib as olib, Field
        odp = ob.Op(None, None, 0, 0, None, None, None, None, None, None, None, None, None, _h=olib().orc_solver_op(osol.h, depth))
        # coefficient sets built by MGnewOp: compare through lambda (uses bX,bY) and a relax (uses B,Pi,zb,mask)
        olay = orc.layout.coarsen(2 ** depth)
        odp.layout = olay
        ophi, orhs = ob.Field(olay, 1, 1), ob.Field(olay, 1, 0)
        rng = np.random.RandomState(depth)
        dom = olay.domain
        gph = 500.0 + rng.rand(dom[3] + 3, dom[2] + 3)
        grh = 1e-9 * rng.rand(dom[3] + 1, dom[2] + 1)
        ophi.set_global(gph, (-1, -1)); orhs.set_global(grh, (0, 0))
        gphi = gpu.amr.LevelData(gd.layout, 1, 1); grhs = gpu.amr.LevelData(gd.layout, 1, 0)
        gphi.set_global(gph, (-1, -1)); grhs.set_global(grh, (0, 0))
        odp.relax(ophi, orhs, 1); gd.relax(gphi, grhs, 1)
        assert_same(gphi, ophi, f"MGnewOp depth {depth} relax")
        odp.average_operator(oop, depth); gd.AverageOperator(gop, depth)
        odp.relax(ophi, orhs, 1); gd.relax(gphi, grhs, 1)
        assert_same(gphi, ophi, f"AverageOperator depth {depth} relax")
    assert gpu.factory.MGnewOp(0, osol.depth) is None  # cannot coarsen further -> NULL


def test_vector_ops(gpu_ctx):
    """a12: assign, incr, axby, scale, setToZero, dotProduct"""
    cfg, orc, gpu = make(gpu_ctx, "C3", 1, None)
    gop = gpu.factory.AMRnewOp(0)
    L = ob.lib()
    ox, oy = ob.Field(orc.layout, 1, 0), ob.Field(orc.layout, 1, 0)
    gx, gy = gpu.new_like("rhs"), gpu.new_like("rhs")
    L.orc_assign(ox.h, orc.F["rhs"].h); gop.assign(gx, gpu.F["rhs"])
    assert_same(gx, ox, "assign")
    L.orc_scale(ox.h, 3.25); gop.scale(gx, 3.25)
    L.orc_incr(ox.h, orc.F["rhs"].h, -0.5); gop.incr(gx, gpu.F["rhs"], -0.5)
    assert_same(gx, ox, "scale+incr")
    L.orc_axby(oy.h, ox.h, orc.F["rhs"].h, 2.0, -7.0); gop.axby(gy, gx, gpu.F["rhs"], 2.0, -7.0)
    assert_same(gy, oy, "axby")
    assert gop.dotProduct(gx, gy) == pytest.approx(L.orc_dot(ox.h, oy.h), rel=1e-13)
    gop.setToZero(gy)
    assert np.all(gy.get_global() == 0.0)


@pytest.mark.parametrize("name,scale,mb", CASES + [("C5", 2, None)])
def test_fixed_vcycles_parity(gpu_ctx, name, scale, mb):
    """parity protocol of SURVEY.md 8d: identical inputs, a fixed number of FAS V-cycles, compare head by
    relative L2 (<= 1e-10 demanded; bit-exact expected) and the residual history."""
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    ncyc = 5
    osp = ob.make_solver_params(bottom=10, fixed_cycles=ncyc)
    it, ohist = orc.solver().solve(orc.F["head"], orc.F["rhs"], osp)
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=ncyc)
    assert git == it == ncyc
    oh, gh = orc.F["head"].get_global(), gpu.F["head"].get_global()
    assert rel_l2(gh, oh) <= 1e-10
    assert np.array_equal(gh, oh), f"head differs, max {np.abs(gh - oh).max():g}"
    assert np.array_equal(ghist, ohist), (ghist, ohist)
    assert stats.kernel_launches > 0 and stats.cell_updates == ncyc * mg.cell_updates_per_cycle()


@pytest.mark.parametrize("name,scale", [("C2", 2), ("C5", 1)])
def test_solve_with_stop_test(gpu_ctx, name, scale):
    """AMRMultiGrid stop logic (eps, hang, normThresh, imin, iterMin): same iteration count and history"""
    cfg, orc, gpu = make(gpu_ctx, name, scale)
    osp = ob.make_solver_params(bottom=10, eps=1e-10, hang=1e-4, imin=20, iter_min=2, max_iter=100)
    it, ohist = orc.solver().solve(orc.F["head"], orc.F["rhs"], osp)
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    mg.params.imin, mg.params.iter_min = 20, 2
    git, ghist, stats = mg.solve([gpu.F["head"]], [gpu.F["rhs"]])
    assert git == it
    assert np.array_equal(ghist, ohist)
    assert np.array_equal(gpu.F["head"].get_global(), orc.F["head"].get_global())


def test_box_decomposition_invariance(gpu_ctx):
    """GSRB colour = parity of the global index => results do not depend on how the level is cut into boxes"""
    cfg = syn.config("C5", 1)
    res = []
    for mb in (64, 32):
        boxes = syn.domain_split(cfg.nx, cfg.ny, mb, 2)
        orc = OracleSide(cfg, boxes)
        orc.init_bcoef()
        gpu = GpuSide(gpu_ctx, orc)
        gop = gpu.factory.AMRnewOp(0)
        gop.relax(gpu.F["head"], gpu.F["rhs"], 3)
        res.append(gpu.F["head"].get_global())
    assert np.array_equal(res[0], res[1])


def test_ghost_utilities(gpu_ctx):
    """a15/a16: mixBCValues, ExtrapGhostCells, CopyGhostCells on the level's domain ghosts"""
    cfg, orc, gpu = make(gpu_ctx, "C3", 1, None, bc_vals=((3.0, 0.5), (-1.0, 2.0)))
    amr = gpu.amr
    L = ob.lib()
    import ctypes as C
    dx = np.array(cfg.dx)
    for homog in (0, 1):
        L.orc_apply_bc(orc.F["head"].h, C.byref(orc.bc), dx.ctypes.data_as(C.POINTER(C.c_double)), homog)
        amr.check(amr.lib().sg_apply_bc(gpu.F["head"].h, C.byref(gpu.bc), dx.ctypes.data_as(C.POINTER(C.c_double)), homog))
        for b in (0, len(orc.boxes) - 1):
            got, exp = gpu.F["head"].download_box(b), orc.F["head"].fab(b)[0]
            assert np.array_equal(got[:, 1:-1, :], exp[:, 1:-1, :]) and np.array_equal(got[:, :, 1:-1], exp[:, :, 1:-1])
    for fn_o, fn_g in ((L.orc_extrap_ghost, amr.ExtrapGhostCells), (L.orc_copy_ghost, amr.CopyGhostCells)):
        fn_o(orc.F["B"].h); fn_g(gpu.F["B"])
        for b in (0, len(orc.boxes) - 1):
            got, exp = gpu.F["B"].download_box(b), orc.F["B"].fab(b)[0]
            assert np.array_equal(got, exp)


def test_against_golden_fixture_c1(gpu_ctx):
    """the CUDA path against the committed fixture (tests/golden/c1_1lev_vcycles.npz): QuickStart 1lev, 5 FAS V-cycles"""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "c1_1lev_vcycles.npz"))
    cfg, orc, gpu = make(gpu_ctx, "C1", 1, None)
    assert np.array_equal(gpu.F["bX"].get_global(), z["bX0"]) and np.array_equal(gpu.F["bY"].get_global(), z["bY0"])
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=5)
    assert np.array_equal(ghist, z["resnorm"])
    assert np.array_equal(gpu.F["head"].get_global(), z["head5"])


@pytest.mark.parametrize("use_nl", [1, 0])
@pytest.mark.parametrize("name,scale", [("C2", 2), ("C5", 1)])
def test_helmholtz_alpha_and_linear_variants(gpu_ctx, name, scale, use_nl):
    """alpha != 0 (the aCoef-reading kernel variants; what the implicit gap solve's operator needs: alpha = 1, NL = 0) and
    solver.use_NL = false: relax in every mode, residual, restriction and fixed V-cycles stay bit-exact"""
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes, prm_over=dict(use_NL=use_nl, bcoeff_otf=0))
    orc.alpha = 0.75
    rng = np.random.RandomState(7)
    orc.F["a"].set_global(1e-9 * (1.0 + rng.rand(cfg.ny, cfg.nx)), (0, 0))
    orc.init_bcoef()
    gpu = GpuSide(gpu_ctx, orc)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    for mode in (0, 1, 3, 4, 5):
        gpu_ctx.set_relax_mode(mode)
        oop.relax(orc.F["head"], orc.F["rhs"], 2)
        gop.relax(gpu.F["head"], gpu.F["rhs"], 2)
        assert_same(gpu.F["head"], orc.F["head"], f"relax alpha!=0 mode {mode}")
    gpu_ctx.set_relax_mode(1)
    ores, gres = ob.Field(orc.layout, 1, 0), gpu.new_like("rhs")
    oop.residual(ores, orc.F["head"], orc.F["rhs"])
    gop.residual(gres, gpu.F["head"], gpu.F["rhs"])
    assert_same(gres, ores, "residual alpha!=0")
    glam = gpu.new_like("rhs")
    gop.lambda_(glam)
    assert_same(glam, oop.lambda_field(), "lambda alpha!=0")
    osp = ob.make_solver_params(bottom=10, fixed_cycles=3)
    it, ohist = orc.solver().solve(orc.F["head"], orc.F["rhs"], osp)
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=3)
    assert np.array_equal(ghist, ohist), (ghist, ohist)
    assert_same(gpu.F["head"], orc.F["head"], "V-cycles alpha!=0")


def test_error_conventions(gpu_ctx):
    """where the reference aborts or asserts, the C ABI returns a status: SG_ERR_ABORT for the MayDay::Abort sites,
    SG_ERR_INVALID for CH_assert-style argument errors -- never a silent fallback"""
    from suhmo_b200.capi import ERR_ABORT, ERR_INVALID, SuhmoGpuError
    cfg, orc, gpu = make(gpu_ctx, "C3", 1, None)
    cfg2, orc2, gpu2 = make(gpu_ctx, "C5", 1, None)
    gop = gpu.factory.AMRnewOp(0)
    res = gpu.new_like("rhs")
    for call in (lambda: gop.restrictResidual(gop.createCoarser(gpu.F["rhs"]), gpu.F["head"], None, gpu.F["rhs"], True),
                 lambda: gop.UpdateOperator(gpu.F["head"], None, 0, 0, True),
                 lambda: gop.AMRResidualNF(res, gpu.F["head"], None, gpu.F["rhs"], True)):
        with pytest.raises(SuhmoGpuError) as e:
            call()
        assert e.value.code == ERR_ABORT
    with pytest.raises(SuhmoGpuError) as e:   # a field of another level handed to the operator
        gop.relax(gpu2.F["head"], gpu.F["rhs"], 1)
    assert e.value.code == ERR_INVALID
    with pytest.raises(SuhmoGpuError) as e:   # no coarser level, but a coarse phi is passed
        gop.relaxNF(gpu.F["head"], gpu.F["head"], gpu.F["rhs"], 1)
    assert e.value.code == ERR_INVALID
    assert gpu.factory.refToFiner(0) == 2
    with pytest.raises(SuhmoGpuError) as e:
        gpu.factory.refToFiner(3)              # "Domain not found in AMR hierarchy"
    assert e.value.code == ERR_ABORT

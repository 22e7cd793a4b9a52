"""GPU parity of the entry points round 1 left without a test (stand-alone field kernels, the NC / NoBoundary / UpdateResidual
variants) and of the virtuals of the cited classes that the FAS path never calls (AMRRestrict, AMRProlong, preCond, getFlux,
finerOperatorChanged, mDotProduct, buildCopier / assignCopier, setAlphaAndBeta, diagonalScale, divideByIdentityCoef,
homogeneousCFInterp).  Everything through the C ABI against the CPU oracle, bit-exact."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests.problem import (AmrGpuSide, AmrOracleSide, GpuSide, OracleSide, amr_hierarchy, fabs_equal, fields_equal)

pytestmark = pytest.mark.gpu
CELL, XFACE, YFACE = 0, 1, 2


def make(ctx, name, scale=1, **kw):
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes, **kw)
    orc.init_bcoef()
    return cfg, orc, GpuSide(ctx, orc)


def make_amr(ctx, nlev=3, **kw):
    cfg, lv = amr_hierarchy()
    orc = AmrOracleSide(cfg, lv[:nlev], **kw)
    orc.average_down("head")
    orc.init_bcoef()
    return cfg, orc, AmrGpuSide(ctx, orc)


def same(gpu_ld, orc_f, what, ghosts=False):
    d, eq = fabs_equal(gpu_ld, orc_f) if ghosts else fields_equal(gpu_ld, orc_f)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


def twin(gpu, orc, ncomp, ng, cent=CELL, rng=None, lo=-1.0, hi=1.0):
    """a pair (device field, oracle field) holding the same random numbers, drawn once per cell / face of the level (ghost
    ring included) so that what two boxes share -- a face, a ghost cell that is the neighbour's valid cell -- has one value:
    the device stores a uniform level as one merged rectangle"""
    of = ob.Field(orc.layout, ncomp, ng, cent)
    gf = gpu.amr.LevelData(gpu.layout, ncomp, ng, cent)
    if rng is not None:
        d = orc.layout.domain
        nx, ny = d[2] - d[0] + 1 + 2 * ng + (cent == XFACE), d[3] - d[1] + 1 + 2 * ng + (cent == YFACE)
        g = rng.uniform(lo, hi, size=(ncomp, ny, nx))
        of.set_global(g, (d[0] - ng, d[1] - ng))
        gf.set_global(g, (d[0] - ng, d[1] - ng))
    return gf, of


@pytest.mark.parametrize("name", ["C1", "C4", "C5"])
def test_divergence_and_nonlinear_level(gpu_ctx, name):
    """util/DivergenceF.ChF:23-57 (north-star "Gradient/Divergence") and NonLinear_level as stand-alone calls"""
    cfg, orc, gpu = make(gpu_ctx, name)
    rng = np.random.RandomState(7)
    gux, oux = twin(gpu, orc, 1, 0, XFACE, rng)
    guy, ouy = twin(gpu, orc, 1, 0, YFACE, rng)
    gdiv, odiv = twin(gpu, orc, 1, 0, CELL, rng)
    dx = np.ascontiguousarray(cfg.dx, dtype=np.float64)
    ob.lib().orc_divergence(oux.h, ouy.h, dx.ctypes.data_as(C.POINTER(C.c_double)), odiv.h)
    from suhmo_b200.capi import check, lib
    check(lib().sg_divergence(gdiv.h, gux.h, guy.h, dx.ctypes.data_as(C.POINTER(C.c_double))))
    same(gdiv, odiv, "divergence")
    # a linear velocity field has the exact divergence a + b
    gnl, onl = twin(gpu, orc, 1, 0)
    gdn, odn = twin(gpu, orc, 1, 0)
    ob.lib().orc_compute_nl(C.byref(orc.prm), orc.F["head"].h, orc.F["B"].h, orc.F["mask"].h, orc.F["Pi"].h, orc.F["zb"].h, onl.h, odn.h)
    check(lib().sg_nonlinear_level(C.byref(gpu.prm), gnl.h, gdn.h, gpu.F["head"].h, gpu.F["B"].h, gpu.F["mask"].h, gpu.F["Pi"].h, gpu.F["zb"].h))
    same(gnl, onl, "NonLinear_level nl")
    same(gdn, odn, "NonLinear_level dnl")
    if name == "C4":
        assert (orc.F["mask"].get_global() < 0).any() and (onl.get_global()[orc.F["mask"].get_global() < 0] == 0).all()


def test_time_varying_recharge(gpu_ctx):
    """FORT_COMPUTE_TIMEVARYINGRECHARGE (src/AmrHydroF.ChF:353-373): the seasonal forcing of the SHMIP F runs (config 4)"""
    cfg, orc, gpu = make(gpu_ctx, "C4")
    rng = np.random.RandomState(3)
    gzs, ozs = twin(gpu, orc, 1, 1, CELL, rng, 0.0, 2000.0)
    grc, orc_r = twin(gpu, orc, 1, 0)
    from suhmo_b200.capi import check, lib
    for TK, bg in ((-21.0, 7.93e-11), (11.0, 7.93e-11), (4.5, 0.0)):
        ob.lib().orc_time_varying_recharge(ozs.h, orc_r.h, TK, bg)
        check(lib().sg_time_varying_recharge(gzs.h, grc.h, TK, bg))
        same(grc, orc_r, f"time-varying recharge T={TK}")
    assert orc_r.get_global().max() > 0.0  # the warm case melts somewhere


@pytest.mark.parametrize("name", ["C2", "C4", "C5"])
def test_wflx_level_and_gradient_cc(gpu_ctx, name):
    """AmrHydro::WFlx_level and Gradient::compGradientCC as stand-alone calls (single level)"""
    cfg, orc, gpu = make(gpu_ctx, name)
    from suhmo_b200.capi import check, lib
    dx = np.ascontiguousarray(cfg.dx, dtype=np.float64)
    dxp = dx.ctypes.data_as(C.POINTER(C.c_double))
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    oop.relax(orc.F["head"], orc.F["rhs"], 1)
    gop.relax(gpu.F["head"], gpu.F["rhs"], 1)
    # gradient: MAC gradient on the faces of every box, then EdgeToCell (util/Gradient.cpp:478-624)
    ob.lib().orc_exchange_faces(orc.F["head"].h)
    ob.lib().orc_apply_bc(orc.F["head"].h, C.byref(orc.bc), dxp, 0)
    gpu.F["head"].exchange(False)
    check(lib().sg_apply_bc(gpu.F["head"].h, C.byref(gpu.bc), dxp, 0))
    ogx, ogy = ob.Field(orc.layout, 1, 0, XFACE), ob.Field(orc.layout, 1, 0, YFACE)
    ograd, ggrad = ob.Field(orc.layout, 2, 1), gpu.amr.LevelData(gpu.layout, 2, 1)
    msk_o = orc.F["mask"].h if cfg.use_mask_grad else None
    msk_g = gpu.F["mask"].h if cfg.use_mask_grad else None
    ob.lib().orc_mac_gradient(orc.F["head"].h, msk_o, dxp, ogx.h, ogy.h)
    ob.lib().orc_edge_to_cell(ogx.h, ogy.h, ograd.h)
    check(lib().sg_gradient_cc(ggrad.h, gpu.F["head"].h, msk_g, dxp))
    same(ggrad, ograd, "compGradientCC")
    # WFlx_level: the whole B(h) evaluation into fresh face fields
    obX, obY = orc.F["bX"], orc.F["bY"]
    oop.update_operator(orc.F["head"])
    gbX, gbY = gpu.amr.LevelData(gpu.layout, 1, 0, XFACE), gpu.amr.LevelData(gpu.layout, 1, 0, YFACE)
    check(lib().sg_wflx_level(gpu_ctx.h, C.byref(gpu.prm), gbX.h, gbY.h, gpu.F["head"].h, None, gpu.F["B"].h, gpu.F["mask"].h, dxp))
    same(gbX, obX, "WFlx_level bX")
    same(gbY, obY, "WFlx_level bY")


def test_apply_no_boundary_and_device_view(gpu_ctx):
    cfg, orc, gpu = make(gpu_ctx, "C5")
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    ores, gres = ob.Field(orc.layout, 1, 0), gpu.new_like("rhs")
    # applyOpNoBoundary = exchange + operator, the physical ghost cells as the caller left them (here: the inhomogeneous BC fill)
    dx = np.ascontiguousarray(cfg.dx, dtype=np.float64)
    dxp = dx.ctypes.data_as(C.POINTER(C.c_double))
    ob.lib().orc_apply_bc(orc.F["head"].h, C.byref(orc.bc), dxp, 0)
    from suhmo_b200.capi import check, lib
    check(lib().sg_apply_bc(gpu.F["head"].h, C.byref(gpu.bc), dxp, 0))
    oop.apply(ores, orc.F["head"], False)  # same ghost values -> same result as the no-boundary form
    gop.applyOpNoBoundary(gres, gpu.F["head"])
    same(gres, ores, "applyOpNoBoundary")
    base, pitch, cstride, off = C.c_void_p(), C.c_longlong(), C.c_longlong(), C.c_longlong()
    plo, phi_ = (C.c_int * 2)(), (C.c_int * 2)()
    check(lib().sg_field_device_view(gpu.F["head"].h, C.byref(base), C.byref(pitch), C.byref(cstride), plo, phi_, C.byref(off)))
    assert base.value and pitch.value >= cfg.nx + 2 and cstride.value >= pitch.value * cfg.ny
    assert (plo[0], plo[1], phi_[0], phi_[1]) == (0, 0, cfg.nx - 1, cfg.ny - 1) and off.value > 0
    import torch

    class View:  # what a zero-copy caller does: wrap the raw pointer (CUDA array interface), no copy
        __cuda_array_interface__ = {"shape": (int(cstride.value),), "typestr": "<f8", "data": (int(base.value), False), "version": 2}

    gpu_ctx.sync()
    t = torch.as_tensor(View(), device="cuda:0")
    exp = gpu.F["head"].get_global()
    assert float(t[off.value]) == exp[0, 0] and float(t[off.value + 3 * pitch.value + 5]) == exp[3, 5]


def test_amr_nc_variants_and_update_residual(gpu_ctx):
    """AMRResidualNC / AMROperatorNC (no coarser level: the base level under a finer one) and AMRUpdateResidual"""
    cfg, orc, gpu = make_amr(gpu_ctx, 3)
    oops = [orc.level_op(l) for l in range(3)]
    gops = [gpu.factory.AMRnewOp(l) for l in range(3)]
    olof, glof = ob.Field(orc.layouts[0], 1, 0), gpu.new_like(0, "rhs")
    oops[0].amr_operator(olof, orc.F[1]["head"], orc.F[0]["head"], None, False, oops[1])
    gops[0].AMROperatorNC(glof, gpu.F[1]["head"], gpu.F[0]["head"], False, gops[1])
    same(glof, olof, "AMROperatorNC")
    oops[0].amr_residual(olof, orc.F[1]["head"], orc.F[0]["head"], None, orc.F[0]["rhs"], False, oops[1])
    gops[0].AMRResidualNC(glof, gpu.F[1]["head"], gpu.F[0]["head"], gpu.F[0]["rhs"], False, gops[1])
    same(glof, olof, "AMRResidualNC")
    # AMRUpdateResidual(residual, correction, coarseCorrection): residual <- residual - L(correction) (NF form, in place)
    for l in (1, 2):
        ores, gres = ob.Field(orc.layouts[l], 1, 0), gpu.new_like(l, "rhs")
        ores.copy_from(orc.F[l]["rhs"])
        gops[l].assign(gres, gpu.F[l]["rhs"])
        oops[l].amr_residual(ores, None, orc.F[l]["head"], orc.F[l - 1]["head"], ores, False, None)
        gops[l].AMRUpdateResidual(gres, gpu.F[l]["head"], gpu.F[l - 1]["head"])
        same(gres, ores, f"AMRUpdateResidual L{l}")


def test_amr_restrict_prolong_plain(gpu_ctx):
    """AMRRestrict (own scratch) and AMRProlong (own coarsened-fine copy), src/AMRNonLinearPoissonOp.cpp:1011-1025,1073-1103"""
    cfg, orc, gpu = make_amr(gpu_ctx, 3)
    oops = [orc.level_op(l) for l in range(3)]
    gops = [gpu.factory.AMRnewOp(l) for l in range(3)]
    for l in (1, 2):
        clay = orc.layouts[l].coarsen(2)
        oresC, oscr = ob.Field(clay, 1, 1), ob.Field(orc.layouts[l], 1, 1)
        gresC = gops[l].createCoarsened(gpu.F[l]["head"])
        for skip in (True, False):
            src_o = orc.F[l]["head"] if skip else orc.F[l]["rhs"]
            src_g = gpu.F[l]["head"] if skip else gpu.F[l]["rhs"]
            oops[l].amr_restrict_s(oresC, src_o, orc.F[l]["head"], orc.F[l - 1]["head"], oscr, skip)
            gops[l].AMRRestrict(gresC, src_g, gpu.F[l]["head"], gpu.F[l - 1]["head"], skip)
            same(gresC, oresC, f"AMRRestrict L{l} skip_res={skip}")
        ocorr, gcorr = ob.Field(orc.layouts[l - 1], 1, 1), gpu.new_like(l - 1, "head")
        dom = orc.layouts[l - 1].domain
        gl = np.random.RandomState(10 + l).rand(dom[3] + 3, dom[2] + 3)
        ocorr.set_global(gl, (-1, -1))
        gcorr.set_global(gl, (-1, -1))
        oops[l].amr_prolong_s(orc.F[l]["head"], ocorr, oresC)
        gops[l].AMRProlong(gpu.F[l]["head"], gcorr)
        same(gpu.F[l]["head"], orc.F[l]["head"], f"AMRProlong L{l}")


@pytest.mark.parametrize("name", ["C1", "C4", "C5"])
def test_precond_getflux_scalings(gpu_ctx, name):
    """preCond (2- and 3-argument forms), getFlux per direction, mDotProduct, setAlphaAndBeta, diagonalScale / divideByIdentityCoef"""
    cfg, orc, gpu = make(gpu_ctx, name)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    rng = np.random.RandomState(11)
    # preCond(phi, rhs): phi = rhs / lambda on the valid cells, then two GSRB iterations
    gphi, ophi = twin(gpu, orc, 1, 1, CELL, rng)
    oop.precond(ophi, orc.F["rhs"])
    gop.preCond(gphi, gpu.F["rhs"])
    same(gphi, ophi, "preCond (2 arguments)")
    oop.precond3(orc.F["head"], ophi, orc.F["rhs"])
    gop.preCond(gpu.F["head"], gphi, gpu.F["rhs"])
    same(gpu.F["head"], orc.F["head"], "preCond (3 arguments)")
    # getFlux on every face of every box, both directions, with a refinement factor and a scale
    for d, cent in ((0, XFACE), (1, YFACE)):
        gfl, ofl = twin(gpu, orc, 1, 0, cent)
        oop.get_flux(ofl, orc.F["head"], d, 2, 0.375)
        gop.getFlux(gfl, gpu.F["head"], d, 2, 0.375)
        same(gfl, ofl, f"getFlux dir {d}")
    # mDotProduct against the single dot products
    fs_g, fs_o = [], []
    for k in range(3):
        g, o = twin(gpu, orc, 1, 0, CELL, rng)
        fs_g.append(g); fs_o.append(o)
    md = gop.mDotProduct(gpu.F["rhs"], fs_g)
    for k in range(3):
        assert md[k] == gop.dotProduct(gpu.F["rhs"], fs_g[k])
        assert md[k] == pytest.approx(ob.lib().orc_dot(orc.F["rhs"].h, fs_o[k].h), rel=1e-12)
    # alpha != 0 through setAlphaAndBeta: a(x) random, residual and lambda must follow
    ga, oa = twin(gpu, orc, 1, 0, CELL, rng, 0.5, 1.5)
    gpu.F["a"].upload([oa.fab(b)[0].copy() for b in range(len(orc.boxes))])
    for b in range(len(orc.boxes)):
        orc.F["a"].fab(b)[0][...] = oa.fab(b)[0]
    oop.set_alpha_beta(1e-9, -0.75)
    gop.setAlphaAndBeta(1e-9, -0.75)
    ores, gres = ob.Field(orc.layout, 1, 0), gpu.new_like("rhs")
    oop.residual(ores, orc.F["head"], orc.F["rhs"])
    gop.residual(gres, gpu.F["head"], gpu.F["rhs"])
    same(gres, ores, "residual after setAlphaAndBeta")
    oop.relax(orc.F["head"], orc.F["rhs"], 2)
    gop.relax(gpu.F["head"], gpu.F["rhs"], 2)
    same(gpu.F["head"], orc.F["head"], "relax after setAlphaAndBeta")
    # TGA scalings with the identity coefficient
    oop.diagonal_scale(ores)
    gop.diagonalScale(gres)
    same(gres, ores, "diagonalScale")
    oop.divide_by_identity_coef(ores)
    gop.divideByIdentityCoef(gres)
    same(gres, ores, "divideByIdentityCoef")


def test_copier_surface(gpu_ctx):
    """buildCopier / assignCopier: copyTo between the coarsened fine layout and the coarser level through a prebuilt copier"""
    cfg, orc, gpu = make_amr(gpu_ctx, 2)
    gop = gpu.factory.AMRnewOp(1)
    oop = orc.level_op(1)
    clay = orc.layouts[1].coarsen(2)
    oresC, oscr = ob.Field(clay, 1, 1), ob.Field(orc.layouts[1], 1, 1)
    gresC, gscr = gop.createCoarsened(gpu.F[1]["head"]), gpu.new_like(1, "head")
    oop.amr_restrict_s(oresC, orc.F[1]["head"], orc.F[1]["head"], orc.F[0]["head"], oscr, True)
    gop.AMRRestrictS(gresC, gpu.F[1]["head"], gpu.F[1]["head"], gpu.F[0]["head"], gscr, True)
    otmp, gtmp = ob.Field(orc.layouts[0], 1, 0), gpu.new_like(0, "rhs")
    ob.copy_to(otmp, oresC)
    cop = gop.buildCopier(gtmp, gresC)
    gop.assignCopier(gtmp, gresC, cop)
    same(gtmp, otmp, "assignCopier")
    from suhmo_b200.capi import SuhmoGpuError, lib
    with pytest.raises(SuhmoGpuError):
        gop.assignCopier(gpu.F[1]["head"], gresC, cop)  # a copier built for other layouts
    lib().sg_copier_destroy(cop)


@pytest.mark.parametrize("name", ["C2", "C4", "C5"])
def test_finer_operator_changed(gpu_ctx, name):
    """finerOperatorChanged: a multigrid operator re-derives ALL its coefficients from the operator above it"""
    cfg, orc, gpu = make(gpu_ctx, name)
    osol = orc.solver()
    o0 = ob.Op(orc.layout, None, 0, 0, None, None, None, None, None, None, None, None, None, _h=ob.lib().orc_solver_op(osol.h, 0))
    o1 = ob.Op(orc.layout.coarsen(2), None, 0, 0, None, None, None, None, None, None, None, None, None, _h=ob.lib().orc_solver_op(osol.h, 1))
    g0, g1 = gpu.factory.MGnewOp(0, 0), gpu.factory.MGnewOp(0, 1)
    # change the finest data on both sides, then notify the depth-1 operators
    rng = np.random.RandomState(5)
    for k in ("B", "zb"):
        for b in range(len(orc.boxes)):
            a, _ = orc.F[k].fab(b)
            a *= 1.0 + 0.01 * rng.rand(*a.shape)
        ob.lib().orc_exchange_full(orc.F[k].h)
        gpu.push(k)
    o0.update_operator(orc.F["head"])
    g0.UpdateOperator(gpu.F["head"], None, 0, 0, False)
    o1.finer_operator_changed(o0, 2)
    g1.finerOperatorChanged(g0, 2)
    # compare through what the operator computes with them: lambda (bCoef), one relaxation (B, Pi, zb, mask) on the coarse level
    clay = orc.layout.coarsen(2)
    ophi, orhs = ob.Field(clay, 1, 1), ob.Field(clay, 1, 0)
    gphi, grhs = g0.createCoarser(gpu.F["head"]), g0.createCoarser(gpu.F["rhs"])
    o0.restrict_r(ophi, orc.F["head"]); g0.restrictR(gphi, gpu.F["head"])
    o0.restrict_residual(orhs, orc.F["head"], orc.F["rhs"]); g0.restrictResidual(grhs, gpu.F["head"], None, gpu.F["rhs"], False)
    glam = g0.createCoarser(gpu.F["rhs"])
    g1.lambda_(glam)
    ob.lib().orc_op_reset_lambda(o1.h)
    same(glam, o1.lambda_field(), "lambda after finerOperatorChanged")
    o1.relax(ophi, orhs, 2)
    g1.relax(gphi, grhs, 2)
    same(gphi, ophi, "coarse relax after finerOperatorChanged")


def test_homogeneous_cf_interp(gpu_ctx):
    """homogeneousCFInterp (dead under FAS, part of the surface): c1*near + c2*far on every coarse-fine ghost cell"""
    cfg, orc, gpu = make_amr(gpu_ctx, 3)
    for l in (1, 2):
        dxf = np.ascontiguousarray(orc.dx[l], dtype=np.float64)
        dxc = 2.0 * dxf
        ob.lib().orc_homogeneous_cf_interp(orc.F[l]["head"].h, dxf.ctypes.data_as(C.POINTER(C.c_double)), dxc.ctypes.data_as(C.POINTER(C.c_double)))
        gop = gpu.factory.AMRnewOp(l)
        gop.homogeneousCFInterp(gpu.F[l]["head"])
        same(gpu.F[l]["head"], orc.F[l]["head"], f"homogeneousCFInterp L{l}", ghosts=True)
    # base level: no coarse-fine region, a no-op
    gpu.factory.AMRnewOp(0).homogeneousCFInterp(gpu.F[0]["head"])
    same(gpu.F[0]["head"], orc.F[0]["head"], "homogeneousCFInterp L0")

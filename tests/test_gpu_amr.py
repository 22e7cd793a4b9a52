"""GPU parity of the AMR (multi-level) operator surface against the CPU oracle: QuadCFInterp, flux register / reflux,
AMROperator / AMRResidual, AMRRestrictS, AMRProlongS[_2], UpdateOperator with a coarser level, relaxNF, AMRNorm, and the
multi-level FAS V-cycle.  Bit-exact (np.array_equal) everywhere; the north-star tolerance 1e-10 relative L2 is asserted too."""
import numpy as np
import pytest

from oracle import binding as ob
from tests.problem import AmrGpuSide, AmrOracleSide, amr_hierarchy, fabs_equal, fields_equal, rel_l2

pytestmark = pytest.mark.gpu


HIERS = ["C5", "C4", "C5_256", "C5_BR"]  # tests/problem.py:amr_hierarchy -- 64^2 3-level, valley 2-level with ice mask < 0, 256^2 3-level,
# 64^2 3-level with the Berger-Rigoutsos shapes a regrid leaves (narrow boxes, refined boxes along the physical boundary on both levels)


def make(ctx, nlev=None, hier="C5", **kw):
    cfg, lv = amr_hierarchy(hier)
    nlev = nlev or len(lv)
    orc = AmrOracleSide(cfg, lv[:nlev], **kw)
    orc.average_down("head")
    orc.init_bcoef()
    gpu = AmrGpuSide(ctx, orc)
    return cfg, orc, gpu


def same(gpu_ld, orc_f, what, ghosts=False):
    if ghosts:
        d, eq = fabs_equal(gpu_ld, orc_f)
    else:
        d, eq = fields_equal(gpu_ld, orc_f)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


def test_general_layout_roundtrip_and_exchange(gpu_ctx):
    cfg, orc, gpu = make(gpu_ctx)
    for l in range(3):
        for k in ("head", "B", "rhs", "bX", "bY"):
            # level 0 is stored as one merged rectangle: ghost cells between its boxes ARE the neighbours' valid cells
            same(gpu.F[l][k], orc.F[l][k], f"roundtrip L{l} {k}", ghosts=l > 0)
        # exchange: face strips only is what the operator uses; the exported call fills corners too
        ob.lib().orc_exchange_full(orc.F[l]["head"].h)
        gpu.F[l]["head"].exchange(True)
        # (on level 0 a box's corner ghost outside the domain is another box's face ghost in the merged storage: skip ghosts)
        same(gpu.F[l]["head"], orc.F[l]["head"], f"exchange L{l}", ghosts=l > 0)


@pytest.mark.parametrize("hier", HIERS)
def test_cf_interp_bit_exact(gpu_ctx, hier):
    cfg, orc, gpu = make(gpu_ctx, hier=hier)
    for l in range(1, orc.nlev):
        ob.cf_interp(orc.F[l]["head"], orc.F[l - 1]["head"], 2, orc.dx[l][0])
        gop = gpu.factory.AMRnewOp(l)
        gop.coarseFineInterp(gpu.F[l]["head"], gpu.F[l - 1]["head"])
        same(gpu.F[l]["head"], orc.F[l]["head"], f"QuadCFInterp L{l}", ghosts=True)


@pytest.mark.parametrize("hier", HIERS)
def test_amr_operator_and_reflux_bit_exact(gpu_ctx, hier):
    cfg, orc, gpu = make(gpu_ctx, hier=hier)
    nl = orc.nlev
    oops = [orc.level_op(l) for l in range(nl)]
    gops = [gpu.factory.AMRnewOp(l) for l in range(nl)]
    for l in range(nl):
        olof, glof = ob.Field(orc.layouts[l], 1, 0), gpu.new_like(l, "rhs")
        fine = l + 1 if l < nl - 1 else None
        crse = l - 1 if l > 0 else None
        oops[l].amr_operator(olof, orc.F[fine]["head"] if fine else None, orc.F[l]["head"], orc.F[crse]["head"] if crse is not None else None,
                             False, oops[fine] if fine else None)
        gops[l].AMROperator(glof, gpu.F[fine]["head"] if fine else None, gpu.F[l]["head"], gpu.F[crse]["head"] if crse is not None else None,
                            False, gops[fine] if fine else None)
        same(glof, olof, f"AMROperator L{l}")
        oops[l].amr_residual(olof, orc.F[fine]["head"] if fine else None, orc.F[l]["head"], orc.F[crse]["head"] if crse is not None else None,
                             orc.F[l]["rhs"], False, oops[fine] if fine else None)
        gops[l].AMRResidual(glof, gpu.F[fine]["head"] if fine else None, gpu.F[l]["head"], gpu.F[crse]["head"] if crse is not None else None,
                            gpu.F[l]["rhs"], False, gops[fine] if fine else None)
        same(glof, olof, f"AMRResidual L{l}")
        if fine:
            # AMRNorm: zero under the finer level, then the norm
            assert gops[l].AMRNorm(glof, gpu.F[fine]["rhs"], 2, 0) == ob.amr_norm(olof, orc.layouts[fine], 2, 0)
            assert gops[l].AMRNorm(glof, gpu.F[fine]["rhs"], 2, 2) == pytest.approx(ob.amr_norm(olof, orc.layouts[fine], 2, 2), rel=1e-13)
            ob.zero_covered(olof, orc.layouts[fine], 2)
            gops[l].zeroCovered(glof, gpu.F[fine]["rhs"])
            same(glof, olof, f"zeroCovered L{l}")


@pytest.mark.parametrize("hier", HIERS)
def test_restrict_prolong_bit_exact(gpu_ctx, hier):
    cfg, orc, gpu = make(gpu_ctx, hier=hier)
    oops = [orc.level_op(l) for l in range(orc.nlev)]
    gops = [gpu.factory.AMRnewOp(l) for l in range(orc.nlev)]
    for l in range(1, orc.nlev):
        clay = orc.layouts[l].coarsen(2)
        oresC, oscr = ob.Field(clay, 1, 1), ob.Field(orc.layouts[l], 1, 1)
        gresC, gscr = gops[l].createCoarsened(gpu.F[l]["head"]), gpu.new_like(l, "head")
        for skip in (True, False):
            src_o = orc.F[l]["head"] if skip else orc.F[l]["rhs"]
            src_g = gpu.F[l]["head"] if skip else gpu.F[l]["rhs"]
            oops[l].amr_restrict_s(oresC, src_o, orc.F[l]["head"], orc.F[l - 1]["head"], oscr, skip)
            gops[l].AMRRestrictS(gresC, src_g, gpu.F[l]["head"], gpu.F[l - 1]["head"], gscr, skip)
            same(gresC, oresC, f"AMRRestrictS L{l} skip_res={skip}")
            # copyTo onto the coarser level
            otmp, gtmp = ob.Field(orc.layouts[l - 1], 1, 0), gpu.new_like(l - 1, "rhs")
            ob.copy_to(otmp, oresC)
            gresC.copyTo(gtmp)
            same(gtmp, otmp, f"copyTo L{l}->L{l - 1}")
        ocorr, gcorr = ob.Field(orc.layouts[l - 1], 1, 1), gpu.new_like(l - 1, "head")
        rng = np.random.RandomState(l)
        dom = orc.layouts[l - 1].domain
        gl = rng.rand(dom[3] + 3, dom[2] + 3)
        ocorr.set_global(gl, (-1, -1))
        gcorr.set_global(gl, (-1, -1))
        oops[l].amr_prolong_s(orc.F[l]["head"], ocorr, oresC)
        gops[l].AMRProlongS(gpu.F[l]["head"], gcorr)
        same(gpu.F[l]["head"], orc.F[l]["head"], f"AMRProlongS L{l}")
        oresC.setval(0.0)   # the oracle's scratch starts from zeros like the operator's own scratch on the device
        oops[l].amr_prolong_s2(orc.F[l]["head"], ocorr, oresC, oops[l - 1])
        gops[l].AMRProlongS_2(gpu.F[l]["head"], gcorr, gops[l - 1])
        same(gpu.F[l]["head"], orc.F[l]["head"], f"AMRProlongS_2 L{l}")


@pytest.mark.parametrize("hier", HIERS)
def test_update_operator_and_relax_nf_bit_exact(gpu_ctx, hier):
    cfg, orc, gpu = make(gpu_ctx, hier=hier)
    oops = [orc.level_op(l) for l in range(orc.nlev)]
    gops = [gpu.factory.AMRnewOp(l) for l in range(orc.nlev)]
    if hier == "C4":  # the point of this hierarchy: masked cells inside the fine boxes and on both sides of the coarse-fine interface
        m1, m0 = orc.F[1]["mask"].get_global(), orc.F[0]["mask"].get_global()
        assert (m1[~np.isnan(m1)] < 0).any() and (m1[~np.isnan(m1)] > 0).any() and (m0 < 0).any() and cfg.use_mask_grad and cfg.cutOffBcoef
    for l in range(1, orc.nlev):
        oops[l].relax_nf(orc.F[l]["head"], orc.F[l - 1]["head"], orc.F[l]["rhs"], 2)
        gops[l].relaxNF(gpu.F[l]["head"], gpu.F[l - 1]["head"], gpu.F[l]["rhs"], 2)
        same(gpu.F[l]["head"], orc.F[l]["head"], f"relaxNF L{l}", ghosts=True)
        oops[l].update_operator_amr(orc.F[l]["head"], orc.F[l - 1]["head"], orc.F[l - 1]["mask"])
        gops[l].UpdateOperator(gpu.F[l]["head"], gpu.F[l - 1]["head"], l, 0, False)
        same(gpu.F[l]["bX"], orc.F[l]["bX"], f"UpdateOperator bX L{l}")
        same(gpu.F[l]["bY"], orc.F[l]["bY"], f"UpdateOperator bY L{l}")
        ores, gres = ob.Field(orc.layouts[l], 1, 0), gpu.new_like(l, "rhs")
        oops[l].residual_nf(ores, orc.F[l]["head"], orc.F[l - 1]["head"], orc.F[l]["rhs"])
        gops[l].residualNF(gres, gpu.F[l]["head"], gpu.F[l - 1]["head"], gpu.F[l]["rhs"], False)
        same(gres, ores, f"residualNF L{l}")


@pytest.mark.parametrize("mode", ["reference flow", "fused", "fused off by tuning key"])
@pytest.mark.parametrize("hier", HIERS)
def test_refined_level_smoother_variants(gpu_ctx, hier, mode):
    """levelGSRB on refined (patch-table) levels: the fused per-patch red+black kernel (default) and the reference's
    exchange-per-colour flow must both reproduce the oracle bit for bit, ghost cells included, for odd and even sweep counts
    (the fused sweep is out of place: an odd count leaves the field in the other buffer)."""
    cfg, orc, gpu = make(gpu_ctx, hier=hier)
    oops = [orc.level_op(l) for l in range(orc.nlev)]
    gops = [gpu.factory.AMRnewOp(l) for l in range(orc.nlev)]
    try:
        gpu_ctx.set_relax_mode(0 if mode == "reference flow" else 1)
        gpu_ctx.set_tuning(7, 1 if mode == "fused off by tuning key" else 0)
        for l in range(1, orc.nlev):
            for n in (1, 3, 4):
                oops[l].relax_nf(orc.F[l]["head"], orc.F[l - 1]["head"], orc.F[l]["rhs"], n)
                gops[l].relaxNF(gpu.F[l]["head"], gpu.F[l - 1]["head"], gpu.F[l]["rhs"], n)
                same(gpu.F[l]["head"], orc.F[l]["head"], f"relaxNF x{n} L{l} ({mode})", ghosts=True)
            # no coarse level handed in: the coarse-fine ghost cells keep their values through the sweeps
            oops[l].relax(orc.F[l]["head"], orc.F[l]["rhs"], 3)
            gops[l].relax(gpu.F[l]["head"], gpu.F[l]["rhs"], 3)
            same(gpu.F[l]["head"], orc.F[l]["head"], f"relax x3 L{l} ({mode})", ghosts=True)
    finally:
        gpu_ctx.set_relax_mode(1)
        gpu_ctx.set_tuning(7, 0)


@pytest.mark.parametrize("nlev,hier", [(2, "C5"), (3, "C5"), (2, "C4"), (2, "C5_256"), (3, "C5_256"), (2, "C5_BR"), (3, "C5_BR")])
def test_amr_fixed_vcycles_parity(gpu_ctx, nlev, hier):
    cfg, orc, gpu = make(gpu_ctx, nlev, hier)
    ncyc = 4
    sp = ob.make_solver_params(bottom=10, fixed_cycles=ncyc)
    osol = orc.solver()
    it, ohist = osol.solve(orc.fields("head"), orc.fields("rhs"), nlev - 1, sp)
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, nlev)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve(gpu.fields("head"), gpu.fields("rhs"), fixed_cycles=ncyc)
    assert git == it == ncyc
    for l in range(nlev):
        oh, gh = orc.F[l]["head"].get_global(), gpu.F[l]["head"].get_global()
        assert rel_l2(gh, oh) <= 1e-10, f"level {l}"
        assert np.array_equal(np.isnan(gh), np.isnan(oh)), f"level {l}: NaN pattern differs"
        m = ~np.isnan(oh)
        assert np.array_equal(gh[m], oh[m]), f"level {l}: head differs, max {np.abs(gh[m] - oh[m]).max():g}"
    assert np.array_equal(ghist, ohist), (ghist, ohist)
    assert stats.cell_updates == ncyc * osol.cell_updates(sp, nlev - 1)


def test_amr_solve_with_stop_test(gpu_ctx):
    cfg, orc, gpu = make(gpu_ctx, 3)
    sp = ob.make_solver_params(bottom=10, eps=1e-6, hang=1e-4, imin=5, iter_min=2, max_iter=30)
    it, ohist = orc.solver().solve(orc.fields("head"), orc.fields("rhs"), 2, sp)
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 3)
    mg.setSolverParameters(4, 4, 10, 1, 30, 1e-6, 1e-4, 1e-7)
    mg.params.imin, mg.params.iter_min = 5, 2
    git, ghist, stats = mg.solve(gpu.fields("head"), gpu.fields("rhs"))
    assert git == it
    assert np.array_equal(ghist, ohist)


def test_against_golden_fixture_three_levels(gpu_ctx):
    """the CUDA path against the committed fixture (tests/golden/amr_3lev_vcycles.npz): 3 levels, 3 FAS V-cycles"""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "amr_3lev_vcycles.npz"))
    cfg, orc, gpu = make(gpu_ctx, 3)
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 3)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, stats = mg.solve(gpu.fields("head"), gpu.fields("rhs"), fixed_cycles=3)
    assert np.array_equal(ghist, z["resnorm"])
    for l in range(3):
        assert np.array_equal(np.nan_to_num(gpu.F[l]["head"].get_global(), nan=0.0), z[f"head3_L{l}"]), f"level {l}"


@pytest.mark.parametrize("hier,nlev", [("C5", 3), ("C4", 2), ("C5_256", 3), ("C5_BR", 3)])
def test_composite_sweeps_fused_vs_reference_flow(gpu_ctx, hier, nlev):
    """The V-cycle driver's shortcuts on the base level under a finer one -- (rhs - L phi) + L phi in one sweep with the cells next to /
    under the finer level redone by sparse kernels, norm-only residual sweeps, the FAS reference state kept on the coarsened-fine
    layout -- against the reference's own sequence (tune key 9 = 1: AMROperator, reflux, axby, AMROperatorNF, incr, whole-level
    axby for the correction).  Both must reproduce the oracle bit for bit, residual history included."""
    cfg, orc, gpu = make(gpu_ctx, nlev, hier)
    ncyc = 3
    sp = ob.make_solver_params(bottom=10, fixed_cycles=ncyc)
    it, ohist = orc.solver().solve(orc.fields("head"), orc.fields("rhs"), nlev - 1, sp)
    keep = {k: [f.download() for f in gpu.fields(k)] for k in ("head", "bX", "bY")}   # what a solve changes
    for key9 in (0, 1):
        for k, fabs in keep.items():
            for l, f in enumerate(gpu.fields(k)):
                f.upload(fabs[l])
        gpu_ctx.set_tuning(9, key9)
        try:
            mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, nlev)
            mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
            git, ghist, stats = mg.solve(gpu.fields("head"), gpu.fields("rhs"), fixed_cycles=ncyc)
        finally:
            gpu_ctx.set_tuning(9, 0)
        assert np.array_equal(ghist, ohist), (key9, ghist, ohist)
        for l in range(nlev):
            same(gpu.F[l]["head"], orc.F[l]["head"], f"head L{l} (tune 9 = {key9})")
        mg.destroy()

"""Multi-level Picard body and regrid transfers on the device against the oracle's independent restatement (oracle/picard_amr.py,
oracle/suhmo_oracle_r3.inc): PiecewiseLinearFillPatch, FineInterp, CoarseAverage, destructiveRegrid, aCoeff_bCoeff, the moulin
recharge, and whole Picard iterations + the explicit gap update on 2- and 3-level hierarchies.  Bit-exact except where exp() and the
order of a sum enter (moulin quadrature: 1e-12 relative)."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob
from oracle import picard_amr as opa
from tests.amr_picard import build_device, build_oracle
from tests.problem import amr_hierarchy, fabs_equal, fields_equal

pytestmark = pytest.mark.gpu


def same(gpu_ld, orc_f, what, ghosts=False):
    d, eq = fabs_equal(gpu_ld, orc_f) if ghosts else fields_equal(gpu_ld, orc_f)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


@pytest.mark.parametrize("hier", ["C5", "C4", "C5_256", "C5_BR"])
def test_interlevel_transfers_bit_exact(gpu_ctx, hier):
    cfg, lv = amr_hierarchy(hier)
    H, st = build_oracle(cfg, lv), build_device(gpu_ctx, cfg, lv)
    L = ob.lib()
    for l in range(1, H.nlev):
        for k in ("head", "B"):
            L.orc_pwl_fill_patch(H.S[l][k].h, H.S[l - 1][k].h, 2)
            st.ops[l].pwlFillPatch(st.S[l][k], st.S[l - 1][k])
            same(st.S[l][k], H.S[l][k], f"PiecewiseLinearFillPatch {k} L{l}", ghosts=True)
        # two components at once (qw, Dcoef_cc in the reference)
        L.orc_pwl_fill_patch(H.S[l]["gradH"].h, H.S[l - 1]["gradH"].h, 2)
        st.ops[l].pwlFillPatch(st.S[l]["gradH"], st.S[l - 1]["gradH"])
        same(st.S[l]["gradH"], H.S[l]["gradH"], f"PiecewiseLinearFillPatch 2 comps L{l}", ghosts=True)
        L.orc_fine_interp(H.S[l]["Re"].h, H.S[l - 1]["head"].h, 2)
        st.ops[l].fineInterp(st.S[l]["Re"], st.S[l - 1]["head"])
        same(st.S[l]["Re"], H.S[l]["Re"], f"FineInterp L{l}")
    for l in range(H.nlev - 1, 0, -1):
        tmp = ob.Field(H.layouts[l].coarsen(2), 1, 0)
        L.orc_coarse_average(H.S[l]["head"].h, tmp.h, 2)
        ob.copy_to(H.S[l - 1]["head"], tmp)
        st.ops[l].averageToCoarse(st.S[l - 1]["head"], st.S[l]["head"])
        same(st.S[l - 1]["head"], H.S[l - 1]["head"], f"CoarseAverage L{l}->L{l - 1}")


def test_regrid_transfer_bit_exact(gpu_ctx):
    """destructiveRegrid: level 1 moves from two of its boxes to the full set; old data survive where they were"""
    from suhmo_b200 import amr
    cfg, lv = amr_hierarchy("C5")
    H, st = build_oracle(cfg, lv[:2]), build_device(gpu_ctx, cfg, lv[:2])
    old_boxes = lv[1][:2]
    dom1 = (0, 0, cfg.nx * 2 - 1, cfg.ny * 2 - 1)
    olay, glay = ob.Layout(old_boxes, dom1, cfg.periodic), amr.DisjointBoxLayout(gpu_ctx, old_boxes, dom1, cfg.periodic)
    oold, gold = ob.Field(olay, 1, 1), amr.LevelData(glay, 1, 1, 0)
    rng = np.random.RandomState(3)
    g = rng.rand(cfg.ny * 2 + 2, cfg.nx * 2 + 2)
    oold.set_global(g, (-1, -1))
    gold.set_global(g, (-1, -1))
    ob.lib().orc_regrid_transfer(H.S[1]["head"].h, oold.h, H.S[0]["head"].h, 2)
    st.ops[1].regridTransfer(st.S[1]["head"], gold, st.S[0]["head"])
    same(st.S[1]["head"], H.S[1]["head"], "destructiveRegrid", ghosts=True)
    ob.lib().orc_regrid_transfer(H.S[1]["B"].h, None, H.S[0]["B"].h, 2)
    st.ops[1].regridTransfer(st.S[1]["B"], None, st.S[0]["B"])
    same(st.S[1]["B"], H.S[1]["B"], "destructiveRegrid without old data", ghosts=True)


def test_moulin_recharge_to_tolerance(gpu_ctx):
    from suhmo_b200.timestep_amr import AmrTimeStep
    cfg, lv = amr_hierarchy("C5")
    cfg.moulins = [(30000.0, 40000.0, 80.0, 3000.0), (70000.0, 20000.0, 40.0, 2500.0), (52000.0, 51000.0, 10.0, 1500.0)]
    H, st = build_oracle(cfg, lv), build_device(gpu_ctx, cfg, lv)
    ots, gts = opa.TimeStep(H), AmrTimeStep(st)
    ots.begin_step()
    gts.begin_step()
    oi = ots.moulin_sources(time=7200.0, runoff=0.3)
    gi = gts.moulin_sources(time=7200.0, runoff=0.3)
    assert np.allclose(gi, oi, rtol=1e-12, atol=0)
    for l in range(H.nlev):
        o, g = H.S[l]["MS"].get_global(), st.S[l]["MS"].get_global()
        assert np.array_equal(np.isnan(o), np.isnan(g))
        m = ~np.isnan(o)
        assert np.allclose(g[m], o[m], rtol=1e-11, atol=1e-300), f"level {l}: {np.abs(g[m] - o[m]).max()}"
        assert o[m].max() > 0


@pytest.mark.parametrize("hier", ["C5", "C4", "C5_BR"])
def test_multilevel_picard_and_gap_update_bit_exact(gpu_ctx, hier):
    from suhmo_b200.timestep_amr import AmrTimeStep
    cfg, lv = amr_hierarchy(hier)
    H, st = build_oracle(cfg, lv), build_device(gpu_ctx, cfg, lv)
    ots, gts = opa.TimeStep(H), AmrTimeStep(st)
    ots.begin_step()
    gts.begin_step()
    names = ("head", "B", "oldB", "oldH")
    for l in range(H.nlev):
        for k in names:
            same(st.S[l][k], H.S[l][k], f"begin_step {k} L{l}", ghosts=l > 0)   # level 0 is one merged rectangle on the device
    sp = ob.make_solver_params(bottom=10, fixed_cycles=3)
    for it in range(2):
        ots.picard_body()
        gts.picard_body()
        for l in range(H.nlev):
            for k in ("gradH", "Re", "qgh", "qgz", "mR", "Pw", "rhs", "Dterm"):
                same(st.S[l][k], H.S[l][k], f"Picard {it} body {k} L{l}")
            for k in ("Bec", "gH", "gZ", "Dc", "Reec", "Qw", "b"):
                for d in range(2):
                    same(st.S[l][k][d], H.S[l][k][d], f"Picard {it} body {k}[{d}] L{l}")
        n, ohist = ots.solver().solve(H.fields("head"), H.fields("rhs"), H.nlev - 1, sp)
        ghist = gts.solve_head(fixed_cycles=3)
        assert np.array_equal(ghist, ohist), (ghist, ohist)
        ots.after_solve()
        gts.after_solve()
        for l in range(H.nlev):
            same(st.S[l]["head"], H.S[l]["head"], f"Picard {it} head L{l}", ghosts=l > 0)
        assert gts.picard_change() == ots.picard_change()
    ots.update_gap(3600.0)
    gts.update_gap(3600.0)
    for l in range(H.nlev):
        for k in ("B", "mR", "Re", "RHSb"):
            same(st.S[l][k], H.S[l][k], f"gap update {k} L{l}")
        same(st.S[l]["B"], H.S[l]["B"], f"gap update B with ghosts L{l}", ghosts=l > 0)


@pytest.mark.parametrize("hier,cur_step", [("C5", 0), ("C5", 60), ("C4", 10)])
def test_multilevel_time_steps_same_decisions(gpu_ctx, hier, cur_step):
    """two whole time steps with convergence-driven head solves and Picard loops: the same number of V-cycles per solve, the same
    Picard iteration count, the same convergence measures, and bit-identical head and gap height on every level"""
    from suhmo_b200.timestep_amr import AmrTimeStep
    cfg, lv = amr_hierarchy(hier)
    H, st = build_oracle(cfg, lv), build_device(gpu_ctx, cfg, lv)
    ots, gts = opa.TimeStep(H), AmrTimeStep(st)
    for step in range(2):
        o = ots.time_step(1800.0, cur_step + step)
        g = gts.time_step(1800.0, cur_step + step)
        assert g["picard_iterations"] == o["picard_iterations"] and g["head_cycles"] == o["head_cycles"], (g, o)
        assert g["x_h"] == o["x_h"], (g["x_h"], o["x_h"])
        for l in range(H.nlev):
            same(st.S[l]["head"], H.S[l]["head"], f"step {step} head L{l}")
            same(st.S[l]["B"], H.S[l]["B"], f"step {step} gap height L{l}")

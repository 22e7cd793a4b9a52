"""The implicit gap-height equation of SolveForGap_nl (src/AmrHydro.cpp:594-662) on an AMR hierarchy, solved on the CPU oracle through
the FAS solver of the head equation with the nonlinear term switched off -- the route sg::AmrHydro::SolveForGap_FAS
(suhmo_b200/host/suhmo_amrhydro.hpp, experimental, off by default) takes on the device, where the reference's own multi-level linear
AMRMultiGrid is not restated.  What this shows: on ONE level the route lands on the stock linear solver's answer, and on THREE levels
it drives the composite residual of the same linear system to (numerically) zero, i.e. it solves the reference's discrete problem --
not that it reproduces the reference's iteration sequence.  CPU only; nothing here runs on a GPU."""
import ctypes as C

import numpy as np

from oracle import binding as ob
from oracle import picard_amr as opa
from suhmo_b200 import synthetic as syn
from tests.amr_picard import build_oracle
from tests.problem import amr_hierarchy

DT = 1800.0


def prepare(cfg, lv):
    """a hierarchy after one Picard iteration, with the right-hand side of the implicit gap equation on every level"""
    H = build_oracle(cfg, lv, use_ImplDiff=1)
    ts = opa.TimeStep(H)
    ts.begin_step()
    ts.picard_body()
    ts.solver().solve(H.fields("head"), H.fields("rhs"), H.nlev - 1, ob.make_solver_params(bottom=10, fixed_cycles=3))
    ts.after_solve()
    for l in range(H.nlev):
        S = H.S[l]
        ts.re_and_qw(l, True)
        ts.melt_rate(l)
        ob.lib().orc_rhs_gap(C.byref(H.q), S["RHSb"].h, S["Pi"].h, S["Pw"].h, S["mR"].h, S["B"].h, S["Dterm"].h, S["mask"].h, S["BH"].h, S["BL"].h,
                             S["MV"].h, DT)
    return H


def gap_through_fas(H, cur_step=3):
    ones = [ob.Field(lay, 1, 0) for lay in H.layouts]
    for o in ones:
        o.setval(1.0)
    f = H.fields
    sol = ob.AmrSolver(H.layouts, H.dx[0], 1.0, DT * H.q.DiffFactor, ob.make_bc((1, 1), (1, 1)), ob.make_params(use_NL=0, bcoeff_otf=0), ones,
                       [S["Dc"][0] for S in H.S], [S["Dc"][1] for S in H.S], f("B"), f("Pi"), f("zb"), f("mask"))
    cur = [ob.Field(lay, 1, 1) for lay in H.layouts]
    for c, S in zip(cur, H.S):
        c.copy_from(S["B"])
    sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10 if cur_step < 50 else 5, iter_min=2, eps=1e-7, hang=1e-6,
                               norm_thresh=1e-7)
    it, hist = sol.solve(cur, f("RHSb"), H.nlev - 1, sp)
    return cur, hist


def test_one_level_lands_on_the_linear_solver():
    cfg = syn.config("C2", 1)
    cfg.max_box_size = 32
    H = prepare(cfg, [syn.domain_split(cfg.nx, cfg.ny, 32, cfg.block_factor)])
    cur, hist = gap_through_fas(H)
    S = H.S[0]
    ones = ob.Field(H.layouts[0], 1, 0)
    ones.setval(1.0)
    lin = ob.Field(H.layouts[0], 1, 1)
    lin.copy_from(S["B"])
    s = ob.LinSolver(H.layouts[0], H.dx[0][0], 1.0, DT * H.q.DiffFactor, ones, S["Dc"][0], S["Dc"][1])
    it, hist_lin = s.solve(lin, S["RHSb"], ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10, iter_min=2, eps=1e-7, hang=1e-6,
                                                                 norm_thresh=1e-7))
    s.free()
    a, b = cur[0].get_global(), lin.get_global()
    assert hist[0] == hist_lin[0] > 0                       # the same initial residual: the same operator and right-hand side
    assert np.abs(a - b).max() <= 1e-12 * np.abs(b).max()
    assert hist[-1] <= 1e-7 * hist[0]


def test_three_levels_solve_the_composite_system():
    cfg, lv = amr_hierarchy("C5")
    H = prepare(cfg, lv)
    before = [S["B"].get_global().copy() for S in H.S]
    cur, hist = gap_through_fas(H)
    assert 2 <= len(hist) - 1 < 20 and hist[-1] <= 1e-7 * hist[0], hist
    for l in range(H.nlev):
        a = cur[l].get_global()
        m = ~np.isnan(a)
        assert np.isfinite(a[m]).all() and np.abs(a[m] - before[l][m]).max() > 0
    # with a = 1 and a diffusion term of relative size dt * DiffFactor * D / dx^2 << 1 the solution sits next to the right-hand side on
    # every cell that no finer level covers (covered cells carry the average of the finer solution)
    for l in range(H.nlev):
        a, r = cur[l].get_global(), H.S[l]["RHSb"].get_global()
        m = ~np.isnan(a)
        if l + 1 < H.nlev:
            for bx in lv[l + 1]:
                m[bx[1] // 2:bx[3] // 2 + 1, bx[0] // 2:bx[2] // 2 + 1] = False
        assert m.any() and np.abs(a[m] - r[m]).max() <= 1e-6 * np.abs(r[m]).max()

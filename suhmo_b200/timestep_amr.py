"""Time step of AmrHydro::timeStepFAS on an AMR hierarchy, every field resident on the device (src/AmrHydro.cpp:2255-3620).

Host orchestration over the C ABI, as in the reference (the time and Picard loops are host code there too): per Picard iteration the
ghost fills (PiecewiseLinearFillPatch, exchange, boundary conditions), centring changes, gradients with coarse-fine interpolation,
Re, water flux, melt rate, RHS_h, aCoeff_bCoeff, the composite FAS head solve, average-down; then the explicit gap-height update.
No field leaves the device between the head solves.  The single-level Picard step with the implicit gap solve lives in
suhmo_b200/timestep.py; the implicit solve on more than one level is not built, although reference inputs ask for it (exec/AMR_multiMoulins/run_C_*lev and
exec/0_convergence_channelized/*_base* set solver.use_ImplDiff = true with max_level > 0): see DESIGN.md section 9.
"""
import ctypes as C

import numpy as np

from . import amr, capi
from .timestep import picard_params

CELL, XFACE, YFACE = 0, 1, 2
_SPEC = dict(head=(1, 1), B=(1, 1), Pi=(1, 1), zb=(1, 1), mask=(1, 1), MV=(1, 1), BH=(1, 1), BL=(1, 1), mR=(1, 1), Pw=(1, 1), Re=(1, 1),
             MS=(1, 1), headLag=(1, 1), oldH=(1, 1), oldB=(1, 1), work=(1, 1), gradH=(2, 1), qgh=(2, 1), qgz=(2, 1), rhs=(1, 0), RHSb=(1, 0),
             Dterm=(1, 0), a=(1, 0))
_FACES = ("Bec", "mRec", "gH", "gZ", "Dc", "Reec", "Qw", "t1", "t2", "IMec", "b")


def _dx(dx):
    a = np.array(dx, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class AmrState:
    """device fields of every level, the operator factory on them (bCoef = the `b` face pair) and the per-level operators"""

    def __init__(self, ctx, cfg, layouts, prm=None, bc=None, moulins=None, **picard_over):
        self.ctx, self.cfg, self.layouts = ctx, cfg, layouts
        self.nlev = len(layouts)
        self.dx = [(cfg.dx[0] / 2 ** l, cfg.dx[1] / 2 ** l) for l in range(self.nlev)]
        self.bc = bc if bc is not None else amr.make_bc(cfg.bc_lo, cfg.bc_hi, cfg.bc_lo_val, cfg.bc_hi_val)
        self.prm = prm if prm is not None else amr.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr,
                                                               cutOffBcoef=cfg.cutOffBcoef, use_mask_grad=cfg.use_mask_grad)
        self.q = picard_params(capi.PicardParams, cfg, **picard_over)
        self.moulins = moulins if moulins is not None else cfg.moulins
        self.S = []
        for lay in layouts:
            S = {k: amr.LevelData(lay, nc, ng, CELL) for k, (nc, ng) in _SPEC.items()}
            for k in _FACES:
                S[k] = (amr.LevelData(lay, 1, 0, XFACE), amr.LevelData(lay, 1, 0, YFACE))
            self.S.append(S)
        f = self.fields
        self.factory = amr.VCAMRNonLinearPoissonOpFactory().define(
            ctx, layouts, [2] * (self.nlev - 1), cfg.dx, self.bc, 0.0, f("a"), -1.0, [S["b"][0] for S in self.S], [S["b"][1] for S in self.S],
            self.prm, f("B"), f("Pi"), f("zb"), f("mask"))
        self.ops = [self.factory.AMRnewOp(l) for l in range(self.nlev)]
        self.mg = None

    def fields(self, name):
        return [S[name] for S in self.S]


class AmrTimeStep:
    def __init__(self, state):
        self.st, self.L = state, capi.lib()

    def _ck(self, status):
        capi.check(status)

    # ---- ghost cells -----------------------------------------------------------------------------------------------------
    def fill_cf_linear(self, l, names):
        """PiecewiseLinearFillPatch::fillInterp of the named fields on level l > 0"""
        st = self.st
        for k in names:
            st.ops[l].pwlFillPatch(st.S[l][k], st.S[l - 1][k])

    def fill_cf_quadratic(self, l, name):
        st = self.st
        st.ops[l].coarseFineInterp(st.S[l][name], st.S[l - 1][name])

    def head_bc(self, l):
        st = self.st
        self._ck(self.L.sg_apply_bc(st.S[l]["head"].h, C.byref(st.bc), _dx(st.dx[l])[1], 0))

    def average_down(self, name):
        st = self.st
        for l in range(st.nlev - 1, 0, -1):
            st.ops[l].averageToCoarse(st.S[l - 1][name], st.S[l][name])

    def copy(self, l, dst, src):
        self.st.ops[l].assignLocal(self.st.S[l][dst], self.st.S[l][src])

    # ---- pieces ------------------------------------------------------------------------------------------------------------
    def begin_step(self):
        """src/AmrHydro.cpp:2356-2445: consistent head and gap height (ghost cells, boundary conditions), old copies"""
        st, L = self.st, self.L
        for l in range(st.nlev):
            S = st.S[l]
            if l > 0:
                self.fill_cf_linear(l, ("head", "B"))
            S["head"].exchange(True)
            S["B"].exchange(True)
            amr.CopyGhostCells(S["B"])
            self.head_bc(l)
            self.copy(l, "oldH", "head")
            self.copy(l, "oldB", "B")
            self._ck(L.sg_icemask_ec(S["mask"].h, S["IMec"][0].h, S["IMec"][1].h))

    def grad_head(self, l):
        """compute_grad_head (src/AmrHydro.cpp:1611-1656)"""
        st, L, S = self.st, self.L, self.st.S[l]
        if l > 0:
            self.fill_cf_quadratic(l, "head")
        mask = S["mask"].h if st.cfg.use_mask_grad else None
        self._ck(L.sg_mac_gradient(S["head"].h, mask, _dx(st.dx[l])[1], S["gH"][0].h, S["gH"][1].h))
        self._ck(L.sg_edge_to_cell(S["gH"][0].h, S["gH"][1].h, S["gradH"].h))
        if l > 0:
            self.fill_cf_quadratic(l, "gradH")
        S["gradH"].exchange(True)
        amr.ExtrapGhostCells(S["gradH"])

    def grad_zb(self, l):
        """compute_grad_zb_ec (src/AmrHydro.cpp:1578-1608)"""
        st, L, S = self.st, self.L, self.st.S[l]
        if l > 0:
            self.fill_cf_quadratic(l, "zb")
        mask = S["mask"].h if st.cfg.use_mask_grad else None
        self._ck(L.sg_mac_gradient(S["zb"].h, mask, _dx(st.dx[l])[1], S["gZ"][0].h, S["gZ"][1].h))

    def reynolds_and_flux(self, l, fresh_gradient):
        """evaluate_Re_quadratic, fill, CellToEdge, evaluate_Qw_ec (src/AmrHydro.cpp:2703-2760; :3256-3290 with a fresh gradient)"""
        st, L, S = self.st, self.L, self.st.S[l]
        if fresh_gradient:
            self.grad_head(l)
        self._ck(L.sg_compute_re(C.byref(st.prm), S["Re"].h, S["B"].h, S["gradH"].h))
        if l > 0:
            self.fill_cf_linear(l, ("Re",))
        S["Re"].exchange(True)
        self._ck(L.sg_cell_to_edge(S["Re"].h, S["Reec"][0].h, S["Reec"][1].h))
        for d in range(2):
            self._ck(L.sg_compute_qw(C.byref(st.prm), S["Bec"][d].h, S["Reec"][d].h, S["gH"][d].h, S["Qw"][d].h))

    def melting(self, l):
        st, L, S = self.st, self.L, self.st.S[l]
        for d in range(2):
            self._ck(L.sg_compute_scaprod(S["Qw"][d].h, S["gH"][d].h, S["gZ"][d].h, S["t1"][d].h, S["t2"][d].h))
        self._ck(L.sg_edge_to_cell(S["t1"][0].h, S["t1"][1].h, S["qgh"].h))
        self._ck(L.sg_edge_to_cell(S["t2"][0].h, S["t2"][1].h, S["qgz"].h))
        self._ck(L.sg_calc_melting_rate(C.byref(st.q), S["head"].h, S["zb"].h, S["Pi"].h, S["mask"].h, S["B"].h, S["qgh"].h, S["qgz"].h,
                                        S["Pw"].h, S["mR"].h))

    def moulin_sources(self, time=0.0, runoff=0.0):
        """src/AmrHydro.cpp:2800-2836 (the branch taken after a regrid): quadrature, normalisation, average-down, ghost cells"""
        st = self.st
        integ = amr.moulin_source_terms(st.ops, st.fields("MS"), st.moulins, runoff, time)
        self.average_down("MS")
        for l in range(st.nlev):
            if l > 0:
                self.fill_cf_quadratic(l, "MS")
            st.S[l]["MS"].exchange(True)
            amr.ExtrapGhostCells(st.S[l]["MS"])
        return integ

    def picard_body(self):
        """src/AmrHydro.cpp:2477-3105: everything of one Picard iteration before the head solve"""
        st, L = self.st, self.L
        for l in range(st.nlev):
            S = st.S[l]
            if l > 0:
                self.fill_cf_linear(l, ("head", "B", "mR"))
            for k in ("head", "B", "mR"):
                S[k].exchange(True)
            amr.CopyGhostCells(S["B"])
            self.head_bc(l)
            self.copy(l, "headLag", "head")
            amr.ExtrapGhostCells(S["mR"])
            self._ck(L.sg_cell_to_edge(S["B"].h, S["Bec"][0].h, S["Bec"][1].h))
            self._ck(L.sg_cell_to_edge(S["mR"].h, S["mRec"][0].h, S["mRec"][1].h))
        for l in range(st.nlev):
            S = st.S[l]
            self.grad_head(l)
            self.grad_zb(l)
            for d in range(2):
                self._ck(L.sg_compute_dcoeff(S["Dc"][d].h, S["mRec"][d].h, S["Bec"][d].h, S["IMec"][d].h, st.q.rho_i, st.cfg.cutOffBcoef))
        for l in range(st.nlev):
            self.reynolds_and_flux(l, False)
        for l in range(st.nlev):
            S = st.S[l]
            self.melting_inputs_and_rhs(l)
        for l in range(st.nlev):
            S = st.S[l]
            st.ops[l].setToZero(S["a"])
            for d in range(2):
                self._ck(L.sg_compute_bcoeff(C.byref(st.prm), S["Bec"][d].h, S["Reec"][d].h, S["IMec"][d].h, S["b"][d].h))

    def melting_inputs_and_rhs(self, l):
        """src/AmrHydro.cpp:2920-3079 on one level: q.grad(h), q.grad(zb), diffusive term, melt rate, RHS_h"""
        st, L, S = self.st, self.L, self.st.S[l]
        for d in range(2):
            self._ck(L.sg_compute_scaprod(S["Qw"][d].h, S["gH"][d].h, S["gZ"][d].h, S["t1"][d].h, S["t2"][d].h))
        self._ck(L.sg_edge_to_cell(S["t1"][0].h, S["t1"][1].h, S["qgh"].h))
        self._ck(L.sg_edge_to_cell(S["t2"][0].h, S["t2"][1].h, S["qgz"].h))
        self._ck(L.sg_compute_difterm(S["B"].h, _dx(st.dx[l])[1], S["Dterm"].h, S["Dc"][0].h, S["Dc"][1].h))
        self._ck(L.sg_calc_melting_rate(C.byref(st.q), S["head"].h, S["zb"].h, S["Pi"].h, S["mask"].h, S["B"].h, S["qgh"].h, S["qgz"].h,
                                        S["Pw"].h, S["mR"].h))
        self._ck(L.sg_rhs_head(C.byref(st.q), S["rhs"].h, S["mR"].h, S["B"].h, S["BH"].h, S["BL"].h, S["MV"].h, S["MS"].h, S["Dterm"].h, S["mask"].h))

    def solve_head(self, fixed_cycles=0, cur_step=0, bottom=None):
        """SolveForHead_nl (src/AmrHydro.cpp:666-769): the reference's constants unless a fixed number of V-cycles is asked for"""
        st = self.st
        if st.mg is None:
            st.mg = amr.AMRFASMultiGrid().define(st.factory, st.nlev)
        else:
            st.mg.refresh()
        early = cur_step < 50
        if fixed_cycles:
            st.mg.setSolverParameters(4, 4, bottom or 10, 1, 100, 1e-10, 1e-4, 1e-7)
        else:
            st.mg.setSolverParameters(4, 4, bottom or (10 if early else 16), 1, 100, 1e-10 if early else 1e-7, 1e-4 if early else 0.01, 1e-7)
            st.mg.params.imin, st.mg.params.iter_min = (20 if early else 5), 2
        it, hist, stats = st.mg.solve(st.fields("head"), st.fields("rhs"), fixed_cycles=fixed_cycles)
        return hist

    def after_solve(self):
        """src/AmrHydro.cpp:3134-3165: average the head down, refill its ghost cells"""
        st = self.st
        self.average_down("head")
        for l in range(st.nlev):
            if l > 0:
                self.fill_cf_linear(l, ("head",))
            st.S[l]["head"].exchange(True)
            self.head_bc(l)

    def picard_change(self):
        """max |h_lag - h| / max h over the hierarchy (src/AmrHydro.cpp:3168-3185); head is positive in every SUHMO set-up"""
        st = self.st
        mx = max(st.ops[l].norm(st.S[l]["head"], 0) for l in range(st.nlev))
        res = 0.0
        for l in range(st.nlev):
            w = st.S[l]["work"]
            st.ops[l].axby(w, st.S[l]["headLag"], st.S[l]["head"], 1.0, -1.0)
            if l + 1 < st.nlev:
                st.ops[l].zeroCovered(w, st.S[l + 1]["head"])   # computeMax looks at the cells no finer level covers
            res = max(res, st.ops[l].norm(w, 0) / mx)
        return res

    def picard_iteration(self, fixed_cycles=0, cur_step=0, bottom=None):
        self.picard_body()
        hist = self.solve_head(fixed_cycles, cur_step, bottom)
        self.after_solve()
        return hist

    def update_gap(self, dt):
        """explicit gap-height update (src/AmrHydro.cpp:3248-3423, 3590-3595)"""
        st, L = self.st, self.L
        for l in range(st.nlev):
            S = st.S[l]
            self.reynolds_and_flux(l, True)
            self.melting(l)
            self._ck(L.sg_rhs_gap(C.byref(st.q), S["RHSb"].h, S["Pi"].h, S["Pw"].h, S["mR"].h, S["B"].h, S["Dterm"].h, S["mask"].h, S["BH"].h,
                                  S["BL"].h, S["MV"].h, dt))
            self._ck(L.sg_gap_euler(S["B"].h, S["oldB"].h, S["RHSb"].h, dt))
            if l > 0:
                self.fill_cf_linear(l, ("B",))
            S["B"].exchange(True)
            amr.CopyGhostCells(S["B"])
        self.average_down("B")

    def time_step(self, dt, cur_step=0, eps_picard=1.0e-6, max_picard=100):
        """one time step as the reference runs it: Picard iterations until the lagged change of head passes the reference's test
        (src/AmrHydro.cpp:3187-3229), then the gap update.  Returns {"picard_iterations", "x_h", "head_cycles"}."""
        self.begin_step()
        out = {"x_h": [], "head_cycles": []}
        ite = 0
        while True:
            hist = self.picard_iteration(0, cur_step)
            out["head_cycles"].append(len(hist) - 1)
            x_h = self.picard_change()
            out["x_h"].append(x_h)
            if ite > max_picard:
                raise RuntimeError("does not converge (Picard iterations > 100)")
            done = (x_h < 0.05 and ite > 2) if cur_step < 2 else (x_h < 0.05) if cur_step < 50 else (x_h < eps_picard)
            ite += 1
            if done:
                break
        out["picard_iterations"] = ite
        self.update_gap(dt)
        return out

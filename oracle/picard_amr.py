"""AmrHydro::timeStepFAS on an AMR hierarchy, restated for the CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle/suhmo_oracle.h).

Written from src/AmrHydro.cpp:2255-3620 top to bottom, one statement of the reference per call into the oracle library, for any
number of levels (one level included).  It deliberately shares no code with the product's orchestration (suhmo_b200/timestep.py):
the parity tests run the two against each other field by field.  Only the explicit gap-height update is restated for more than
one level (the implicit solve of SolveForGap_nl on an AMR hierarchy -- asked for by exec/AMR_multiMoulins/run_C_*lev -- is not restated).

State: per level l a dict S[l] of oracle fields
  persistent   head B Pi zb mask MV BH BL mR Pw Re MS gradH(2 comps)             (1 ghost cell)
  per step     rhs RHSb Dterm a (0 ghost), headLag oldH oldB (1 ghost), qgh qgz (2 comps, 1 ghost)
  faces (x, y) Bec mRec gH gZ Dc Reec Qw t1 t2 IMec bX bY
"""
import ctypes as C

import numpy as np

from . import binding as ob

CELL, XFACE, YFACE = 0, 1, 2
CELL_1G = ("head", "B", "Pi", "zb", "mask", "MV", "BH", "BL", "mR", "Pw", "Re", "MS", "headLag", "oldH", "oldB")
CELL_2C = ("gradH", "qgh", "qgz")
CELL_0G = ("rhs", "RHSb", "Dterm", "a")
FACES = ("Bec", "mRec", "gH", "gZ", "Dc", "Reec", "Qw", "t1", "t2", "IMec")


def _dxp(dx):
    a = np.array(dx, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class Hierarchy:
    """fields of every level + the constants of the run"""

    def __init__(self, cfg, layouts, dx, prm, bc, q, moulins=None):
        self.cfg, self.layouts, self.dx, self.prm, self.bc, self.q = cfg, layouts, dx, prm, bc, q
        self.nlev = len(layouts)
        self.moulins = moulins
        self.S = []
        for lay in layouts:
            S = {k: ob.Field(lay, 1, 1) for k in CELL_1G}
            S.update({k: ob.Field(lay, 2, 1) for k in CELL_2C})
            S.update({k: ob.Field(lay, 1, 0) for k in CELL_0G})
            for k in FACES + ("b",):
                S[k] = (ob.Field(lay, 1, 0, XFACE), ob.Field(lay, 1, 0, YFACE))
            self.S.append(S)

    def fields(self, name):
        return [S[name] for S in self.S]


class TimeStep:
    def __init__(self, H):
        self.H, self.L = H, ob.lib()
        self.use_mask = bool(H.cfg.use_mask_grad)

    # ---- the Chombo calls the reference makes, on oracle fields -------------------------------------------------------
    def fill_interp(self, l, name):            # PiecewiseLinearFillPatch::fillInterp, :2373-2380 and friends
        self.L.orc_pwl_fill_patch(self.H.S[l][name].h, self.H.S[l - 1][name].h, 2)

    def quad_cfi(self, l, name):               # QuadCFInterp::coarseFineInterp
        self.L.orc_cf_interp(self.H.S[l][name].h, self.H.S[l - 1][name].h, 2, self.H.dx[l][0])

    def exchange(self, f): self.L.orc_exchange_full(f.h)
    def bc_head(self, l, f): self.L.orc_apply_bc(f.h, C.byref(self.H.bc), _dxp(self.H.dx[l])[1], 0)   # mixBCValues(..., false)

    def average_down(self, name):              # CoarseAverage::averageToCoarse, finest level first
        H = self.H
        for l in range(H.nlev - 1, 0, -1):
            tmp = ob.Field(H.layouts[l].coarsen(2), 1, 0)
            self.L.orc_coarse_average(H.S[l][name].h, tmp.h, 2)
            ob.copy_to(H.S[l - 1][name], tmp)

    # ---- I: before the Picard loop (:2356-2445) ----------------------------------------------------------------------
    def begin_step(self):
        H, L = self.H, self.L
        for l in range(H.nlev):
            S = H.S[l]
            if l > 0:
                self.fill_interp(l, "head")
                self.fill_interp(l, "B")
            self.exchange(S["head"])
            self.exchange(S["B"])
            L.orc_copy_ghost(S["B"].h)
            self.bc_head(l, S["head"])
            S["oldH"].copy_from(S["head"])
            S["oldB"].copy_from(S["B"])
            L.orc_icemask_ec(S["mask"].h, S["IMec"][0].h, S["IMec"][1].h)   # m_iceMask_ec is persistent state (set at init / regrid)

    # ---- gradients (:1611-1656, 1578-1608) -----------------------------------------------------------------------------
    def compute_grad_head(self, l):
        H, L, S = self.H, self.L, self.H.S[l]
        if l > 0:
            self.quad_cfi(l, "head")           # levelGradientMAC's coarse-fine BC, util/Gradient.cpp:85-93
        L.orc_mac_gradient(S["head"].h, S["mask"].h if self.use_mask else None, _dxp(H.dx[l])[1], S["gH"][0].h, S["gH"][1].h)
        L.orc_edge_to_cell(S["gH"][0].h, S["gH"][1].h, S["gradH"].h)
        if l > 0:
            self.quad_cfi(l, "gradH")
        self.exchange(S["gradH"])
        L.orc_extrap_ghost(S["gradH"].h)

    def compute_grad_zb_ec(self, l):
        H, L, S = self.H, self.L, self.H.S[l]
        if l > 0:
            self.quad_cfi(l, "zb")
        L.orc_mac_gradient(S["zb"].h, S["mask"].h if self.use_mask else None, _dxp(H.dx[l])[1], S["gZ"][0].h, S["gZ"][1].h)

    def re_and_qw(self, l, compute_grad):       # evaluate_Re_quadratic + fill + CellToEdge + evaluate_Qw_ec (:2703-2760, 3256-3290)
        H, L, S = self.H, self.L, self.H.S[l]
        if compute_grad:
            self.compute_grad_head(l)
        L.orc_compute_re(C.byref(H.prm), S["B"].h, S["gradH"].h, S["Re"].h)
        if l > 0:
            self.fill_interp(l, "Re")
        self.exchange(S["Re"])
        L.orc_cell_to_edge(S["Re"].h, S["Reec"][0].h, S["Reec"][1].h)
        for d in range(2):
            L.orc_compute_qw(C.byref(H.prm), S["Bec"][d].h, S["Reec"][d].h, S["gH"][d].h, S["Qw"][d].h)

    def melt_rate(self, l):                      # COMPUTESCAPROD, EdgeToCell x2, Calc_meltingRate (:2964-2990 + :3022, 3295-3330)
        L, S, q = self.L, self.H.S[l], self.H.q
        for d in range(2):
            L.orc_compute_scaprod(S["Qw"][d].h, S["gH"][d].h, S["gZ"][d].h, S["t1"][d].h, S["t2"][d].h)
        L.orc_edge_to_cell(S["t1"][0].h, S["t1"][1].h, S["qgh"].h)
        L.orc_edge_to_cell(S["t2"][0].h, S["t2"][1].h, S["qgz"].h)
        L.orc_calc_melting_rate(C.byref(q), S["head"].h, S["zb"].h, S["Pi"].h, S["mask"].h, S["B"].h, S["qgh"].h, S["qgz"].h, S["Pw"].h, S["mR"].h)

    # ---- moulins (:2800-2836) --------------------------------------------------------------------------------------------
    def moulin_sources(self, time=0.0, runoff=0.0):
        H, L = self.H, self.L
        n = len(H.moulins)
        pos = np.ascontiguousarray([[m[0], m[1]] for m in H.moulins], dtype=np.float64).ravel()
        flux = np.ascontiguousarray([m[2] for m in H.moulins], dtype=np.float64)
        sig = np.ascontiguousarray([m[3] for m in H.moulins], dtype=np.float64)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
        integ = np.zeros(n)
        tmp = [ob.Field(lay, n, 0) for lay in H.layouts]
        for l in range(H.nlev - 1, -1, -1):
            L.orc_moulin_nonorm(tmp[l].h, _dxp(H.dx[l])[1], n, dp(pos), dp(sig))
            L.orc_moulin_integral(tmp[l].h, H.layouts[l + 1].h if l + 1 < H.nlev else None, _dxp(H.dx[l])[1], n, dp(integ))
        for l in range(H.nlev):
            L.orc_moulin_source(H.S[l]["MS"].h, tmp[l].h, n, dp(integ), dp(flux), float(runoff), float(time))
        self.average_down("MS")
        for l in range(H.nlev):
            if l > 0:
                self.quad_cfi(l, "MS")
            self.exchange(H.S[l]["MS"])
            L.orc_extrap_ghost(H.S[l]["MS"].h)
        return integ

    # ---- II: one pass of `while (!converged_h)` up to the head solve (:2477-3119) -------------------------------------------
    def picard_body(self):
        H, L = self.H, self.L
        for l in range(H.nlev):                                       # :2482-2532
            S = H.S[l]
            if l > 0:
                self.fill_interp(l, "head")
                self.fill_interp(l, "B")
                self.fill_interp(l, "mR")
            self.exchange(S["head"])
            self.exchange(S["B"])
            self.exchange(S["mR"])
            L.orc_copy_ghost(S["B"].h)
            self.bc_head(l, S["head"])
            S["headLag"].copy_from(S["head"])
            L.orc_extrap_ghost(S["mR"].h)
            L.orc_cell_to_edge(S["B"].h, S["Bec"][0].h, S["Bec"][1].h)
            L.orc_cell_to_edge(S["mR"].h, S["mRec"][0].h, S["mRec"][1].h)
        for l in range(H.nlev):                                       # :2539-2572
            S = H.S[l]
            self.compute_grad_head(l)
            self.compute_grad_zb_ec(l)
            for d in range(2):
                L.orc_compute_dcoeff(S["Dc"][d].h, S["mRec"][d].h, S["Bec"][d].h, S["IMec"][d].h, H.q.rho_i, H.cfg.cutOffBcoef)
        for l in range(H.nlev):                                       # :2703-2760
            self.re_and_qw(l, False)
        for l in range(H.nlev):                                       # :2920-3079
            S = H.S[l]
            for d in range(2):
                L.orc_compute_scaprod(S["Qw"][d].h, S["gH"][d].h, S["gZ"][d].h, S["t1"][d].h, S["t2"][d].h)
            L.orc_edge_to_cell(S["t1"][0].h, S["t1"][1].h, S["qgh"].h)
            L.orc_edge_to_cell(S["t2"][0].h, S["t2"][1].h, S["qgz"].h)
            L.orc_compute_difterm(S["B"].h, _dxp(H.dx[l])[1], S["Dterm"].h, S["Dc"][0].h, S["Dc"][1].h)
            L.orc_calc_melting_rate(C.byref(H.q), S["head"].h, S["zb"].h, S["Pi"].h, S["mask"].h, S["B"].h, S["qgh"].h, S["qgz"].h, S["Pw"].h, S["mR"].h)
            L.orc_rhs_head(C.byref(H.q), S["rhs"].h, S["mR"].h, S["B"].h, S["BH"].h, S["BL"].h, S["MV"].h, S["MS"].h, S["Dterm"].h, S["mask"].h)
        for l in range(H.nlev):                                       # aCoeff_bCoeff, :3087-3105
            S = H.S[l]
            S["a"].setval(0.0)
            for d in range(2):
                L.orc_compute_bcoeff(C.byref(H.prm), S["Bec"][d].h, S["Reec"][d].h, S["IMec"][d].h, S["b"][d].h)

    def solver(self):
        H, f = self.H, self.H.fields
        bX, bY = [S["b"][0] for S in H.S], [S["b"][1] for S in H.S]
        return ob.AmrSolver(H.layouts, H.dx[0], 0.0, -1.0, H.bc, H.prm, f("a"), bX, bY, f("B"), f("Pi"), f("zb"), f("mask"))

    def after_solve(self):                                            # :3134-3165
        H = self.H
        self.average_down("head")
        for l in range(H.nlev):
            if l > 0:
                self.fill_interp(l, "head")
            self.exchange(H.S[l]["head"])
            self.bc_head(l, H.S[l]["head"])

    def picard_change(self):                                          # :3168-3185: max|h_lag - h| / max h over the composite grid
        H, L = self.H, self.L
        mx = max(float(L.orc_norm(S["head"].h, 0)) for S in H.S)       # head is positive in every SUHMO set-up: max h = max |h|
        res = 0.0
        for l, S in enumerate(H.S):
            d = ob.Field(H.layouts[l], 1, 1)
            L.orc_axby(d.h, S["headLag"].h, S["head"].h, 1.0, -1.0)
            if l + 1 < H.nlev:
                ob.zero_covered(d, H.layouts[l + 1])                   # computeMax looks at the cells no finer level covers
            res = max(res, float(L.orc_norm(d.h, 0)) / mx)
        return res

    def picard_iteration(self, sp):
        """one Picard iteration with the head solve under solver parameters sp; returns the residual history"""
        self.picard_body()
        H = self.H
        it, hist = self.solver().solve(H.fields("head"), H.fields("rhs"), H.nlev - 1, sp)
        self.after_solve()
        return hist

    # ---- III: gap height, explicit (:3248-3423, 3590-3595) -------------------------------------------------------------------
    def update_gap(self, dt):
        H, L = self.H, self.L
        for l in range(H.nlev):
            S = H.S[l]
            self.re_and_qw(l, True)                                   # Re and Qw again with the fresh head
            self.melt_rate(l)
            L.orc_rhs_gap(C.byref(H.q), S["RHSb"].h, S["Pi"].h, S["Pw"].h, S["mR"].h, S["B"].h, S["Dterm"].h, S["mask"].h, S["BH"].h, S["BL"].h,
                          S["MV"].h, dt)
            L.orc_gap_euler(S["B"].h, S["oldB"].h, S["RHSb"].h, dt)
            if l > 0:
                self.fill_interp(l, "B")
            self.exchange(S["B"])
            L.orc_copy_ghost(S["B"].h)
        self.average_down("B")

    # ---- the whole step (:2255-3620): Picard iterations until the reference's test passes, then the gap update ------------------------
    def time_step(self, dt, cur_step=0, eps_picard=1.0e-6):
        early = cur_step < 50
        sp = ob.make_solver_params(pre=4, post=4, bottom=10 if early else 16, max_iter=100, imin=20 if early else 5, iter_min=2,
                                   eps=1e-10 if early else 1e-7, hang=1e-4 if early else 0.01, norm_thresh=1e-7)    # :737-762
        self.begin_step()
        out = {"x_h": [], "head_cycles": []}
        ite_idx = 0
        while True:
            hist = self.picard_iteration(sp)
            out["head_cycles"].append(len(hist) - 1)
            max_resH = self.picard_change()
            out["x_h"].append(max_resH)
            if ite_idx > 100:
                raise RuntimeError("Abort")                                # MayDay::Error("Abort"), :3190-3195
            if cur_step < 2:
                converged = max_resH < 0.05 and ite_idx > 2                # m_cur_PicardIte > 2 (:3198)
            elif cur_step < 50:
                converged = max_resH < 0.05
            else:
                converged = max_resH < eps_picard
            ite_idx += 1
            if converged:
                break
        out["picard_iterations"] = ite_idx
        self.update_gap(dt)
        return out

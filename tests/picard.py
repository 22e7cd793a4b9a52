"""One single-level Picard step of AmrHydro::timeStepFAS (src/AmrHydro.cpp:2477-3235, 3248-3408) as a sequence of
kernel calls, written once against a tiny backend interface and run on the oracle (CPU) and on the library (GPU): the time /
Picard loops are host code in the reference, their field kernels are what the library provides.
Gap update: explicit Euler (solver.use_ImplDiff = false) or the implicit diffusion solve SolveForGap_nl (true; src/AmrHydro.cpp:
3378-3455, 594-662)."""
import ctypes as C

import numpy as np

from oracle import binding as ob
from suhmo_b200 import synthetic as syn

CELL, XFACE, YFACE = 0, 1, 2


def picard_params(cls, cfg, **over):
    kw = dict(rho_i=910.0, rho_w=1000.0, gravity=9.8, G=0.05, L=334000.0, ct=7.5e-8, cw=4220.0, ub0=1e-6, basal_friction=1,
              A=cfg.A, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr, DiffFactor=1e-2, n_moulins=len(cfg.moulins) or -1, ramp=1.0,
              distributed_input=cfg.distributed_input, use_mask_rhs_b=int(cfg.ibc == "valley"), use_ImplDiff=0)
    kw.update(over)
    return cls(**kw)


def _dx(cfg):
    a = np.array(cfg.dx, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class OracleBackend:
    """kernel calls on oracle fields"""

    def __init__(self, orc, impl_diff=False):
        self.L, self.orc, self.cfg = ob.lib(), orc, orc.cfg
        self.prm, self.bc = orc.prm, orc.bc
        self.impl_diff = impl_diff
        self.q = picard_params(ob.PicardParams, orc.cfg, use_ImplDiff=int(impl_diff))

    def new(self, ncomp=1, ng=0, cent=CELL):
        return ob.Field(self.orc.layout, ncomp, ng, cent)

    def exchange(self, f): self.L.orc_exchange_full(f.h)
    def copy_ghost(self, f): self.L.orc_copy_ghost(f.h)
    def extrap_ghost(self, f): self.L.orc_extrap_ghost(f.h)
    def apply_bc(self, f): self.L.orc_apply_bc(f.h, C.byref(self.bc), _dx(self.cfg)[1], 0)
    def cell_to_edge(self, c, ex, ey): self.L.orc_cell_to_edge(c.h, ex.h, ey.h)
    def edge_to_cell(self, ex, ey, c2): self.L.orc_edge_to_cell(ex.h, ey.h, c2.h)
    def mac_gradient(self, phi, mask, gx, gy): self.L.orc_mac_gradient(phi.h, None if mask is None else mask.h, _dx(self.cfg)[1], gx.h, gy.h)
    def icemask_ec(self, m, mx, my): self.L.orc_icemask_ec(m.h, mx.h, my.h)
    def compute_re(self, Re, B, gradH): self.L.orc_compute_re(C.byref(self.prm), B.h, gradH.h, Re.h)
    def compute_qw(self, Bec, Reec, gec, Qw): self.L.orc_compute_qw(C.byref(self.prm), Bec.h, Reec.h, gec.h, Qw.h)
    def scaprod(self, a, b1, b2, p1, p2): self.L.orc_compute_scaprod(a.h, b1.h, b2.h, p1.h, p2.h)
    def dcoeff(self, D, MRec, Bec, IMec): self.L.orc_compute_dcoeff(D.h, MRec.h, Bec.h, IMec.h, self.q.rho_i, self.orc.cfg.cutOffBcoef)
    def difterm(self, phi, Dt, D0, D1): self.L.orc_compute_difterm(phi.h, _dx(self.cfg)[1], Dt.h, D0.h, D1.h)
    def melting_rate(self, H, zb, Pi, IM, B, qgh, qgz, Pw, mR): self.L.orc_calc_melting_rate(C.byref(self.q), H.h, zb.h, Pi.h, IM.h, B.h, qgh.h, qgz.h, Pw.h, mR.h)
    def rhs_head(self, R, mR, B, BH, BL, MV, MS, Dt, IM): self.L.orc_rhs_head(C.byref(self.q), R.h, mR.h, B.h, BH.h, BL.h, MV.h, MS.h, Dt.h, IM.h)
    def rhs_gap(self, R, Pi, Pw, mR, B, DT, IM, BH, BL, MV, dt): self.L.orc_rhs_gap(C.byref(self.q), R.h, Pi.h, Pw.h, mR.h, B.h, DT.h, IM.h, BH.h, BL.h, MV.h, dt)
    def gap_euler(self, nB, oB, R, dt): self.L.orc_gap_euler(nB.h, oB.h, R.h, dt)
    def copy(self, dst, src): dst.copy_from(src)

    def bcoeff(self, F):  # aCoeff_bCoeff == COMPUTEBCOEFF on the edge data; the operator's UpdateOperator does the same from h
        self.orc.op().update_operator(F["head"])

    def solve_head(self, F, ncyc):
        it, hist = self.orc.solver().solve(F["head"], F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=ncyc))
        return hist

    def setval(self, f, v): f.setval(v)

    def solve_gap(self, aC, Dc, gap, rhs, dt, cur_step):
        s = ob.LinSolver(self.orc.layout, self.cfg.dx[0], 1.0, dt * self.q.DiffFactor, aC, Dc[0], Dc[1])
        sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10 if cur_step < 50 else 5, iter_min=2, eps=1e-7,
                                   hang=1e-6, norm_thresh=1e-7)
        it, hist = s.solve(gap, rhs, sp)
        s.free()
        return hist


class GpuBackend:
    """the same calls through the C ABI"""

    def __init__(self, gpu, impl_diff=False):
        from suhmo_b200 import amr, capi
        self.amr, self.capi, self.gpu, self.cfg = amr, capi, gpu, gpu.orc.cfg
        self.Lib = capi.lib()
        self.prm, self.bc = gpu.prm, gpu.bc
        self.impl_diff = impl_diff
        self.q = picard_params(capi.PicardParams, self.cfg, use_ImplDiff=int(impl_diff))
        self.mg = None

    def ck(self, st): self.capi.check(st)
    def new(self, ncomp=1, ng=0, cent=CELL): return self.amr.LevelData(self.gpu.layout, ncomp, ng, cent)
    def exchange(self, f): f.exchange(True)
    def copy_ghost(self, f): self.amr.CopyGhostCells(f)
    def extrap_ghost(self, f): self.amr.ExtrapGhostCells(f)
    def apply_bc(self, f): self.ck(self.Lib.sg_apply_bc(f.h, C.byref(self.bc), _dx(self.cfg)[1], 0))
    def cell_to_edge(self, c, ex, ey): self.ck(self.Lib.sg_cell_to_edge(c.h, ex.h, ey.h))
    def edge_to_cell(self, ex, ey, c2): self.ck(self.Lib.sg_edge_to_cell(ex.h, ey.h, c2.h))
    def mac_gradient(self, phi, mask, gx, gy): self.ck(self.Lib.sg_mac_gradient(phi.h, None if mask is None else mask.h, _dx(self.cfg)[1], gx.h, gy.h))
    def icemask_ec(self, m, mx, my): self.ck(self.Lib.sg_icemask_ec(m.h, mx.h, my.h))
    def compute_re(self, Re, B, gradH): self.ck(self.Lib.sg_compute_re(C.byref(self.prm), Re.h, B.h, gradH.h))
    def compute_qw(self, Bec, Reec, gec, Qw): self.ck(self.Lib.sg_compute_qw(C.byref(self.prm), Bec.h, Reec.h, gec.h, Qw.h))
    def scaprod(self, a, b1, b2, p1, p2): self.ck(self.Lib.sg_compute_scaprod(a.h, b1.h, b2.h, p1.h, p2.h))
    def dcoeff(self, D, MRec, Bec, IMec): self.ck(self.Lib.sg_compute_dcoeff(D.h, MRec.h, Bec.h, IMec.h, self.q.rho_i, self.cfg.cutOffBcoef))
    def difterm(self, phi, Dt, D0, D1): self.ck(self.Lib.sg_compute_difterm(phi.h, _dx(self.cfg)[1], Dt.h, D0.h, D1.h))
    def melting_rate(self, H, zb, Pi, IM, B, qgh, qgz, Pw, mR): self.ck(self.Lib.sg_calc_melting_rate(C.byref(self.q), H.h, zb.h, Pi.h, IM.h, B.h, qgh.h, qgz.h, Pw.h, mR.h))
    def rhs_head(self, R, mR, B, BH, BL, MV, MS, Dt, IM): self.ck(self.Lib.sg_rhs_head(C.byref(self.q), R.h, mR.h, B.h, BH.h, BL.h, MV.h, MS.h, Dt.h, IM.h))
    def rhs_gap(self, R, Pi, Pw, mR, B, DT, IM, BH, BL, MV, dt): self.ck(self.Lib.sg_rhs_gap(C.byref(self.q), R.h, Pi.h, Pw.h, mR.h, B.h, DT.h, IM.h, BH.h, BL.h, MV.h, dt))
    def gap_euler(self, nB, oB, R, dt): self.ck(self.Lib.sg_gap_euler(nB.h, oB.h, R.h, dt))

    def copy(self, dst, src):
        op = self.op0()
        op.assignLocal(dst, src)

    def op0(self):
        if not hasattr(self, "_op0"):
            self._op0 = self.gpu.factory.AMRnewOp(0)
        return self._op0

    def bcoeff(self, F):
        self.op0().UpdateOperator(F["head"], None, 0, 0, False)

    def solve_head(self, F, ncyc):
        if self.mg is None:
            self.mg = self.amr.AMRFASMultiGrid().define(self.gpu.factory, 1)
            self.mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
        else:
            self.mg.refresh()  # the reference rebuilds factory + solver per Picard iteration (src/AmrHydro.cpp:704-735)
        it, hist, st = self.mg.solve([F["head"]], [F["rhs"]], fixed_cycles=ncyc)
        return hist


    def setval(self, f, v):
        f.upload([np.full(f.fab_shape(b), float(v)) if f.layout.owned(b) else None for b in range(len(f.layout.boxes))])

    def solve_gap(self, aC, Dc, gap, rhs, dt, cur_step):
        it, hist, st = self.amr.SolveForGap_nl(self.gpu.ctx, [self.gpu.layout], [aC], [Dc[0]], [Dc[1]], [], self.cfg.dx, [gap], [rhs], dt,
                                               self.q.DiffFactor, cur_step)
        return hist


def extra_fields(be, setter):
    """fields of the Picard body beyond the head-solve set, with simple deterministic contents"""
    cfg = be.cfg
    X = {}
    for k in ("mR", "Pw", "MV", "BH", "BL", "MS", "oldB", "gradH", "Re", "qgh", "qgz"):
        X[k] = be.new(2 if k in ("gradH", "qgh", "qgz") else 1, 1, CELL)
    for k in ("Dterm", "RHSb"):
        X[k] = be.new(1, 0, CELL)
    for k in ("Bec", "mRec", "gH", "gZ", "Dc", "Reec", "Qw", "t1", "t2", "IMec"):
        X[k] = (be.new(1, 0, XFACE), be.new(1, 0, YFACE))
    ny, nx = cfg.ny + 2, cfg.nx + 2
    jj, ii = np.meshgrid(np.arange(ny) - 1, np.arange(nx) - 1, indexing="ij")
    setter(X["MV"], np.full((ny, nx), 1e-6))
    setter(X["BH"], 0.012 + 0.002 * np.sin(0.37 * ii) * np.cos(0.21 * jj))   # some cells above, some below the gap height
    setter(X["BL"], np.full((ny, nx), 2.0))
    setter(X["mR"], 1e-7 * (1.0 + 0.3 * np.cos(0.11 * ii + 0.05 * jj)))
    g = syn.fields(cfg, ng=1)
    src = np.zeros((ny, nx))
    src[1:-1, 1:-1] = g["rhs"]
    setter(X["MS"], src)
    return X


def picard_step(be, F, X, dt=3600.0, npicard=2, ncyc=3, cur_step=0):
    """F: head-solve fields (head, B, Pi, zb, mask, rhs, bX, bY); X: extra_fields.  Returns residual histories."""
    use_mask = bool(be.cfg.use_mask_grad)
    hists = []
    be.copy(X["oldB"], F["B"])
    be.icemask_ec(F["mask"], *X["IMec"])
    for _ in range(npicard):
        # ghost fills and centering changes (src/AmrHydro.cpp:2482-2532)
        for k in (F["head"], F["B"], X["mR"]):
            be.exchange(k)
        be.copy_ghost(F["B"])
        be.apply_bc(F["head"])
        be.extrap_ghost(X["mR"])
        be.cell_to_edge(F["B"], *X["Bec"])
        be.cell_to_edge(X["mR"], *X["mRec"])
        # gradients and the diffusion coefficient (:2539-2572)
        be.mac_gradient(F["head"], F["mask"] if use_mask else None, *X["gH"])
        be.edge_to_cell(*X["gH"], X["gradH"])
        be.exchange(X["gradH"])
        be.extrap_ghost(X["gradH"])
        be.mac_gradient(F["zb"], F["mask"] if use_mask else None, *X["gZ"])
        for d in range(2):
            be.dcoeff(X["Dc"][d], X["mRec"][d], X["Bec"][d], X["IMec"][d])
        # Re, Qw (:2703-2789)
        be.compute_re(X["Re"], F["B"], X["gradH"])
        be.exchange(X["Re"])
        be.cell_to_edge(X["Re"], *X["Reec"])
        for d in range(2):
            be.compute_qw(X["Bec"][d], X["Reec"][d], X["gH"][d], X["Qw"][d])
        # RHS of the head equation (:2920-3079)
        for d in range(2):
            be.scaprod(X["Qw"][d], X["gH"][d], X["gZ"][d], X["t1"][d], X["t2"][d])
        be.edge_to_cell(*X["t1"], X["qgh"])
        be.edge_to_cell(*X["t2"], X["qgz"])
        be.difterm(F["B"], X["Dterm"], *X["Dc"])
        be.melting_rate(F["head"], F["zb"], F["Pi"], F["mask"], F["B"], X["qgh"], X["qgz"], X["Pw"], X["mR"])
        be.rhs_head(F["rhs"], X["mR"], F["B"], X["BH"], X["BL"], X["MV"], X["MS"], X["Dterm"], F["mask"])
        # coefficients and the head solve (:3087-3119)
        be.bcoeff(F)
        hists.append(be.solve_head(F, ncyc))
    # gap-height update (:3248-3408)
    be.rhs_gap(X["RHSb"], F["Pi"], X["Pw"], X["mR"], F["B"], X["Dterm"], F["mask"], X["BH"], X["BL"], X["MV"], dt)
    if be.impl_diff:
        # implicit branch (:3378-3391, 3425-3455): a_gh_curr = B incl. ghosts, aCoef = 1, bCoef = Dcoef, SolveForGap_nl, copy back
        cur, aC = be.new(1, 1, CELL), be.new(1, 0, CELL)
        be.copy(cur, F["B"])
        be.setval(aC, 1.0)
        hists.append(be.solve_gap(aC, X["Dc"], cur, X["RHSb"], dt, cur_step))
        be.copy(F["B"], cur)
    else:
        be.gap_euler(F["B"], X["oldB"], X["RHSb"], dt)
    be.exchange(F["B"])
    be.copy_ghost(F["B"])
    return hists

"""Single-level time step of AmrHydro::timeStepFAS (src/AmrHydro.cpp:2255-4172) as host orchestration over the C ABI: the Picard
loop (ghost fills, centring changes, gradients, Re, water flux, melt rate, RHS_h, B(h) coefficients, FAS head solve, lagged
convergence test :3168-3229) and the gap-height update (explicit Euler :3394-3408 or the implicit solve SolveForGap_nl
:3378-3455).  The time / Picard loops are host code in the reference too; every field kernel they call is the library's
(SURVEY.md 8 a18, f1, f2).  The orchestration is written against a small backend interface: `GpuBackend` (here) issues the calls
through libsuhmo_gpu; the parity tests run the same sequence on the CPU oracle with a twin backend that lives in tests/."""
import ctypes as C

import numpy as np

from . import synthetic as syn

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



class Level0:
    """What GpuBackend needs of a level-0 problem on the device: ctx, layout, operator factory, parameter blocks, the Config."""

    def __init__(self, ctx, layout, factory, prm, bc, cfg):
        from . import amr
        self.amr, self.ctx, self.layout, self.factory, self.prm, self.bc, self.cfg = amr, ctx, layout, factory, prm, bc, cfg

    @classmethod
    def from_config(cls, ctx, cfg, use_NL=1, bcoeff_otf=1, seed=12345):
        """Level-0 problem of a synthetic.Config on the device: domainSplit boxes, the IBC's closed-form fields (what the
        reference's initData produces: src/*IBC.cpp), ghost cells filled as AmrHydro does before the first solve
        (src/AmrHydro.cpp:2360-2445), operator factory with alpha = 0, beta = -1 (src/AmrHydro.cpp:704-717).
        Returns (Level0, F) with F the head-solve fields head, B, Pi, zb, mask, rhs, a, bX, bY."""
        from . import amr
        boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        layout = amr.DisjointBoxLayout(ctx, boxes, (0, 0, cfg.nx - 1, cfg.ny - 1), cfg.periodic)
        spec = dict(head=(1, CELL), B=(1, CELL), Pi=(1, CELL), zb=(1, CELL), mask=(1, CELL), rhs=(0, CELL), a=(0, CELL), bX=(0, XFACE),
                    bY=(0, YFACE))
        F = {k: amr.LevelData(layout, 1, ng, cent) for k, (ng, cent) in spec.items()}
        g = syn.fields(cfg, ng=1, seed=seed)
        for k in ("head", "B", "Pi", "zb", "mask"):
            F[k].set_global(g[k], (-1, -1))
            F[k].exchange(True)
            if k != "head":
                amr.CopyGhostCells(F[k])
        F["rhs"].set_global(g["rhs"], (0, 0))
        bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi, cfg.bc_lo_val, cfg.bc_hi_val)
        prm = amr.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr, cutOffBcoef=cfg.cutOffBcoef,
                              use_NL=use_NL, use_mask_grad=cfg.use_mask_grad, bcoeff_otf=bcoeff_otf)
        factory = amr.VCAMRNonLinearPoissonOpFactory().define(ctx, [layout], [], cfg.dx, bc, 0.0, [F["a"]], -1.0, [F["bX"]], [F["bY"]], prm,
                                                              [F["B"]], [F["Pi"]], [F["zb"]], [F["mask"]])
        return cls(ctx, layout, factory, prm, bc, cfg), F


class GpuBackend:
    """the same calls through the C ABI"""

    def __init__(self, gpu, impl_diff=False, cfg=None):
        from . import amr, capi
        self.amr, self.capi, self.gpu = amr, capi, gpu
        self.cfg = cfg if cfg is not None else gpu.cfg
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

    def solve_head_converged(self, F, cur_step):
        """SolveForHead_nl with the reference's parameters and stop logic (src/AmrHydro.cpp:737-766)"""
        early = cur_step < 50
        if self.mg is None:
            self.mg = self.amr.AMRFASMultiGrid().define(self.gpu.factory, 1)
        else:
            self.mg.refresh()
        self.mg.setSolverParameters(4, 4, 10 if early else 16, 1, 100, 1e-10 if early else 1e-7, 1e-4 if early else 0.01, 1e-7)
        self.mg.params.imin, self.mg.params.iter_min = (20 if early else 5), 2
        it, hist, st = self.mg.solve([F["head"]], [F["rhs"]])
        return hist

    def max_abs(self, f):
        return self.op0().norm(f, 0)

    def max_abs_diff(self, a, b):
        if not hasattr(self, "_tmp"):
            self._tmp = self.new(1, 1, CELL)
        self.op0().axby(self._tmp, a, b, 1.0, -1.0)
        return self.op0().norm(self._tmp, 0)

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


def picard_iteration(be, F, X, solve):
    """Body of the Picard loop (src/AmrHydro.cpp:2477-3119): everything between the top of `while (!converged_h)` and the head
    solve.  `solve(F)` runs the head solve and returns its residual history."""
    use_mask = bool(be.cfg.use_mask_grad)
    # ghost fills and centring changes (:2482-2532)
    for k in (F["head"], F["B"], X["mR"]):
        be.exchange(k)
    be.copy_ghost(F["B"])
    be.apply_bc(F["head"])
    if "headLag" in X:
        be.copy(X["headLag"], F["head"])      # h into h_lag, after the ghost fill (:2522)
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
    return solve(F)


def update_gap(be, F, X, dt, cur_step):
    """Gap-height update after the Picard loop (src/AmrHydro.cpp:3248-3455).  Returns the implicit solve's history or None."""
    hist = None
    use_mask = bool(be.cfg.use_mask_grad)
    # Re, Qw and the melt rate once more with the fresh head (evaluate_Re_quadratic(lev, true) ... Calc_meltingRate, :3256-3330)
    be.mac_gradient(F["head"], F["mask"] if use_mask else None, *X["gH"])
    be.edge_to_cell(*X["gH"], X["gradH"])
    be.exchange(X["gradH"])
    be.extrap_ghost(X["gradH"])
    be.compute_re(X["Re"], F["B"], X["gradH"])
    be.exchange(X["Re"])
    be.cell_to_edge(X["Re"], *X["Reec"])
    for d in range(2):
        be.compute_qw(X["Bec"][d], X["Reec"][d], X["gH"][d], X["Qw"][d])
        be.scaprod(X["Qw"][d], X["gH"][d], X["gZ"][d], X["t1"][d], X["t2"][d])
    be.edge_to_cell(*X["t1"], X["qgh"])
    be.edge_to_cell(*X["t2"], X["qgz"])
    be.melting_rate(F["head"], F["zb"], F["Pi"], F["mask"], F["B"], X["qgh"], X["qgz"], X["Pw"], X["mR"])
    be.rhs_gap(X["RHSb"], F["Pi"], X["Pw"], X["mR"], F["B"], X["Dterm"], F["mask"], X["BH"], X["BL"], X["MV"], dt)
    if be.impl_diff:
        # implicit branch (:3378-3391, 3425-3455): a_gh_curr = B incl. ghosts, aCoef = 1, bCoef = Dcoef, SolveForGap_nl, copy back
        if "ghCur" not in X:
            X["ghCur"], X["aCoefGH"] = be.new(1, 1, CELL), be.new(1, 0, CELL)
            be.setval(X["aCoefGH"], 1.0)
        be.copy(X["ghCur"], F["B"])
        hist = be.solve_gap(X["aCoefGH"], X["Dc"], X["ghCur"], X["RHSb"], dt, cur_step)
        be.copy(F["B"], X["ghCur"])
    else:
        be.gap_euler(F["B"], X["oldB"], X["RHSb"], dt)
    be.exchange(F["B"])
    be.copy_ghost(F["B"])
    return hist


def picard_step(be, F, X, dt=3600.0, npicard=2, ncyc=3, cur_step=0):
    """Parity protocol: a fixed number of Picard iterations with a fixed number of V-cycles each, then the gap update.
    F: head-solve fields (head, B, Pi, zb, mask, rhs, bX, bY); X: extra_fields.  Returns the residual histories."""
    hists = []
    be.copy(X["oldB"], F["B"])
    be.icemask_ec(F["mask"], *X["IMec"])
    for _ in range(npicard):
        hists.append(picard_iteration(be, F, X, lambda F_: be.solve_head(F_, ncyc)))
        be.exchange(F["head"])
        be.apply_bc(F["head"])                 # head ghost cells refilled after the solve (:3143-3165)
    h = update_gap(be, F, X, dt, cur_step)
    if h is not None:
        hists.append(h)
    return hists


def time_step(be, F, X, dt, cur_step, eps_picard=1.0e-6, max_picard=100):
    """One time step as the reference runs it: Picard iterations until the lagged change of head, max|h_lag - h| / max h, falls
    below 0.05 (while cur_step < 50; additionally more than two iterations while cur_step < 2) or solver.eps_PicardIte afterwards
    (src/AmrHydro.cpp:3168-3229), each with a head solve under the reference's stop logic (:737-762), then the gap update.
    Returns {"picard_iterations", "x_h": [...], "head_cycles": [...], "gap_cycles"}.  max h is taken as max|h| (head is positive
    in every SUHMO set-up)."""
    be.copy(X["oldB"], F["B"])
    be.icemask_ec(F["mask"], *X["IMec"])
    if "headLag" not in X:
        X["headLag"] = be.new(1, 1, CELL)
    out = {"x_h": [], "head_cycles": []}
    ite = 0
    while True:
        hist = picard_iteration(be, F, X, lambda F_: be.solve_head_converged(F_, cur_step))
        out["head_cycles"].append(len(hist) - 1)
        be.exchange(F["head"])
        be.apply_bc(F["head"])                 # head ghost cells refilled before the test (:3143-3165)
        x_h = be.max_abs_diff(X["headLag"], F["head"]) / be.max_abs(F["head"])
        out["x_h"].append(x_h)
        if ite > max_picard:
            raise RuntimeError("does not converge (Picard iterations > 100)")   # MayDay::Error("Abort"), :3190-3195
        if cur_step < 2:
            done = x_h < 0.05 and ite > 2
        elif cur_step < 50:
            done = x_h < 0.05
        else:
            done = x_h < eps_picard
        ite += 1
        if done:
            break
    out["picard_iterations"] = ite
    g = update_gap(be, F, X, dt, cur_step)
    out["gap_cycles"] = None if g is None else len(g) - 1
    return out

"""One single-level Picard step of AmrHydro::timeStepFAS (src/AmrHydro.cpp:2477-3235, 3248-3408) as a sequence of
kernel calls, written once against a tiny backend interface and run on the oracle (CPU) and on the library (GPU): the time /
Picard loops are host code in the reference, their field kernels are what the library provides.
Gap update: explicit Euler (solver.use_ImplDiff = false) or the implicit diffusion solve SolveForGap_nl (true; src/AmrHydro.cpp:
3378-3455, 594-662)."""
import ctypes as C

import numpy as np

from oracle import binding as ob
from suhmo_b200 import synthetic as syn  # noqa: F401
from suhmo_b200.timestep import GpuBackend as _GpuBackend, _dx, extra_fields, picard_params, picard_step, time_step  # noqa: F401

CELL, XFACE, YFACE = 0, 1, 2



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

    def solve_head_converged(self, F, cur_step):
        early = cur_step < 50
        sp = ob.make_solver_params(pre=4, post=4, bottom=10 if early else 16, max_iter=100, imin=20 if early else 5, iter_min=2,
                                   eps=1e-10 if early else 1e-7, hang=1e-4 if early else 0.01, norm_thresh=1e-7)
        it, hist = self.orc.solver().solve(F["head"], F["rhs"], sp)
        return hist

    def max_abs(self, f): return f.norm(0)

    def max_abs_diff(self, a, b):
        tmp = self.new(1, 1, CELL)
        self.L.orc_axby(tmp.h, a.h, b.h, 1.0, -1.0)
        return tmp.norm(0)

    def solve_gap(self, aC, Dc, gap, rhs, dt, cur_step):
        s = ob.LinSolver(self.orc.layout, self.cfg.dx[0], 1.0, dt * self.q.DiffFactor, aC, Dc[0], Dc[1])
        sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10 if cur_step < 50 else 5, iter_min=2, eps=1e-7,
                                   hang=1e-6, norm_thresh=1e-7)
        it, hist = s.solve(gap, rhs, sp)
        s.free()
        return hist



class GpuBackend(_GpuBackend):
    """the package's backend on a tests.problem.GpuSide (whose Config sits under .orc)"""

    def __init__(self, gpu, impl_diff=False):
        super().__init__(gpu, impl_diff, cfg=gpu.orc.cfg)

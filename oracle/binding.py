"""ctypes binding of the CPU oracle (oracle/libsuhmo_oracle.so).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (suhmo_b200/) never imports this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

CELL, XFACE, YFACE = 0, 1, 2


class Params(C.Structure):
    _fields_ = [("A", C.c_double), ("cutOffbr", C.c_double), ("maxOffbr", C.c_double),
                ("omega", C.c_double), ("nu", C.c_double), ("cutOffBcoef", C.c_int),
                ("use_NL", C.c_int), ("use_mask_grad", C.c_int), ("bcoeff_otf", C.c_int)]


class BC(C.Structure):
    _fields_ = [("lo_type", C.c_int * 2), ("hi_type", C.c_int * 2),
                ("lo_val", C.c_double * 2), ("hi_val", C.c_double * 2)]


class PicardParams(C.Structure):
    _fields_ = [("rho_i", C.c_double), ("rho_w", C.c_double), ("gravity", C.c_double), ("G", C.c_double), ("L", C.c_double),
                ("ct", C.c_double), ("cw", C.c_double), ("ub0", C.c_double), ("basal_friction", C.c_int),
                ("A", C.c_double), ("cutOffbr", C.c_double), ("maxOffbr", C.c_double), ("DiffFactor", C.c_double),
                ("n_moulins", C.c_int), ("ramp", C.c_double), ("distributed_input", C.c_double),
                ("use_mask_rhs_b", C.c_int), ("use_ImplDiff", C.c_int)]


class SolverParams(C.Structure):
    _fields_ = [("pre", C.c_int), ("post", C.c_int), ("bottom", C.c_int), ("max_iter", C.c_int),
                ("imin", C.c_int), ("iter_min", C.c_int), ("eps", C.c_double), ("hang", C.c_double),
                ("norm_thresh", C.c_double), ("fixed_cycles", C.c_int)]


def build(force=False):
    so = os.path.join(_HERE, "libsuhmo_oracle.so")
    src = [os.path.join(_HERE, f) for f in ("suhmo_oracle.c", "suhmo_oracle_r2.inc", "suhmo_oracle_r3.inc", "suhmo_oracle.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "all"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp, ci, cd = C.c_void_p, C.c_int, C.c_double
    ip, dp = C.POINTER(C.c_int), C.POINTER(C.c_double)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("orc_layout_create", vp, ci, ip, ip, ip)
    sig("orc_layout_coarsen", vp, vp, ci)
    sig("orc_layout_coarsenable", ci, vp, ci)
    sig("orc_layout_nbox", ci, vp)
    sig("orc_layout_box", None, vp, ci, ip)
    sig("orc_layout_free", None, vp)
    sig("orc_field_create", vp, vp, ci, ci, ci)
    sig("orc_field_free", None, vp)
    sig("orc_field_fab", dp, vp, ci, ip, ip)
    sig("orc_field_setval", None, vp, cd)
    sig("orc_field_copy", None, vp, vp)
    for n in ("orc_exchange_faces", "orc_exchange_full", "orc_extrap_ghost", "orc_copy_ghost", "orc_set_to_zero"):
        sig(n, None, vp)
    sig("orc_apply_bc", None, vp, C.POINTER(BC), dp, ci)
    sig("orc_cell_to_edge", None, vp, vp, vp)
    sig("orc_edge_to_cell", None, vp, vp, vp)
    sig("orc_mac_gradient", None, vp, vp, dp, vp, vp)
    sig("orc_divergence", None, vp, vp, dp, vp)
    sig("orc_icemask_ec", None, vp, vp, vp)
    sig("orc_coarse_average", None, vp, vp, ci)
    sig("orc_coarse_average_face", None, vp, vp, ci)
    sig("orc_compute_nl", None, C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp)
    sig("orc_compute_re", None, C.POINTER(Params), vp, vp, vp)
    sig("orc_assign", None, vp, vp)
    sig("orc_incr", None, vp, vp, cd)
    sig("orc_axby", None, vp, vp, vp, cd, cd)
    sig("orc_scale", None, vp, cd)
    sig("orc_dot", cd, vp, vp)
    sig("orc_norm", cd, vp, ci)
    sig("orc_op_create", vp, vp, dp, cd, cd, C.POINTER(BC), C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp)
    sig("orc_op_free", None, vp)
    sig("orc_op_reset_lambda", None, vp)
    sig("orc_op_lambda", vp, vp)
    sig("orc_op_relax", None, vp, vp, vp, ci)
    sig("orc_op_residual", None, vp, vp, vp, vp)
    sig("orc_op_apply", None, vp, vp, vp, ci)
    sig("orc_op_restrict_residual", None, vp, vp, vp, vp)
    sig("orc_op_restrict_r", None, vp, vp, vp)
    sig("orc_op_prolong_increment", None, vp, vp, vp)
    sig("orc_op_update_operator", None, vp, vp)
    sig("orc_op_average_operator", None, vp, vp, ci)
    sig("orc_solver_create", vp, vp, dp, cd, cd, C.POINTER(BC), C.POINTER(Params), vp, vp, vp, vp, vp, vp, vp)
    sig("orc_solver_free", None, vp)
    sig("orc_solver_depth", ci, vp)
    sig("orc_solver_op", vp, vp, ci)
    sig("orc_solver_solve", ci, vp, vp, vp, C.POINTER(SolverParams), dp)
    sig("orc_solver_vcycle", None, vp, vp, vp, C.POINTER(SolverParams), ci)
    sig("orc_solver_cell_updates", cd, vp, C.POINTER(SolverParams))
    sig("orc_set_threads", None, ci)
    pvp = C.POINTER(C.c_void_p)
    sig("orc_copy_to", None, vp, vp, ci)
    sig("orc_cf_interp", None, vp, vp, ci, cd)
    sig("orc_op_set_ref_to_coarser", None, vp, ci)
    sig("orc_op_relax_nf", None, vp, vp, vp, vp, ci)
    sig("orc_op_residual_nf", None, vp, vp, vp, vp, vp)
    sig("orc_op_reflux", None, vp, vp, vp, vp, vp)
    sig("orc_op_amr_operator", None, vp, vp, vp, vp, vp, ci, vp)
    sig("orc_op_amr_residual", None, vp, vp, vp, vp, vp, vp, ci, vp)
    sig("orc_op_amr_restrict_s", None, vp, vp, vp, vp, vp, vp, ci)
    sig("orc_op_amr_prolong_s", None, vp, vp, vp, vp)
    sig("orc_op_amr_prolong_s2", None, vp, vp, vp, vp, vp)
    sig("orc_zero_covered", None, vp, vp, ci)
    sig("orc_op_amr_norm", cd, vp, vp, ci, ci)
    sig("orc_op_update_operator_amr", None, vp, vp, vp, vp)
    sig("orc_amr_solver_create", vp, ci, pvp, dp, cd, cd, C.POINTER(BC), C.POINTER(Params), pvp, pvp, pvp, pvp, pvp, pvp, pvp)
    sig("orc_amr_solver_free", None, vp)
    sig("orc_amr_solver_op", vp, vp, ci)
    sig("orc_amr_solver_mg0", vp, vp)
    sig("orc_amr_solver_residual", vp, vp, ci)
    sig("orc_amr_solver_vcycle", None, vp, pvp, pvp, ci, C.POINTER(SolverParams))
    sig("orc_amr_solver_resnorm", cd, vp, pvp, pvp, ci)
    sig("orc_amr_solver_solve", ci, vp, pvp, pvp, ci, C.POINTER(SolverParams), dp)
    sig("orc_lin_solver_create", vp, vp, cd, cd, cd, vp, vp, vp)
    sig("orc_lin_solver_free", None, vp)
    sig("orc_lin_solver_depth", ci, vp)
    sig("orc_lin_solver_bottom_iters", ci, vp)
    sig("orc_linop_lambda", vp, vp, ci)
    sig("orc_linop_relax", None, vp, ci, vp, vp, ci)
    sig("orc_linop_residual", None, vp, ci, vp, vp, vp)
    sig("orc_linop_apply", None, vp, ci, vp, vp)
    sig("orc_linop_restrict_residual", None, vp, ci, vp, vp, vp)
    sig("orc_linop_prolong_increment", None, vp, ci, vp, vp)
    sig("orc_linop_precond", None, vp, ci, vp, vp)
    sig("orc_lin_solver_bottom_solve", ci, vp, vp, vp)
    sig("orc_lin_solver_vcycle", None, vp, vp, vp, C.POINTER(SolverParams))
    sig("orc_lin_solver_solve", ci, vp, vp, vp, C.POINTER(SolverParams), dp)
    sig("orc_amr_solver_cell_updates", cd, vp, C.POINTER(SolverParams), ci)
    pq = C.POINTER(PicardParams)
    sig("orc_compute_qw", None, C.POINTER(Params), vp, vp, vp, vp)
    sig("orc_compute_scaprod", None, vp, vp, vp, vp, vp)
    sig("orc_compute_dcoeff", None, vp, vp, vp, vp, cd, ci)
    sig("orc_compute_difterm", None, vp, dp, vp, vp, vp)
    sig("orc_time_varying_recharge", None, vp, vp, cd, cd)
    sig("orc_calc_melting_rate", None, pq, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("orc_rhs_head", None, pq, vp, vp, vp, vp, vp, vp, vp, vp, vp)
    sig("orc_rhs_gap", None, pq, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, cd)
    sig("orc_gap_euler", None, vp, vp, vp, cd)
    sig("orc_tag_cells_level", None, vp, cd, cd, ci, ip, vp, ci)
    sig("orc_op_set_alpha_beta", None, vp, cd, cd)
    sig("orc_op_precond", None, vp, vp, vp)
    sig("orc_op_precond3", None, vp, vp, vp, vp)
    sig("orc_op_get_flux", None, vp, vp, vp, ci, ci, cd)
    sig("orc_op_finer_operator_changed", None, vp, vp, ci)
    sig("orc_homogeneous_cf_interp", None, vp, dp, dp)
    sig("orc_op_diagonal_scale", None, vp, vp)
    sig("orc_op_divide_by_identity_coef", None, vp, vp)
    sig("orc_pwl_fill_patch", None, vp, vp, ci)
    sig("orc_fine_interp", None, vp, vp, ci)
    sig("orc_regrid_transfer", None, vp, vp, vp, ci)
    sig("orc_compute_bcoeff", None, C.POINTER(Params), vp, vp, vp, vp)
    sig("orc_moulin_nonorm", None, vp, dp, ci, dp, dp)
    sig("orc_moulin_integral", None, vp, vp, dp, ci, dp)
    sig("orc_moulin_source", None, vp, vp, ci, dp, dp, cd, cd)
    _LIB = L
    return L


def _ia(a):
    a = np.ascontiguousarray(a, dtype=np.int32)
    return a, a.ctypes.data_as(C.POINTER(C.c_int))


def _da(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


class Layout:
    """DisjointBoxLayout + ProblemDomain."""

    def __init__(self, boxes, domain, periodic, _h=None):
        self.boxes = np.ascontiguousarray(boxes, dtype=np.int32).reshape(-1, 4)
        self.domain = tuple(int(v) for v in domain)
        self.periodic = tuple(int(v) for v in periodic)
        if _h is None:
            b, bp = _ia(self.boxes)
            d, dpp = _ia(self.domain)
            p, pp = _ia(self.periodic)
            _h = lib().orc_layout_create(len(self.boxes), bp, dpp, pp)
        self.h = _h

    def coarsen(self, r):
        h = lib().orc_layout_coarsen(self.h, r)
        nb = lib().orc_layout_nbox(h)
        out = np.zeros((nb, 4), dtype=np.int32)
        for b in range(nb):
            lib().orc_layout_box(h, b, out[b].ctypes.data_as(C.POINTER(C.c_int)))
        dom = (self.domain[0] // r, self.domain[1] // r, self.domain[2] // r, self.domain[3] // r)
        return Layout(out, dom, self.periodic, _h=h)

    def coarsenable(self, r):
        return bool(lib().orc_layout_coarsenable(self.h, r))


class Field:
    """LevelData<FArrayBox> (cell) or one direction of a LevelData<FluxBox>."""

    def __init__(self, layout, ncomp=1, ng=0, cent=CELL, _h=None):
        self.layout, self.ncomp, self.ng, self.cent = layout, ncomp, ng, cent
        self.h = _h if _h is not None else lib().orc_field_create(layout.h, ncomp, ng, cent)
        self._own = _h is None

    def fab(self, b):
        """numpy view [comp, j, i] of box b's array (ghosts included) and the index of element (0,0)."""
        dims = (C.c_int * 3)()
        lo = (C.c_int * 2)()
        p = lib().orc_field_fab(self.h, b, dims, lo)
        arr = np.ctypeslib.as_array(p, shape=(dims[2], dims[1], dims[0]))
        return arr, (lo[0], lo[1])

    def set_global(self, g, glo):
        """fill every box's whole array (ghosts included) from global array g[comp?, j, i] whose (0,0) is index glo."""
        g = np.asarray(g, dtype=np.float64)
        if g.ndim == 2:
            g = g[None]
        for b in range(len(self.layout.boxes)):
            a, lo = self.fab(b)
            j0, i0 = lo[1] - glo[1], lo[0] - glo[0]
            a[...] = g[:, j0:j0 + a.shape[1], i0:i0 + a.shape[2]]

    def get_global(self, fill=np.nan):
        """valid data of all boxes gathered on the domain (cells) or domain faces."""
        d = self.layout.domain
        nx, ny = d[2] - d[0] + 1 + (self.cent == XFACE), d[3] - d[1] + 1 + (self.cent == YFACE)
        out = np.full((self.ncomp, ny, nx), fill)
        for b, bx in enumerate(self.layout.boxes):
            a, lo = self.fab(b)
            vx, vy = bx[2] - bx[0] + 1 + (self.cent == XFACE), bx[3] - bx[1] + 1 + (self.cent == YFACE)
            oj, oi = bx[1] - lo[1], bx[0] - lo[0]
            out[:, bx[1] - d[1]:bx[1] - d[1] + vy, bx[0] - d[0]:bx[0] - d[0] + vx] = a[:, oj:oj + vy, oi:oi + vx]
        return out[0] if self.ncomp == 1 else out

    def setval(self, v):
        lib().orc_field_setval(self.h, float(v))

    def copy_from(self, other):
        lib().orc_field_copy(self.h, other.h)

    def norm(self, p=0):
        return lib().orc_norm(self.h, p)


def make_bc(lo_type, hi_type, lo_val=(0.0, 0.0), hi_val=(0.0, 0.0)):
    bc = BC()
    for d in range(2):
        bc.lo_type[d], bc.hi_type[d] = lo_type[d], hi_type[d]
        bc.lo_val[d], bc.hi_val[d] = lo_val[d], hi_val[d]
    return bc


def make_params(A=2.5e-25, cutOffbr=0.0, maxOffbr=10000.0, omega=1e-3, nu=1.787e-6, cutOffBcoef=0,
                use_NL=1, use_mask_grad=0, bcoeff_otf=1):
    return Params(A, cutOffbr, maxOffbr, omega, nu, cutOffBcoef, use_NL, use_mask_grad, bcoeff_otf)


def make_solver_params(pre=4, post=4, bottom=16, max_iter=100, imin=5, iter_min=2, eps=1e-7, hang=0.01,
                       norm_thresh=1e-7, fixed_cycles=0):
    return SolverParams(pre, post, bottom, max_iter, imin, iter_min, eps, hang, norm_thresh, fixed_cycles)


def _h(x):
    return None if x is None else x.h


def _harr(fs):
    return (C.c_void_p * len(fs))(*[f.h for f in fs])


def cf_interp(phiF, phiC, r, dxFine):
    lib().orc_cf_interp(phiF.h, phiC.h, r, dxFine)


def copy_to(dst, src, with_ghosts=False):
    lib().orc_copy_to(dst.h, src.h, int(with_ghosts))


def zero_covered(crse, fineLayout, r=2):
    lib().orc_zero_covered(crse.h, fineLayout.h, r)


def amr_norm(coarResid, fineLayout, r, ord_):
    return lib().orc_op_amr_norm(coarResid.h, None if fineLayout is None else fineLayout.h, r, ord_)


class Op:
    def __init__(self, layout, dx, alpha, beta, bc, prm, aCoef, bX, bY, B, Pi, zb, mask, _h=None):
        self.layout = layout
        self.fields = (aCoef, bX, bY, B, Pi, zb, mask)
        self.bc, self.prm = bc, prm
        if _h is None:
            _, dxp = _da(dx)
            _h = lib().orc_op_create(layout.h, dxp, alpha, beta, C.byref(bc), C.byref(prm),
                                     aCoef.h, bX.h, bY.h, B.h, Pi.h, zb.h, mask.h)
        self.h = _h

    def relax(self, phi, rhs, n):
        lib().orc_op_relax(self.h, phi.h, rhs.h, n)

    def residual(self, res, phi, rhs):
        lib().orc_op_residual(self.h, res.h, phi.h, rhs.h)

    def apply(self, lhs, phi, homogeneous=False):
        lib().orc_op_apply(self.h, lhs.h, phi.h, int(homogeneous))

    def restrict_residual(self, resC, phiF, rhsF):
        lib().orc_op_restrict_residual(self.h, resC.h, phiF.h, rhsF.h)

    def restrict_r(self, phiC, phiF):
        lib().orc_op_restrict_r(self.h, phiC.h, phiF.h)

    def prolong_increment(self, phiF, corrC):
        lib().orc_op_prolong_increment(self.h, phiF.h, corrC.h)

    def update_operator(self, phi):
        lib().orc_op_update_operator(self.h, phi.h)

    def average_operator(self, finest, depth):
        lib().orc_op_average_operator(self.h, finest.h, depth)

    def lambda_field(self):
        return Field(self.layout, 1, 0, CELL, _h=lib().orc_op_lambda(self.h))

    # ---- AMR surface (phiCoarse / phiFine / finerOp may be None) ----
    def relax_nf(self, phi, phiC, rhs, n):
        lib().orc_op_relax_nf(self.h, phi.h, _h(phiC), rhs.h, n)

    def residual_nf(self, res, phi, phiC, rhs):
        lib().orc_op_residual_nf(self.h, res.h, phi.h, _h(phiC), rhs.h)

    def reflux(self, phiFine, phi, residual, finerOp):
        lib().orc_op_reflux(self.h, phiFine.h, phi.h, residual.h, finerOp.h)

    def amr_operator(self, lof, phiFine, phi, phiC, homogeneous=False, finerOp=None):
        lib().orc_op_amr_operator(self.h, lof.h, _h(phiFine), phi.h, _h(phiC), int(homogeneous), _h(finerOp))

    def amr_residual(self, res, phiFine, phi, phiC, rhs, homogeneous=False, finerOp=None):
        lib().orc_op_amr_residual(self.h, res.h, _h(phiFine), phi.h, _h(phiC), rhs.h, int(homogeneous), _h(finerOp))

    def amr_restrict_s(self, resC, residual, correction, coarseCorrection, scratch, skip_res):
        lib().orc_op_amr_restrict_s(self.h, resC.h, residual.h, correction.h, _h(coarseCorrection), scratch.h, int(skip_res))

    def amr_prolong_s(self, correction, coarseCorrection, temp):
        lib().orc_op_amr_prolong_s(self.h, correction.h, coarseCorrection.h, temp.h)

    def amr_prolong_s2(self, correction, coarseCorrection, temp, crseOp):
        lib().orc_op_amr_prolong_s2(self.h, correction.h, coarseCorrection.h, temp.h, crseOp.h)

    def update_operator_amr(self, phi, phiC, maskC):
        lib().orc_op_update_operator_amr(self.h, phi.h, _h(phiC), _h(maskC))

    # ---- virtuals the FAS path never calls (oracle/suhmo_oracle_r2.inc part 1) ----
    def set_alpha_beta(self, alpha, beta):
        lib().orc_op_set_alpha_beta(self.h, alpha, beta)

    def precond(self, phi, rhs):
        lib().orc_op_precond(self.h, phi.h, rhs.h)

    def precond3(self, phi, res, rhs):
        lib().orc_op_precond3(self.h, phi.h, res.h, rhs.h)

    def get_flux(self, flux, phi, dir_, ref=1, scale=1.0):
        lib().orc_op_get_flux(self.h, flux.h, phi.h, dir_, ref, scale)

    def finer_operator_changed(self, finer, factor):
        lib().orc_op_finer_operator_changed(self.h, finer.h, factor)

    def diagonal_scale(self, rhs):
        lib().orc_op_diagonal_scale(self.h, rhs.h)

    def divide_by_identity_coef(self, rhs):
        lib().orc_op_divide_by_identity_coef(self.h, rhs.h)


class Solver:
    def __init__(self, layout, dx, alpha, beta, bc, prm, aCoef, bX, bY, B, Pi, zb, mask):
        self.layout = layout
        self.keep = (bc, prm, aCoef, bX, bY, B, Pi, zb, mask)
        _, dxp = _da(dx)
        self.h = lib().orc_solver_create(layout.h, dxp, alpha, beta, C.byref(bc), C.byref(prm),
                                         aCoef.h, bX.h, bY.h, B.h, Pi.h, zb.h, mask.h)

    @property
    def depth(self):
        return lib().orc_solver_depth(self.h)

    def solve(self, phi, rhs, sp):
        n = max(sp.max_iter, sp.fixed_cycles) + 2
        hist = np.zeros(n)
        it = lib().orc_solver_solve(self.h, phi.h, rhs.h, C.byref(sp), hist.ctypes.data_as(C.POINTER(C.c_double)))
        return it, hist[:it + 1]

    def cell_updates(self, sp):
        return lib().orc_solver_cell_updates(self.h, C.byref(sp))

    def free(self):
        lib().orc_solver_free(self.h)
        self.h = None


class AmrSolver:
    """AMRFASMultiGrid over several levels (fields: lists per level)."""

    def __init__(self, layouts, dx0, alpha, beta, bc, prm, aCoef, bX, bY, B, Pi, zb, mask):
        self.layouts = layouts
        self.keep = (bc, prm, aCoef, bX, bY, B, Pi, zb, mask)
        _, dxp = _da(dx0)
        self.h = lib().orc_amr_solver_create(len(layouts), _harr(layouts), dxp, alpha, beta, C.byref(bc), C.byref(prm),
                                             _harr(aCoef), _harr(bX), _harr(bY), _harr(B), _harr(Pi), _harr(zb), _harr(mask))

    def op(self, lev):
        o = Op(self.layouts[lev], None, 0, 0, None, None, None, None, None, None, None, None, None, _h=lib().orc_amr_solver_op(self.h, lev))
        return o

    def residual_field(self, lev):
        return Field(self.layouts[lev], 1, 0, CELL, _h=lib().orc_amr_solver_residual(self.h, lev))

    def resnorm(self, phi, rhs, l_max):
        return lib().orc_amr_solver_resnorm(self.h, _harr(phi), _harr(rhs), l_max)

    def vcycle(self, phi, rhs, l_max, sp):
        lib().orc_amr_solver_vcycle(self.h, _harr(phi), _harr(rhs), l_max, C.byref(sp))

    def solve(self, phi, rhs, l_max, sp):
        n = max(sp.max_iter, sp.fixed_cycles) + 2
        hist = np.zeros(n)
        it = lib().orc_amr_solver_solve(self.h, _harr(phi), _harr(rhs), l_max, C.byref(sp), hist.ctypes.data_as(C.POINTER(C.c_double)))
        return it, hist[:it + 1]

    def cell_updates(self, sp, l_max):
        return lib().orc_amr_solver_cell_updates(self.h, C.byref(sp), l_max)

    def free(self):
        lib().orc_amr_solver_free(self.h)
        self.h = None


def tag_cells_level(field, vmin, vmax, tags_grow=0, tags_grow_dir=(0, 0), tags=None):
    d = field.layout.domain
    nx, ny = d[2] - d[0] + 1, d[3] - d[1] + 1
    acc = tags is not None
    out = np.ascontiguousarray(tags, dtype=np.uint8) if acc else np.zeros((ny, nx), dtype=np.uint8)
    gd, gp = _ia(tags_grow_dir)
    lib().orc_tag_cells_level(field.h, float(vmin), float(vmax), int(tags_grow), gp, out.ctypes.data_as(C.c_void_p), int(acc))
    return out


class LinSolver:
    """SolveForGap_nl's stock VCAMRPoissonOp2 + linear AMRMultiGrid + RelaxSolver on one AMR level."""

    def __init__(self, layout, dx, alpha, beta, aCoef, bX, bY):
        self.layout = layout
        self.keep = (aCoef, bX, bY)
        self.h = lib().orc_lin_solver_create(layout.h, float(dx), alpha, beta, aCoef.h, bX.h, bY.h)

    @property
    def depth(self):
        return lib().orc_lin_solver_depth(self.h)

    @property
    def bottom_iters(self):
        return lib().orc_lin_solver_bottom_iters(self.h)

    def layout_at(self, depth):
        return self.layout if depth == 0 else self.layout.coarsen(1 << depth)

    def lambda_field(self, depth=0):
        return Field(self.layout_at(depth), 1, 0, CELL, _h=lib().orc_linop_lambda(self.h, depth))

    def relax(self, phi, rhs, iterations, depth=0):
        lib().orc_linop_relax(self.h, depth, phi.h, rhs.h, iterations)

    def residual(self, res, phi, rhs, depth=0):
        lib().orc_linop_residual(self.h, depth, res.h, phi.h, rhs.h)

    def applyOp(self, lhs, phi, depth=0):
        lib().orc_linop_apply(self.h, depth, lhs.h, phi.h)

    def restrictResidual(self, resC, phi, rhs, depth=0):
        lib().orc_linop_restrict_residual(self.h, depth, resC.h, phi.h, rhs.h)

    def prolongIncrement(self, phi, corrC, depth=0):
        lib().orc_linop_prolong_increment(self.h, depth, phi.h, corrC.h)

    def preCond(self, phi, rhs, depth=0):
        lib().orc_linop_precond(self.h, depth, phi.h, rhs.h)

    def bottom_solve(self, phi, rhs):
        return lib().orc_lin_solver_bottom_solve(self.h, phi.h, rhs.h)

    def vcycle(self, corr, res, sp):
        lib().orc_lin_solver_vcycle(self.h, corr.h, res.h, C.byref(sp))

    def solve(self, phi, rhs, sp):
        n = max(sp.max_iter, sp.fixed_cycles) + 2
        hist = np.zeros(n)
        it = lib().orc_lin_solver_solve(self.h, phi.h, rhs.h, C.byref(sp), hist.ctypes.data_as(C.POINTER(C.c_double)))
        return it, hist[:it + 1]

    def free(self):
        lib().orc_lin_solver_free(self.h)
        self.h = None

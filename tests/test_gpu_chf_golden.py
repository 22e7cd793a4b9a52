"""The CUDA library, through the C ABI, against golden vectors of the REFERENCE's own Fortran kernels
(tests/golden/chf_kernels.npz: the .ChF sources of the reference executed through tools/chf_translate.py, see
tests/golden/make_chf_golden.py).  Same cases as tests/test_oracle_chf_golden.py, no oracle in between: lambda, two levelGSRB
iterations (every relax mode), residual, applyOp, restrictResidual, restrictR, prolongIncrement on one doubly periodic box with
alpha != 0, the ice mask < 0 on a fifth of the cells and both cut-offs of COMPUTENONLINEARTERMS active; NonLinear_level alone;
the MAC gradient with and without mask; the divergence.  Bit for bit."""
import ctypes as C
import os

import numpy as np
import pytest

from suhmo_b200 import amr
from suhmo_b200.capi import check, lib

pytestmark = pytest.mark.gpu

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "chf_kernels.npz"))
NX, NY = int(Z["nx"]), int(Z["ny"])
DX = tuple(float(v) for v in Z["dx"])
A, OMEGA, NU, CUT, MX = (float(v) for v in Z["prm"])
CELL, XF, YF = 0, 1, 2


def wrap(a):
    g = np.zeros((a.shape[0] + 2, a.shape[1] + 2))
    g[1:-1, 1:-1] = a
    g[1:-1, 0], g[1:-1, -1] = a[:, -1], a[:, 0]
    g[0, 1:-1], g[-1, 1:-1] = a[-1, :], a[0, :]
    return g


def same(ld, name):
    got, exp = ld.get_global(), Z[name]
    assert got.shape == exp.shape, (name, got.shape, exp.shape)
    assert np.array_equal(got, exp), f"{name}: max abs diff {np.nanmax(np.abs(got - exp)):g} (the reference kernel's output, expected bit for bit)"


class Side:
    def __init__(self, ctx):
        self.ctx = ctx
        self.lay = amr.DisjointBoxLayout(ctx, np.array([[0, 0, NX - 1, NY - 1]], dtype=np.int32), (0, 0, NX - 1, NY - 1), (1, 1), None)
        self.F = {}
        for k in ("head", "B", "Pi", "zb", "mask"):
            self.F[k] = self.field(wrap(Z[k]), 1)
        self.F["rhs"] = self.field(Z["rhs"])
        self.F["a"] = self.field(Z["aC"])
        self.F["bX"] = self.field(Z["bX"], 0, XF)
        self.F["bY"] = self.field(Z["bY"], 0, YF)
        self.bc = amr.make_bc((0, 0), (0, 0))
        self.prm = amr.make_params(A=A, omega=OMEGA, nu=NU, cutOffbr=CUT, maxOffbr=MX)
        F = self.F
        self.factory = amr.VCAMRNonLinearPoissonOpFactory().define(ctx, [self.lay], [], DX, self.bc, float(Z["alpha"]), [F["a"]], float(Z["beta"]),
                                                                   [F["bX"]], [F["bY"]], self.prm, [F["B"]], [F["Pi"]], [F["zb"]], [F["mask"]])
        self.op = self.factory.AMRnewOp(0)

    def field(self, a, ng=0, cent=CELL):
        a = np.asarray(a)
        f = amr.LevelData(self.lay, a.shape[0] if a.ndim == 3 else 1, ng, cent)
        f.set_global(a, (-ng, -ng))
        return f

    def new(self, ng=0, cent=CELL, ncomp=1):
        return amr.LevelData(self.lay, ncomp, ng, cent)


@pytest.mark.parametrize("mode", [0, 1, 2, 3, 4, 5])
def test_operator_against_reference_kernels(gpu_ctx, mode):
    s = Side(gpu_ctx)
    op, F = s.op, s.F
    lam = s.new()
    op.lambda_(lam)
    same(lam, "lam")
    gpu_ctx.set_relax_mode(mode)
    try:
        for it in (1, 2):
            op.relax(F["head"], F["rhs"], 1)
            same(F["head"], f"gsrb_iter{it}")
        # the same two iterations in ONE call from the initial state (modes 3-5: one launch of the two-iteration kernels)
        h2 = s.field(wrap(Z["head"]), 1)
        op.relax(h2, F["rhs"], 2)
        same(h2, "gsrb_iter2")
    finally:
        gpu_ctx.set_relax_mode(1)
    res, lof = s.new(), s.new()
    op.residual(res, F["head"], F["rhs"])
    same(res, "residual")
    op.applyOp(lof, F["head"], False)
    same(lof, "applyop")
    resc, phic = op.createCoarser(F["rhs"]), op.createCoarser(F["head"])
    op.restrictResidual(resc, F["head"], None, F["rhs"], False)
    same(resc, "restrict_res")
    op.restrictR(phic, F["head"])
    same(phic, "restrict_r")
    corr = op.createCoarser(F["head"])
    g = np.zeros((NY // 2 + 2, NX // 2 + 2))
    g[1:-1, 1:-1] = Z["prolong_corr"]
    corr.set_global(g, (-1, -1))
    op.prolongIncrement(F["head"], corr)
    same(F["head"], "prolong_out")


def test_twin_kernel_every_level_against_reference_kernels(gpu_ctx):
    """the default relax mode with k_gsrb_twin on every level (tune key 19): the two reference iterations in one launch"""
    s = Side(gpu_ctx)
    gpu_ctx.set_relax_mode(1)
    gpu_ctx.set_tuning(19, 1)
    try:
        assert s.op.smoother_kind() == "k_gsrb_twin"
        s.op.relax(s.F["head"], s.F["rhs"], 2)
        same(s.F["head"], "gsrb_iter2")
    finally:
        gpu_ctx.set_tuning(19, 0)


def test_nonlinear_level_gradient_divergence(gpu_ctx):
    s = Side(gpu_ctx)
    F = s.F
    nl, dnl = s.new(), s.new()
    check(lib().sg_nonlinear_level(C.byref(s.prm), nl.h, dnl.h, F["head"].h, F["B"].h, F["mask"].h, F["Pi"].h, F["zb"].h))
    same(nl, "nl")
    same(dnl, "dnl")
    dx = (C.c_double * 2)(*DX)
    for has in (0, 1):
        gx, gy = s.new(0, XF), s.new(0, YF)
        check(lib().sg_mac_gradient(F["head"].h, F["mask"].h if has else None, dx, gx.h, gy.h))
        same(gx, f"macgrad_x_mask{has}")
        same(gy, f"macgrad_y_mask{has}")
    div = s.new()
    s.op.setToZero(div)
    check(lib().sg_divergence(div.h, s.field(Z["div_ux"], 0, XF).h, s.field(Z["div_uy"], 0, YF).h, dx))
    same(div, "div")


def test_reynolds_face_coefficient_and_picard_kernels(gpu_ctx):
    """COMPUTERE, COMPUTEBCOEFF (with and without cutOffBcoef), COMPUTEQW, COMPUTESCAPROD, COMPUTEDCOEFF, COMPUTEDIFTERM2D,
    COMPUTE_TIMEVARYINGRECHARGE (src/AmrHydroF.ChF:81-373)"""
    s = Side(gpu_ctx)
    L = lib()
    prm = amr.make_params(A=A, omega=OMEGA, nu=NU)
    Re = s.new()
    check(L.sg_compute_re(C.byref(prm), Re.h, s.field(Z["B"]).h, s.field(Z["gradH"]).h))
    same(Re, "Re")
    Bec, Reec, IMec = (s.field(Z[k], 0, XF) for k in ("Bec", "Reec", "IMec"))
    for c in (0, 1):
        p = amr.make_params(A=A, omega=OMEGA, nu=NU, cutOffBcoef=c)
        bc = s.new(0, XF)
        check(L.sg_compute_bcoeff(C.byref(p), Bec.h, Reec.h, IMec.h, bc.h))
        same(bc, f"bcoeff_cut{c}")
    qw = s.new(0, XF)
    check(L.sg_compute_qw(C.byref(prm), Bec.h, Reec.h, s.field(Z["gradHec"], 0, XF).h, qw.h))
    same(qw, "Qw")
    p1, p2 = s.new(0, XF), s.new(0, XF)
    check(L.sg_compute_scaprod(qw.h, s.field(Z["sp_b1"], 0, XF).h, s.field(Z["sp_b2"], 0, XF).h, p1.h, p2.h))
    same(p1, "sp_p1")
    same(p2, "sp_p2")
    for c in (0, 1):
        D = s.new(0, XF)
        check(L.sg_compute_dcoeff(D.h, s.field(Z["MRec"], 0, XF).h, Bec.h, IMec.h, 910.0, c))
        same(D, f"dcoeff_cut{c}")
    dx = (C.c_double * 2)(*DX)
    dterm = s.new()
    check(L.sg_compute_difterm(s.field(wrap(Z["B"]), 1).h, dx, dterm.h, s.field(Z["D0"], 0, XF).h, s.field(Z["D1"], 0, YF).h))
    same(dterm, "difterm")
    rech = s.new()
    check(L.sg_time_varying_recharge(s.field(Z["zs"]).h, rech.h, 4.5, 7.93e-11))
    same(rech, "recharge")


def test_extrap_and_copy_ghost_cells(gpu_ctx):
    """ExtrapGhostCells / CopyGhostCells on cell data (util/ExtrapGhostCells.cpp:94-269 + SIMPLEEXTRAPBC / SIMPLECOPYBC), corners included"""
    lay = amr.DisjointBoxLayout(gpu_ctx, np.array([[0, 0, NX - 1, NY - 1]], dtype=np.int32), (0, 0, NX - 1, NY - 1), (0, 0), None)
    for name, fn in (("extrap", amr.ExtrapGhostCells), ("copy", amr.CopyGhostCells)):
        g = np.zeros((NY + 2, NX + 2))
        g[1:-1, 1:-1] = Z["head"]
        f = amr.LevelData(lay, 1, 1, CELL)
        f.set_global(g, (-1, -1))
        fn(f)
        got = f.download_box(0).reshape(NY + 2, NX + 2)
        exp = Z[f"ghost_{name}"]
        assert np.array_equal(got, exp), f"ghost_{name}: max abs diff {np.abs(got - exp).max():g}"


def test_prolong_2_nl_through_amr_prolong_s2(gpu_ctx):
    """sg_op_AMRProlongS_2 against the reference's PROLONG_2_NL (src/AMRNonLinearPoissonOpF.ChF:646-709) as AMRProlongS_2 calls it
    (src/AMRNonLinearPoissonOp.cpp:1141-1206): one fine box inside the doubly periodic coarse box"""
    bx = [int(v) for v in Z["prolong2_box"]]
    layC = amr.DisjointBoxLayout(gpu_ctx, np.array([[0, 0, NX - 1, NY - 1]], dtype=np.int32), (0, 0, NX - 1, NY - 1), (1, 1), None)
    layF = amr.DisjointBoxLayout(gpu_ctx, np.array([bx], dtype=np.int32), (0, 0, 2 * NX - 1, 2 * NY - 1), (1, 1), None)
    lays = [layC, layF]

    def ones(lay, ng=0, cent=CELL):
        f = amr.LevelData(lay, 1, ng, cent)
        f.upload([np.ones(f.fab_shape(b)) for b in range(len(lay.boxes))])
        return f

    F = {k: [ones(l, ng, c) for l in lays] for k, (ng, c) in dict(a=(0, CELL), bX=(0, XF), bY=(0, YF), B=(1, CELL), Pi=(1, CELL), zb=(1, CELL),
                                                                  mask=(1, CELL)).items()}
    factory = amr.VCAMRNonLinearPoissonOpFactory().define(gpu_ctx, lays, [2], DX, amr.make_bc((0, 0), (0, 0)), 0.0, F["a"], -1.0, F["bX"], F["bY"],
                                                          amr.make_params(A=A, omega=OMEGA, nu=NU, cutOffbr=CUT, maxOffbr=MX), F["B"], F["Pi"],
                                                          F["zb"], F["mask"])
    opC, opF = factory.AMRnewOp(0), factory.AMRnewOp(1)
    coarse = amr.LevelData(layC, 1, 1, CELL)
    coarse.set_global(wrap(Z["prolong2_coarse"]), (-1, -1))
    fine = amr.LevelData(layF, 1, 1, CELL)
    g = np.zeros((2 * NY + 2, 2 * NX + 2))
    g[bx[1] + 1:bx[3] + 2, bx[0] + 1:bx[2] + 2] = Z["prolong2_fine_in"]
    fine.set_global(g, (-1, -1))
    opF.AMRProlongS_2(fine, coarse, opC)
    got = fine.get_global()[bx[1]:bx[3] + 1, bx[0]:bx[2] + 1]
    assert np.array_equal(got, Z["prolong2_out"]), f"PROLONG_2_NL: max abs diff {np.abs(got - Z['prolong2_out']).max():g}"

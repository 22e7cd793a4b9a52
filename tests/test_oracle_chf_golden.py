"""The C oracle against golden vectors of the REFERENCE's own Fortran kernels (tests/golden/chf_kernels.npz, made by executing the
.ChF sources of /root/reference through tools/chf_translate.py -- see tests/golden/make_chf_golden.py; the reference tree is not
needed here).  Bit for bit.  This pins the oracle's kernel arithmetic -- COMPUTENONLINEARTERMS, COMPUTERE, COMPUTEBCOEFF, the
Picard-body kernels, SUMFACESNL, GSRBHELMHOLTZVCNL2D, VCNLCOMPUTE{OP,RES}2D, RESTRICT{RES}VCNL, PROLONGNL, NEWMACGRAD, DIVERGENCE --
to the reference source text; what stays unpinned is the absent Chombo layer around them."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding as ob

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "chf_kernels.npz"))
NX, NY = int(Z["nx"]), int(Z["ny"])
DX = tuple(float(v) for v in Z["dx"])
A, OMEGA, NU, CUT, MX = (float(v) for v in Z["prm"])
CELL, XF, YF = ob.CELL, ob.XFACE, ob.YFACE


def layout(periodic=(1, 1)):
    return ob.Layout(np.array([[0, 0, NX - 1, NY - 1]], dtype=np.int32), (0, 0, NX - 1, NY - 1), periodic)


def field(lay, a, ng=0, cent=CELL):
    """a: valid data [nj, ni] / [ncomp, nj, ni] (ng = 0) or the ghosted array (ng = 1)"""
    a = np.asarray(a)
    f = ob.Field(lay, a.shape[0] if a.ndim == 3 else 1, ng, cent)
    f.set_global(a, (-ng, -ng))
    return f


def wrap(a):
    g = np.zeros((a.shape[0] + 2, a.shape[1] + 2))
    g[1:-1, 1:-1] = a
    g[1:-1, 0], g[1:-1, -1] = a[:, -1], a[:, 0]
    g[0, 1:-1], g[-1, 1:-1] = a[-1, :], a[0, :]
    return g


def dxp():
    a = np.array(DX, dtype=np.float64)
    return a, a.ctypes.data_as(C.POINTER(C.c_double))


def same(got, name):
    exp = Z[name]
    assert got.shape == exp.shape, (name, got.shape, exp.shape)
    assert np.array_equal(got, exp), f"{name}: max abs diff {np.abs(got - exp).max():g} (the reference kernel's output, expected bit for bit)"


def test_nonlinear_terms_every_branch():
    lay = layout()
    assert (Z["mask"] < 0).any() and (Z["B"] < CUT).any() and (Z["B"] > MX).any()
    prm = ob.make_params(A=A, omega=OMEGA, nu=NU, cutOffbr=CUT, maxOffbr=MX)
    nl, dnl = ob.Field(lay, 1, 0), ob.Field(lay, 1, 0)
    f = {k: field(lay, Z[k]) for k in ("head", "B", "mask", "Pi", "zb")}
    ob.lib().orc_compute_nl(C.byref(prm), f["head"].h, f["B"].h, f["mask"].h, f["Pi"].h, f["zb"].h, nl.h, dnl.h)
    same(nl.get_global(), "nl")
    same(dnl.get_global(), "dnl")


def test_reynolds_and_face_coefficient():
    lay = layout()
    prm = ob.make_params(A=A, omega=OMEGA, nu=NU)
    Re = ob.Field(lay, 1, 0)
    ob.lib().orc_compute_re(C.byref(prm), field(lay, Z["B"]).h, field(lay, Z["gradH"]).h, Re.h)
    same(Re.get_global(), "Re")
    for c in (0, 1):
        prm = ob.make_params(A=A, omega=OMEGA, nu=NU, cutOffBcoef=c)
        bc = ob.Field(lay, 1, 0, XF)
        ob.lib().orc_compute_bcoeff(C.byref(prm), field(lay, Z["Bec"], 0, XF).h, field(lay, Z["Reec"], 0, XF).h, field(lay, Z["IMec"], 0, XF).h, bc.h)
        same(bc.get_global(), f"bcoeff_cut{c}")


def test_picard_body_kernels():
    lay = layout()
    prm = ob.make_params(A=A, omega=OMEGA, nu=NU)
    L = ob.lib()
    Bec, Reec, IMec = (field(lay, Z[k], 0, XF) for k in ("Bec", "Reec", "IMec"))
    qw = ob.Field(lay, 1, 0, XF)
    L.orc_compute_qw(C.byref(prm), Bec.h, Reec.h, field(lay, Z["gradHec"], 0, XF).h, qw.h)
    same(qw.get_global(), "Qw")
    p1, p2 = ob.Field(lay, 1, 0, XF), ob.Field(lay, 1, 0, XF)
    L.orc_compute_scaprod(qw.h, field(lay, Z["sp_b1"], 0, XF).h, field(lay, Z["sp_b2"], 0, XF).h, p1.h, p2.h)
    same(p1.get_global(), "sp_p1")
    same(p2.get_global(), "sp_p2")
    for c in (0, 1):
        D = ob.Field(lay, 1, 0, XF)
        L.orc_compute_dcoeff(D.h, field(lay, Z["MRec"], 0, XF).h, Bec.h, IMec.h, 910.0, c)
        same(D.get_global(), f"dcoeff_cut{c}")
    dterm = ob.Field(lay, 1, 0)
    _, p = dxp()
    L.orc_compute_difterm(field(lay, wrap(Z["B"]), 1).h, p, dterm.h, field(lay, Z["D0"], 0, XF).h, field(lay, Z["D1"], 0, YF).h)
    same(dterm.get_global(), "difterm")
    rech = ob.Field(lay, 1, 0)
    L.orc_time_varying_recharge(field(lay, Z["zs"]).h, rech.h, 4.5, 7.93e-11)
    same(rech.get_global(), "recharge")


@pytest.fixture
def periodic_op():
    lay = layout((1, 1))
    prm = ob.make_params(A=A, omega=OMEGA, nu=NU, cutOffbr=CUT, maxOffbr=MX)
    bc = ob.make_bc((0, 0), (0, 0))
    F = {k: field(lay, wrap(Z[k]), 1) for k in ("B", "Pi", "zb", "mask")}
    F["a"] = field(lay, Z["aC"])
    F["bX"] = field(lay, Z["bX"], 0, XF)
    F["bY"] = field(lay, Z["bY"], 0, YF)
    F["head"] = field(lay, wrap(Z["head"]), 1)
    F["rhs"] = field(lay, Z["rhs"])
    op = ob.Op(lay, DX, float(Z["alpha"]), float(Z["beta"]), bc, prm, F["a"], F["bX"], F["bY"], F["B"], F["Pi"], F["zb"], F["mask"])
    return lay, op, F


def test_lambda_gsrb_residual_restrict_prolong(periodic_op):
    """resetLambda + SUMFACESNL; two levelGSRB iterations (exchange, NonLinear_level, GSRBHELMHOLTZVCNL2D per colour); then, on that
    state, residualI, applyOpI, restrictResidual, restrictR and prolongIncrement -- one doubly periodic box, so no BC function enters"""
    lay, op, F = periodic_op
    same(op.lambda_field().get_global(), "lam")
    for it in (1, 2):
        op.relax(F["head"], F["rhs"], 1)
        same(F["head"].get_global(), f"gsrb_iter{it}")
    res, lof = ob.Field(lay, 1, 0), ob.Field(lay, 1, 0)
    op.residual(res, F["head"], F["rhs"])
    same(res.get_global(), "residual")
    op.apply(lof, F["head"], False)
    same(lof.get_global(), "applyop")
    lc = lay.coarsen(2)
    resc, phic = ob.Field(lc, 1, 0), ob.Field(lc, 1, 1)
    op.restrict_residual(resc, F["head"], F["rhs"])
    same(resc.get_global(), "restrict_res")
    op.restrict_r(phic, F["head"])
    same(phic.get_global(), "restrict_r")
    corr = ob.Field(lc, 1, 1)
    g = np.zeros((NY // 2 + 2, NX // 2 + 2))
    g[1:-1, 1:-1] = Z["prolong_corr"]
    corr.set_global(g, (-1, -1))
    op.prolong_increment(F["head"], corr)
    same(F["head"].get_global(), "prolong_out")


def test_mac_gradient_and_divergence():
    lay = layout((1, 1))
    _, p = dxp()
    phi, mask = field(lay, wrap(Z["head"]), 1), field(lay, wrap(Z["mask"]), 1)
    for has in (0, 1):
        gx, gy = ob.Field(lay, 1, 0, XF), ob.Field(lay, 1, 0, YF)
        ob.lib().orc_mac_gradient(phi.h, mask.h if has else None, p, gx.h, gy.h)
        same(gx.get_global(), f"macgrad_x_mask{has}")
        same(gy.get_global(), f"macgrad_y_mask{has}")
    div = ob.Field(lay, 1, 0)
    div.setval(0.0)
    ob.lib().orc_divergence(field(lay, Z["div_ux"], 0, XF).h, field(lay, Z["div_uy"], 0, YF).h, p, div.h)
    same(div.get_global(), "div")


def test_extrap_and_copy_ghost_cells():
    """ExtrapGhostCells / CopyGhostCells on cell data (util/ExtrapGhostCells.cpp:94-269 + SIMPLEEXTRAPBC / SIMPLECOPYBC), corners included"""
    lay = layout((0, 0))
    for name, fn in (("extrap", ob.lib().orc_extrap_ghost), ("copy", ob.lib().orc_copy_ghost)):
        g = np.zeros((NY + 2, NX + 2))
        g[1:-1, 1:-1] = Z["head"]
        f = field(lay, g, 1)
        fn(f.h)
        got = f.fab(0)[0][0]
        exp = Z[f"ghost_{name}"]
        assert np.array_equal(got, exp), f"ghost_{name}: max abs diff {np.abs(got - exp).max():g}"


def test_prolong_2_nl_through_amr_prolong_s2():
    """AMRProlongS_2 (src/AMRNonLinearPoissonOp.cpp:1141-1206) around PROLONG_2_NL (src/AMRNonLinearPoissonOpF.ChF:646-709): one fine box
    inside the doubly periodic coarse box, so the scratch's ghost cells are coarse cells (copyTo) and no BC function enters"""
    bx = [int(v) for v in Z["prolong2_box"]]
    layC = layout()
    layF = ob.Layout(np.array([bx], dtype=np.int32), (0, 0, 2 * NX - 1, 2 * NY - 1), (1, 1))
    bc = ob.make_bc((0, 0), (0, 0))
    prm = ob.make_params(A=A, omega=OMEGA, nu=NU, cutOffbr=CUT, maxOffbr=MX)

    def op(lay, dx):
        f = [ob.Field(lay, 1, 0), ob.Field(lay, 1, 0, XF), ob.Field(lay, 1, 0, YF)] + [ob.Field(lay, 1, 1) for _ in range(4)]
        for x in f:
            x.setval(1.0)
        return ob.Op(lay, dx, 0.0, -1.0, bc, prm, *f)

    opC, opF = op(layC, DX), op(layF, (DX[0] / 2, DX[1] / 2))
    coarse = field(layC, wrap(Z["prolong2_coarse"]), 1)
    fine = ob.Field(layF, 1, 1)
    g = np.zeros((2 * NY + 2, 2 * NX + 2))
    g[bx[1] + 1:bx[3] + 2, bx[0] + 1:bx[2] + 2] = Z["prolong2_fine_in"]
    fine.set_global(g, (-1, -1))
    temp = ob.Field(layF.coarsen(2), 1, 1)
    opF.amr_prolong_s2(fine, coarse, temp, opC)
    got = fine.get_global()[bx[1]:bx[3] + 1, bx[0]:bx[2] + 1]
    same(got, "prolong2_out")

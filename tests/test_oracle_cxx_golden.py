"""The C oracle against golden vectors of the REFERENCE's own C++ cell loops (tests/golden/cxx_kernels.npz, made by executing the
BoxIterator loop bodies of src/AmrHydro.cpp through tools/cxx_translate.py -- see tests/golden/make_cxx_golden.py; the reference tree
is not needed here).  Bit for bit.  This pins the part of the Picard body that the reference writes in C++ rather than Fortran:
Calc_meltingRate (:2175-2252), the right-hand side of the head equation (:3044-3077), CalcRHS_gapHeightFAS (:2070-2171) in its four
mask / implicit variants, the explicit gap-height update (:3394-3408), the moulin quadrature (:1867-2069), getFlux
(src/VCAMRNonLinearPoissonOp.cpp:792-841) and setup_iceMask_EC (src/HydroIBC.cpp:138-184).  The CUDA library is held bit for bit to the same oracle
functions on every configuration by tests/test_gpu_picard.py and tests/test_gpu_amr_picard.py."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import binding as ob

Z = np.load(os.path.join(os.path.dirname(__file__), "golden", "cxx_kernels.npz"))
NX, NY = int(Z["nx"]), int(Z["ny"])
RHO_I, RHO_W, GRAV, G, L, CT, CW, UB0, A, CUT, MX, DIFF, DIST = (float(v) for v in Z["prm"])


def layout():
    return ob.Layout(np.array([[0, 0, NX - 1, NY - 1]], dtype=np.int32), (0, 0, NX - 1, NY - 1), (0, 0))


def ghosted(lay, a):
    a = np.asarray(a)
    f = ob.Field(lay, a.shape[0] if a.ndim == 3 else 1, 1)
    f.set_global(a, (-1, -1))
    return f


def params(**over):
    kw = dict(rho_i=RHO_I, rho_w=RHO_W, gravity=GRAV, G=G, L=L, ct=CT, cw=CW, ub0=UB0, basal_friction=1, A=A, cutOffbr=CUT, maxOffbr=MX,
              DiffFactor=DIFF, n_moulins=3, ramp=float(Z["ramp"]), distributed_input=DIST, use_mask_rhs_b=0, use_ImplDiff=0)
    kw.update(over)
    return ob.PicardParams(**kw)


def same(got, name):
    exp = Z[name]
    assert got.shape == exp.shape, (name, got.shape, exp.shape)
    assert np.array_equal(got, exp), f"{name}: max abs diff {np.abs(got - exp).max():g} (the reference loop's output, expected bit for bit)"


def whole(f):
    return f.fab(0)[0][0].copy()


@pytest.fixture
def fields():
    lay = layout()
    F = {k: ghosted(lay, Z[k]) for k in ("H", "zb", "Pi", "IM", "B", "qgh", "qgz", "MV", "BH", "BL", "MS")}
    F["Dterm"] = ob.Field(lay, 1, 0)
    F["Dterm"].set_global(Z["Dterm"], (0, 0))
    F["lay"] = lay
    return F


def test_melting_rate_over_the_ghosted_array(fields):
    F, q = fields, params()
    Pw, mR = ob.Field(F["lay"], 1, 1), ob.Field(F["lay"], 1, 1)
    ob.lib().orc_calc_melting_rate(C.byref(q), F["H"].h, F["zb"].h, F["Pi"].h, F["IM"].h, F["B"].h, F["qgh"].h, F["qgz"].h, Pw.h, mR.h)
    same(whole(Pw), "Pw")
    same(whole(mR), "mR")
    m = Z["mR"]
    assert (m == 0).any() and (m > 0).any() and (Z["B"] < 1e-6).any() and (Z["IM"] < 0).any()


@pytest.mark.parametrize("tag,nm", [("moulins", 3), ("distributed", -1)])
def test_rhs_head(fields, tag, nm):
    F, q = fields, params(n_moulins=nm, ramp=float(Z["ramp"]) if nm > 0 else 1.0)
    mR = ghosted(F["lay"], Z["mR"])
    rhs = ob.Field(F["lay"], 1, 0)
    ob.lib().orc_rhs_head(C.byref(q), rhs.h, mR.h, F["B"].h, F["BH"].h, F["BL"].h, F["MV"].h, F["MS"].h, F["Dterm"].h, F["IM"].h)
    same(rhs.get_global(), "rhs_head_" + tag)


@pytest.mark.parametrize("mask_rhs", [0, 1])
@pytest.mark.parametrize("impl", [0, 1])
def test_rhs_gap_and_explicit_update(fields, mask_rhs, impl):
    F, q = fields, params(use_mask_rhs_b=mask_rhs, use_ImplDiff=impl)
    mR, Pw = ghosted(F["lay"], Z["mR"]), ghosted(F["lay"], Z["Pw"])
    rhs = ob.Field(F["lay"], 1, 0)
    dt = float(Z["dt"])
    ob.lib().orc_rhs_gap(C.byref(q), rhs.h, F["Pi"].h, Pw.h, mR.h, F["B"].h, F["Dterm"].h, F["IM"].h, F["BH"].h, F["BL"].h, F["MV"].h, dt)
    same(rhs.get_global(), f"rhs_gap_mask{mask_rhs}_impl{impl}")
    if not mask_rhs and not impl:
        newB = ob.Field(F["lay"], 1, 1)
        ob.lib().orc_gap_euler(newB.h, F["B"].h, rhs.h, dt)
        same(newB.get_global(), "gap_euler")


def test_moulin_recharge_one_level():
    """Calc_moulin_integral + Calc_moulin_source_term_distributed (src/AmrHydro.cpp:1867-2069) on one level: the nine-point Gauss-Legendre
    values per moulin, their integrals and the normalised source term -- exp() and the order of the sums included"""
    lo = [int(v) for v in Z["moulin_lo"]]
    lay = ob.Layout(np.array([[lo[0], lo[1], lo[0] + NX - 1, lo[1] + NY - 1]], dtype=np.int32), (0, 0, 63, 63), (0, 0))
    n = len(Z["moulin_sigma"])
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))  # noqa: E731
    dx, pos, sig, flux = (np.ascontiguousarray(Z[k], dtype=np.float64) for k in ("moulin_dx", "moulin_pos", "moulin_sigma", "moulin_flux"))
    tmp, src = ob.Field(lay, n, 0), ob.Field(lay, 1, 0)
    ob.lib().orc_moulin_nonorm(tmp.h, dp(dx), n, dp(pos), dp(sig))
    got = tmp.fab(0)[0].copy()
    assert np.array_equal(got, Z["moulin_nonorm"]), np.abs(got - Z["moulin_nonorm"]).max()
    integ = np.zeros(n)
    ob.lib().orc_moulin_integral(tmp.h, None, dp(dx), n, dp(integ))
    assert np.array_equal(integ, Z["moulin_integral"]), (integ, Z["moulin_integral"])
    ob.lib().orc_moulin_source(src.h, tmp.h, n, dp(integ), dp(flux), float(Z["moulin_runoff"]), float(Z["moulin_time"]))
    got = src.fab(0)[0][0].copy()
    assert np.array_equal(got, Z["moulin_source"]), np.abs(got - Z["moulin_source"]).max()


@pytest.mark.parametrize("d", [0, 1])
@pytest.mark.parametrize("ref", [1, 2])
def test_get_flux(d, ref):
    """VCAMRNonLinearPoissonOp::getFlux (src/VCAMRNonLinearPoissonOp.cpp:792-841): flux = -bCoef * ((phi_hi - phi_lo) * (beta * ref / dx))
    on every face of the box, the evaluation order of the reference's statements"""
    lay = layout()
    dx = tuple(float(v) for v in Z["flux_dx"])
    one = lambda ng=0, cent=ob.CELL: ob.Field(lay, 1, ng, cent)  # noqa: E731
    bX, bY = one(0, ob.XFACE), one(0, ob.YFACE)
    bX.set_global(Z["flux_b0"], (0, 0))
    bY.set_global(Z["flux_b1"], (0, 0))
    op = ob.Op(lay, dx, 0.0, float(Z["flux_beta"]), ob.make_bc((0, 0), (0, 0)), ob.make_params(), one(), bX, bY, one(1), one(1), one(1), one(1))
    phi = ghosted(lay, Z["flux_phi"])
    flux = one(0, ob.XFACE if d == 0 else ob.YFACE)
    ob.lib().orc_op_get_flux(op.h, flux.h, phi.h, d, ref, 1.0)
    same(flux.fab(0)[0][0].copy(), f"flux_dir{d}_ref{ref}")


def test_icemask_ec():
    """HydroIBC::setup_iceMask_EC (src/HydroIBC.cpp:138-184): +1 / -1 where both cells agree, 0 across an ice edge and on the domain faces"""
    lo, dom = [int(v) for v in Z["imec_lo"]], tuple(int(v) for v in Z["imec_domain"])
    lay = ob.Layout(np.array([[lo[0], lo[1], lo[0] + NX - 1, lo[1] + NY - 1]], dtype=np.int32), dom, (0, 0))
    mask = ob.Field(lay, 1, 1)
    mask.fab(0)[0][0][...] = Z["imec_mask"]
    mx, my = ob.Field(lay, 1, 0, ob.XFACE), ob.Field(lay, 1, 0, ob.YFACE)
    ob.lib().orc_icemask_ec(mask.h, mx.h, my.h)
    same(mx.fab(0)[0][0].copy(), "imec_dir0")
    same(my.fab(0)[0][0].copy(), "imec_dir1")

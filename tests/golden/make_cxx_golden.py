"""Golden vectors of the REFERENCE's own C++ cell loops, made by executing the `BoxIterator` loop bodies of src/AmrHydro.cpp through
tools/cxx_translate.py (this container only: the reference tree does not travel).  Output: tests/golden/cxx_kernels.npz, which
tests/test_oracle_cxx_golden.py holds the C oracle to, bit for bit.

    python tests/golden/make_cxx_golden.py            # rewrites the .npz

One box of 12 x 10 cells, seeded random fields of the magnitudes the solver sees, chosen so that every branch of every loop is taken.
What stands between the loop bodies and the arrays is stated at each case (which box the BoxIterator runs over, what the FArrayBox
calls before the loop do).
"""
import os
import re
import sys
from types import SimpleNamespace

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools import cxx_translate as X  # noqa: E402

REF = "/root/reference/src/AmrHydro.cpp"
NX, NY = 12, 10


def ghosted(rng, lo, hi, ncomp=1):
    return lo + (hi - lo) * rng.rand(ncomp, NY + 2, NX + 2)


def main():
    src = open(REF).read()
    rng = np.random.RandomState(20261019)
    # suhmo_params as the inputs set them (src/suhmo_params.cpp:51-74); cutOffbr / maxOffbr inside the range of the gap height
    P = SimpleNamespace(m_rho_i=910.0, m_rho_w=1000.0, m_gravity=9.8, m_G=0.05, m_L=334000.0, m_ct=7.5e-8, m_cw=4220.0, m_ub=[1.0e-6, 0.0],
                        m_basal_friction=True, m_A=5.0e-25, m_cutOffbr=0.008, m_maxOffbr=0.012, m_DiffFactor=1.0e-2, m_n_moulins=3,
                        m_distributed_input=3.0e-9)
    out = {"nx": NX, "ny": NY, "prm": np.array([P.m_rho_i, P.m_rho_w, P.m_gravity, P.m_G, P.m_L, P.m_ct, P.m_cw, P.m_ub[0], P.m_A, P.m_cutOffbr,
                                                  P.m_maxOffbr, P.m_DiffFactor, P.m_distributed_input])}
    H = ghosted(rng, 1200.0, 1500.0)
    zb = ghosted(rng, 0.0, 100.0)
    Pi = ghosted(rng, 4.5e6, 1.35e7)
    IM = np.where(rng.rand(1, NY + 2, NX + 2) < 0.2, -1.0, 1.0)
    B = ghosted(rng, 0.005, 0.015)
    B[0][rng.rand(NY + 2, NX + 2) < 0.1] = 5.0e-7          # the `B < 1e-6` clause of Calc_meltingRate
    qgh = ghosted(rng, -2.0e-6, 2.0e-6, 2)
    qgz = ghosted(rng, -2.0e-6, 2.0e-6, 2)
    MV = ghosted(rng, 5.0e-7, 2.0e-6)
    BH = ghosted(rng, 0.008, 0.014)
    BL = ghosted(rng, 1.5, 2.5)
    MS = ghosted(rng, 0.0, 4.0e-8)
    Dterm = 1.0e-9 * (rng.rand(1, NY, NX) - 0.5)
    out.update(H=H[0], zb=zb[0], Pi=Pi[0], IM=IM[0], B=B[0], qgh=qgh, qgz=qgz, MV=MV[0], BH=BH[0], BL=BL[0], MS=MS[0], Dterm=Dterm[0])

    # ---- 1. Calc_meltingRate (src/AmrHydro.cpp:2175-2252): BoxIterator over Pressw.box(), i.e. the ghosted array
    body = X.box_loops(X.strip_comments(X.function_text(src, "AmrHydro::Calc_meltingRate")))
    assert len(body) == 1
    cell = X.compile_cell(body[0], ["Pressw", "newH", "zb", "MV", "Pressi", "tmp_cc", "tmp2_cc", "B", "mR", "IM", "mMR_A", "mMR_B", "mMR_C"])
    z = lambda n=1, g=1: np.zeros((n, NY + 2 * g, NX + 2 * g))  # noqa: E731
    A = dict(Pressw=z(), newH=H, zb=zb, MV=MV, Pressi=Pi, tmp_cc=qgh, tmp2_cc=qgz, B=B, mR=z(), IM=IM, mMR_A=z(), mMR_B=z(), mMR_C=z())
    X.run_box(cell, A, P, {}, NY + 2, NX + 2)
    Pw, mR = A["Pressw"], A["mR"]
    assert (mR[0] == 0).any() and (mR[0] > 0).any()
    out.update(Pw=Pw[0], mR=mR[0])

    # ---- 2. right-hand side of the head equation (the loop inside timeStepFAS, src/AmrHydro.cpp:3044-3077): BoxIterator over
    # RHSh.box() (no ghost cells); rho_coef (:3023), ramp (:2448) are locals of timeStepFAS
    ts = X.function_text(src, "AmrHydro::timeStepFAS")
    loops = X.box_loops(X.strip_comments(ts))
    rhs_h = [b for b in loops if "rho_coef" in b]
    assert len(rhs_h) == 1
    cell = X.compile_cell(rhs_h[0], ["B", "RHSh", "bumpHeight", "bumpSpacing", "DiffusiveTerm", "MV", "mR", "IM", "moulinSrc"])
    inner = lambda a: a[:, 1:-1, 1:-1].copy()  # noqa: E731
    rho_coef = (1.0 / P.m_rho_w - 1.0 / P.m_rho_i)
    for tag, nm, ramp in (("moulins", 3, 0.6), ("distributed", -1, 1.0)):
        P.m_n_moulins = nm
        A = dict(B=inner(B), RHSh=z(1, 0), bumpHeight=inner(BH), bumpSpacing=inner(BL), DiffusiveTerm=Dterm, MV=inner(MV), mR=inner(mR),
                 IM=inner(IM), moulinSrc=inner(MS))
        X.run_box(cell, A, P, dict(rho_coef=rho_coef, ramp=ramp), NY, NX)
        out["rhs_head_" + tag] = A["RHSh"][0]
    out["ramp"] = np.array(0.6)

    # ---- 3. CalcRHS_gapHeightFAS (src/AmrHydro.cpp:2070-2171): before the loop RHS and RHS_A are copies of the melt rate scaled by
    # 1/rho_i (FArrayBox::copy, operator*=), RHS_B and RHS_C zero; BoxIterator over RHS.box() (no ghost cells)
    body = X.box_loops(X.strip_comments(X.function_text(src, "AmrHydro::CalcRHS_gapHeightFAS")))
    assert len(body) == 1
    cell = X.compile_cell(body[0], ["B", "RHS", "DT", "RHS_A", "RHS_B", "RHS_C", "CD", "Pressi", "IM", "Pw", "meltR", "BH", "BL", "MV"])
    dt = 1800.0
    out["dt"] = np.array(dt)
    for mask_rhs in (0, 1):
        for impl in (0, 1):
            r0 = inner(mR) * (1.0 / P.m_rho_i)
            A = dict(B=inner(B), RHS=r0.copy(), DT=Dterm, RHS_A=r0.copy(), RHS_B=z(1, 0), RHS_C=z(1, 0), CD=z(1, 0), Pressi=inner(Pi), IM=inner(IM),
                     Pw=inner(Pw), meltR=inner(mR), BH=inner(BH), BL=inner(BL), MV=inner(MV))
            with np.errstate(all="ignore"):   # the channelisation degree CD = RHS_A / (RHS_A + RHS_B) is 0/0 where there is no melt
                X.run_box(cell, A, P, dict(m_use_mask_rhs_b=bool(mask_rhs), m_use_ImplDiff=bool(impl), a_dt=dt), NY, NX)
            out[f"rhs_gap_mask{mask_rhs}_impl{impl}"] = A["RHS"][0]
    g = out["rhs_gap_mask0_impl0"]
    Bi = inner(B)[0]
    assert (Bi < P.m_cutOffbr).any() and (Bi > P.m_maxOffbr).any() and ((Bi >= P.m_cutOffbr) & (Bi <= P.m_maxOffbr)).any() and np.isfinite(g).all()

    # ---- 4. explicit gap-height update (src/AmrHydro.cpp:3394-3408): BoxIterator over RHS.box()
    eul = [b for b in loops if "oldB" in b and "a_dt" in b]
    assert len(eul) == 1
    cell = X.compile_cell(eul[0], ["oldB", "newB", "RHS"])
    A = dict(oldB=inner(B), newB=z(1, 0), RHS=out["rhs_gap_mask0_impl0"][None].copy())
    X.run_box(cell, A, P, dict(a_dt=dt), NY, NX)
    out["gap_euler"] = A["newB"][0]

    # ---- 5. moulin recharge on one level (src/AmrHydro.cpp:1867-2069): the Gauss-Legendre loop of Calc_moulin_integral over
    # moulinSrcTmp.box() with the weights and nodes declared at the top of the function, its normalisation loop (sum over the box, i
    # fastest, moulin innermost), and the loop of Calc_moulin_source_term_distributed.  No finer level: the zeroing of covered cells in
    # between is Chombo box calculus, not arithmetic.
    mi = X.strip_comments(X.function_text(src, "AmrHydro::Calc_moulin_integral"))
    consts = X.real_decls(mi[:mi.index("for (int lev")])
    assert set(consts) == {"v_m1", "v_c1", "v_p1", "l_m1", "l_c1", "l_p1"}, consts
    lo = (16, 24)                                          # the box sits away from the origin: iv[0], iv[1] are absolute indices
    dx = [[75.0, 82.0], [37.5, 41.0]]                      # m_amrDx of two levels; the loops run on lev = 1
    P.m_n_moulins = 2
    P.m_sigma = [70.0, 110.0]
    P.m_moulin_position = [(lo[0] + 3.3) * dx[1][0], (lo[1] + 6.1) * dx[1][1], (lo[0] + 9.7) * dx[1][0], (lo[1] + 2.4) * dx[1][1]]
    P.m_moulin_flux = [30.0, 12.0]
    P.m_runoff = 0.3
    loops_mi = X.box_loops(mi)
    assert len(loops_mi) == 2
    tmp = np.zeros((2, NY, NX))
    X.run_box(X.compile_cell(loops_mi[0], ["moulinSrcTmp"]), dict(moulinSrcTmp=tmp), P, dict(consts, m_amrDx=dx, lev=1), NY, NX, lo)
    integ = [0.0, 0.0]
    X.run_box(X.compile_cell(loops_mi[1], ["moulinSrcTmp"]), dict(moulinSrcTmp=tmp), P, dict(m_amrDx=dx, lev=1, a_moulinsInteg=integ), NY, NX, lo)
    sd = X.strip_comments(X.function_text(src, "AmrHydro::Calc_moulin_source_term_distributed"))
    loops_sd = X.box_loops(sd)
    assert len(loops_sd) == 1
    ms = np.zeros((1, NY, NX))
    final = [0.0, 0.0]
    time = 7200.0
    X.run_box(X.compile_cell(loops_sd[0], ["moulinSrcTmp", "moulinSrc"]), dict(moulinSrcTmp=tmp, moulinSrc=ms), P,
              dict(m_amrDx=dx, curr_level=1, a_moulinsInteg=integ, a_moulinsIntegFinal=final, Pi=3.14159265358979323846, m_time=time, m_restart_time=0.0),
              NY, NX, lo)
    assert tmp.max() > 1e-4 and all(v > 0 for v in integ) and abs(sum(final) - (30.0 + 12.0) * max(1.0 - 0.3 * np.sin(2 * np.pi * time / 86400.0), 0.0)) < 1e-9
    out.update(moulin_lo=np.array(lo, dtype=np.int32), moulin_dx=np.array(dx[1]), moulin_pos=np.array(P.m_moulin_position), moulin_sigma=np.array(P.m_sigma),
               moulin_flux=np.array(P.m_moulin_flux), moulin_runoff=np.array(P.m_runoff), moulin_time=np.array(time), moulin_nonorm=tmp,
               moulin_integral=np.array(integ), moulin_source=ms[0])

    # ---- 6. VCAMRNonLinearPoissonOp::getFlux (src/VCAMRNonLinearPoissonOp.cpp:792-841), what reflux evaluates on both sides of a
    # coarse-fine face: BoxIterator over the face box, data with one ghost cell, `scale` from the statement just above the loop
    vc = open("/root/reference/src/VCAMRNonLinearPoissonOp.cpp").read()
    gf = X.strip_comments(X.function_text(vc, "void VCAMRNonLinearPoissonOp::getFlux"))
    loops_gf = X.box_loops(gf)
    assert len(loops_gf) == 1
    scale_expr = re.search(r"Real\s+scale\s*=\s*([^;]+);", gf).group(1)
    cell = X.compile_cell(loops_gf[0], ["a_data", "a_flux", "bCoefDir"])
    phi = ghosted(rng, 1200.0, 1500.0)
    dxv = [37.5, 41.0]
    out.update(flux_phi=phi[0], flux_dx=np.array(dxv), flux_beta=np.array(-1.0))
    for d in (0, 1):
        bco = -(1e-3 + 1e-3 * rng.rand(1, NY + (d == 1), NX + (d == 0)))
        out[f"flux_b{d}"] = bco[0]
        for ref in (1, 2):
            scale = eval(scale_expr, {"m_beta": -1.0, "a_ref": ref, "m_dx_vect": dxv, "a_dir": d})
            fl = np.zeros_like(bco)
            X.run_box(cell, dict(a_data=X.Fab(phi, (-1, -1)), a_flux=fl, bCoefDir=bco), P, dict(a_dir=d, scale=scale), NY + (d == 1), NX + (d == 0))
            out[f"flux_dir{d}_ref{ref}"] = fl[0]

    # ---- 7. HydroIBC::setup_iceMask_EC (src/HydroIBC.cpp:138-184), the edge-centred ice mask WFlx_level multiplies bCoef with: a box on
    # the low-x side of a larger domain, the cell mask with one ghost cell
    ib = X.strip_comments(X.function_text(open("/root/reference/src/HydroIBC.cpp").read(), "HydroIBC::setup_iceMask_EC"))
    loops_im = X.box_loops(ib)
    assert len(loops_im) == 1
    cell = X.compile_cell(loops_im[0], ["thisIM", "thisIMEC_dir"])
    blo, dom = (0, 2), (0, 0, 23, 15)
    imask = np.where(rng.rand(1, NY + 2, NX + 2) < 0.3, -1.0, 1.0)
    out.update(imec_lo=np.array(blo, dtype=np.int32), imec_domain=np.array(dom, dtype=np.int32), imec_mask=imask[0])
    for d in (0, 1):
        ec = np.full((1, NY + (d == 1), NX + (d == 0)), 7.0)
        face_box = SimpleNamespace(smallEnd=lambda k, _d=d: dom[k], bigEnd=lambda k, _d=d: dom[2 + k] + (1 if k == _d else 0))
        X.run_box(cell, dict(thisIM=X.Fab(imask, (blo[0] - 1, blo[1] - 1)), thisIMEC_dir=ec), P, dict(dir=d, face_box=face_box),
                  NY + (d == 1), NX + (d == 0), blo)
        assert set(np.unique(ec)) == {-1.0, 0.0, 1.0} and (ec[0][:, 0] == 0).all() == (d == 0)
        out[f"imec_dir{d}"] = ec[0]

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "cxx_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays")


if __name__ == "__main__":
    main()

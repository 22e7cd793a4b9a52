"""Golden vectors of the REFERENCE's own Fortran kernels, made by executing the .ChF sources under /root/reference through
tools/chf_translate.py (this container only: the reference tree does not travel).  Output: tests/golden/chf_kernels.npz, which
tests/test_oracle_chf_golden.py holds the C oracle to, bit for bit.

    python tests/golden/make_chf_golden.py            # rewrites the .npz

One box of 12 x 10 cells, seeded random fields of the magnitudes the solver sees.  The operator cases run on a DOUBLY PERIODIC
domain, where the ghost cells of a single box are its own wrapped cells: no boundary-condition function (absent Chombo, an
inferred piece) enters, only the kernels and the few lines of C++ that call them (cited at each case).
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from tools import chf_translate as T  # noqa: E402

REF = "/root/reference"
NX, NY = 12, 10
DX = (37.5, 41.0)


def fab(a, lo=(0, 0)):
    """numpy [ncomp, nj, ni] (or [nj, ni] -> one component) as a Fortran array argument with lower corner lo"""
    a = a if a.ndim == 3 else a[None]
    return T.Fab(a, lo)


def wrap(a):
    """a[NY, NX] -> ghosted [NY+2, NX+2] with the periodic images in the face strips (Copier::exchange of one periodic box;
    the corners, which no 5-point kernel reads, are left zero)"""
    g = np.zeros((a.shape[0] + 2, a.shape[1] + 2))
    g[1:-1, 1:-1] = a
    g[1:-1, 0], g[1:-1, -1] = a[:, -1], a[:, 0]
    g[0, 1:-1], g[-1, 1:-1] = a[-1, :], a[0, :]
    return g


def main():
    K = {}
    K.update(T.load(f"{REF}/src/AmrHydroF.ChF"))
    K.update(T.load(f"{REF}/src/VCAMRNonLinearPoissonOpF.ChF"))
    K.update(T.load(f"{REF}/src/AMRNonLinearPoissonOpF.ChF", names={"PROLONGNL", "PROLONG_2_NL"}))
    K.update(T.load(f"{REF}/util/GradientF.ChF", names={"NEWMACGRAD"}))
    K.update(T.load(f"{REF}/util/DivergenceF.ChF"))
    K.update(T.load(f"{REF}/util/ExtrapBCF.ChF", names={"SIMPLEEXTRAPBC", "SIMPLECOPYBC"}))
    rng = np.random.RandomState(20261018)
    box = T.Box((0, 0), (NX - 1, NY - 1))
    out = {"nx": NX, "ny": NY, "dx": np.array(DX)}

    # ---- fields of solver-like magnitude
    head = 1200.0 + 300.0 * rng.rand(NY, NX)
    B = 0.005 + 0.01 * rng.rand(NY, NX)
    Pi = 9.0e6 * (0.5 + rng.rand(NY, NX))
    zb = 100.0 * rng.rand(NY, NX)
    mask = np.where(rng.rand(NY, NX) < 0.2, -1.0, 1.0)
    rhs = 1e-6 * (rng.rand(NY, NX) - 0.3)
    A, omega, nu = 5.0e-25, 1.0e-3, 1.787e-6
    cut, mx = 0.008, 0.012                          # cutOffbr / maxOffbr inside the range of B: every branch of COMPUTENONLINEARTERMS
    out.update(head=head, B=B, Pi=Pi, zb=zb, mask=mask, rhs=rhs, prm=np.array([A, omega, nu, cut, mx]))

    # ---- 1. COMPUTENONLINEARTERMS (src/AmrHydroF.ChF:23-68), called on the valid box by NonLinear_level (src/AmrHydro.cpp:1542-1574)
    nl, dnl = np.zeros((1, NY, NX)), np.zeros((1, NY, NX))
    K["COMPUTENONLINEARTERMS"](phi=fab(head), ab=fab(B), im=fab(mask), api=fab(Pi), azb=fab(zb), region=box, nlfunc=fab(nl), dnlfunc=fab(dnl),
                               aparam=A, brparam=cut, brparammax=mx)
    out.update(nl=nl[0], dnl=dnl[0])

    # ---- 2. COMPUTERE (src/AmrHydroF.ChF:81-112)
    grad = 1e-2 * (rng.rand(2, NY, NX) - 0.5)
    Re = np.zeros((1, NY, NX))
    K["COMPUTERE"](ab=fab(B), agradh=fab(grad), region=box, re=fab(Re), omegaparam=omega, nuparam=nu)
    out.update(gradH=grad, Re=Re[0])

    # ---- 3. COMPUTEBCOEFF (src/AmrHydroF.ChF:199-231) on the x-face box, with and without cutOffBcoef
    fbox = T.Box((0, 0), (NX, NY - 1))
    Bec, Reec = 0.005 + 0.01 * rng.rand(NY, NX + 1), 2000.0 * rng.rand(NY, NX + 1)
    IMec = rng.choice([-1.0, 0.0, 1.0], size=(NY, NX + 1))
    out.update(Bec=Bec, Reec=Reec, IMec=IMec)
    for c in (0, 1):
        bc = np.zeros((1, NY, NX + 1))
        K["COMPUTEBCOEFF"](ab=fab(Bec), are=fab(Reec), region=fbox, bcoeff=fab(bc), imec=fab(IMec), omegaparam=omega, nuparam=nu, cutoffb=c)
        out[f"bcoeff_cut{c}"] = bc[0]

    # ---- 4. Picard-body kernels (src/AmrHydroF.ChF:125-373)
    gH = 1e-2 * (rng.rand(NY, NX + 1) - 0.5)
    qw = np.zeros((1, NY, NX + 1))
    K["COMPUTEQW"](ab=fab(Bec), are=fab(Reec), agradh=fab(gH), region=fbox, qw=fab(qw), omegaparam=omega, nuparam=nu)
    p1, p2 = np.zeros((1, NY, NX + 1)), np.zeros((1, NY, NX + 1))
    v1, v2 = rng.rand(NY, NX + 1), rng.rand(NY, NX + 1)
    K["COMPUTESCAPROD"](vara=fab(qw[0].copy()), var1b=fab(v1), var2b=fab(v2), region=fbox, prod1=fab(p1), prod2=fab(p2))
    out.update(gradHec=gH, Qw=qw[0], sp_b1=v1, sp_b2=v2, sp_p1=p1[0], sp_p2=p2[0])
    MRec = 1e-4 * rng.rand(NY, NX + 1)
    for c in (0, 1):
        D = np.zeros((1, NY, NX + 1))
        K["COMPUTEDCOEFF"](region=fbox, dcoeff=fab(D), dx=DX, rho=910.0, mrec=fab(MRec), bec=fab(Bec), imec=fab(IMec), cutoffb=c)
        out[f"dcoeff_cut{c}"] = D[0]
    out.update(MRec=MRec)
    # COMPUTEDIFTERM2D: div(D grad B) on the valid cells, ghost cells of B by periodic wrap
    D0, D1 = 1e-5 * rng.rand(NY, NX + 1), 1e-5 * rng.rand(NY + 1, NX)
    Bg = wrap(B)
    dterm = np.zeros((1, NY, NX))
    K["COMPUTEDIFTERM2D"](phi=fab(Bg, (-1, -1)), region=box, dx=DX, dterm=fab(dterm), dcoef0=fab(D0), dcoef1=fab(D1))
    out.update(D0=D0, D1=D1, difterm=dterm[0])
    zs = 3000.0 * rng.rand(NY, NX)
    rech = np.zeros((1, NY, NX))
    K["COMPUTE_TIMEVARYINGRECHARGE"](azs=fab(zs), region=box, recharge=fab(rech), tk=4.5, backgroundinput=7.93e-11)
    out.update(zs=zs, recharge=rech[0])

    # ---- 5. the operator on a doubly periodic box: alpha != 0 so that aCoef matters too
    alpha, beta = 0.75, -1.0
    aC = 1e-9 * (1.0 + rng.rand(NY, NX))
    bX = -(1e-3 + 1e-3 * rng.rand(NY, NX + 1))
    bY = -(1e-3 + 1e-3 * rng.rand(NY + 1, NX))
    bX[:, -1] = bX[:, 0]    # periodic: the face on the high boundary is the image of face 0
    bY[-1, :] = bY[0, :]
    out.update(alpha=alpha, beta=beta, aC=aC, bX=bX, bY=bY)
    # resetLambda (src/VCAMRNonLinearPoissonOp.cpp:505-534): lambda = aCoef * alpha, then SUMFACESNL per direction with 1/dx^2
    lam = (aC * alpha)[None].copy()
    for d, b in ((0, bX), (1, bY)):
        K["SUMFACESNL"](lhs=fab(lam), beta=beta, bcoefs=fab(b), box=box, dir=d, scale=1.0 / (DX[d] * DX[d]))
    out.update(lam=lam[0])

    def nonlinear(phi_valid):
        n_, d_ = np.zeros((1, NY, NX)), np.zeros((1, NY, NX))
        K["COMPUTENONLINEARTERMS"](phi=fab(phi_valid), ab=fab(B), im=fab(mask), api=fab(Pi), azb=fab(zb), region=box, nlfunc=fab(n_),
                                   dnlfunc=fab(d_), aparam=A, brparam=cut, brparammax=mx)
        return n_, d_

    # levelGSRB (src/VCAMRNonLinearPoissonOp.cpp:654-760): per colour exchange, [BC: none on periodic sides], NonLinear_level, GSRBHELMHOLTZVCNL2D
    phi = head.copy()
    for it in range(2):
        for colour in (0, 1):
            pg = wrap(phi)
            n_, d_ = nonlinear(phi)
            K["GSRBHELMHOLTZVCNL2D"](phi=fab(pg, (-1, -1)), rhs=fab(rhs), region=box, dx=DX, alpha=alpha, acoef=fab(aC), beta=beta,
                                     bcoef0=fab(bX), bcoef1=fab(bY), nlfunc=fab(n_), nldfunc=fab(d_), **{"lambda": fab(lam)}, redblack=colour)
            phi = pg[1:-1, 1:-1].copy()
        out[f"gsrb_iter{it + 1}"] = phi.copy()
    # residualI / applyOpI (src/VCAMRNonLinearPoissonOp.cpp:98-167, 273-345): exchange, NonLinear_level, VCNLCOMPUTE{RES,OP}2D
    pg = wrap(phi)
    n_, d_ = nonlinear(phi)
    res, lof = np.zeros((1, NY, NX)), np.zeros((1, NY, NX))
    common = dict(phi=fab(pg, (-1, -1)), alpha=alpha, acoef=fab(aC), beta=beta, bcoef0=fab(bX), bcoef1=fab(bY), nlfunc=fab(n_), region=box, dx=DX)
    K["VCNLCOMPUTERES2D"](res=fab(res), rhs=fab(rhs), **common)
    K["VCNLCOMPUTEOP2D"](lofphi=fab(lof), **common)
    out.update(residual=res[0], applyop=lof[0])
    # restrictResidual / restrictR (src/VCAMRNonLinearPoissonOp.cpp:347-460): coarse arrays set to zero, then the kernels over the FINE box
    resc, phic = np.zeros((1, NY // 2, NX // 2)), np.zeros((1, NY // 2, NX // 2))
    K["RESTRICTRESVCNL2D"](res=fab(resc), phi=fab(pg, (-1, -1)), rhs=fab(rhs), alpha=alpha, acoef=fab(aC), beta=beta, bcoef0=fab(bX),
                           bcoef1=fab(bY), nlfunc=fab(n_), region=box, dx=DX)
    K["RESTRICTVCNL"](phicoarse=fab(phic), phifine=fab(phi), region=box, dx=DX[0])
    out.update(restrict_res=resc[0], restrict_r=phic[0])
    # prolongIncrement (src/AMRNonLinearPoissonOp.cpp:856-886): PROLONGNL over the fine box, m = 2
    corr = 1e-2 * (rng.rand(NY // 2, NX // 2) - 0.5)
    pf = phi[None].copy()
    K["PROLONGNL"](phi=fab(pf), coarse=fab(corr), region=box, m=2)
    out.update(prolong_corr=corr, prolong_out=pf[0])

    # ---- 6. NEWMACGRAD, normal derivative (util/GradientF.ChF:30-88; Gradient::singleBoxMacGrad util/Gradient.cpp:250-330), with and without mask
    hg, mg = wrap(head), wrap(mask)
    for has in (0, 1):
        gx, gy = np.zeros((NY, NX + 1)), np.zeros((NY + 1, NX))
        K["NEWMACGRAD"](edgegrad=T.Fab(gx, (0, 0), one=True), mask=T.Fab(mg, (-1, -1), one=True), phi=T.Fab(hg, (-1, -1), one=True),
                        edgegrid=T.Box((0, 0), (NX, NY - 1)), dx=DX, dir=0, hasmask=has, edgedir=0)
        K["NEWMACGRAD"](edgegrad=T.Fab(gy, (0, 0), one=True), mask=T.Fab(mg, (-1, -1), one=True), phi=T.Fab(hg, (-1, -1), one=True),
                        edgegrid=T.Box((0, 0), (NX - 1, NY)), dx=DX, dir=1, hasmask=has, edgedir=1)
        out[f"macgrad_x_mask{has}"], out[f"macgrad_y_mask{has}"] = gx, gy
    # ---- 7. DIVERGENCE (util/DivergenceF.ChF:23-57): div starts at zero, one call per direction
    ux, uy = rng.rand(NY, NX + 1), rng.rand(NY + 1, NX)
    div = np.zeros((1, NY, NX))
    for d, u in ((0, ux), (1, uy)):
        K["DIVERGENCE"](uedge=fab(u), div=fab(div), gridint=box, dx=DX[d], idir=d)
    out.update(div_ux=ux, div_uy=uy, div=div[0])

    # ---- 8. ExtrapGhostCells / CopyGhostCells on cell data, one ghost cell, non-periodic domain (util/ExtrapGhostCells.cpp:94-269): per
    # direction the one-cell strip outside the domain (adjCellLo / adjCellHi), grown by the ghost radius tangentially so that the
    # second direction also fills the corners from the first direction's strip; SIMPLEEXTRAPBC / SIMPLECOPYBC on it
    for name, kern in (("extrap", "SIMPLEEXTRAPBC"), ("copy", "SIMPLECOPYBC")):
        g = np.zeros((1, NY + 2, NX + 2))
        g[0, 1:-1, 1:-1] = head
        hi = (NX - 1, NY - 1)
        for d in (0, 1):
            t = 1 - d
            for hilo, pos in ((0, -1), (1, hi[d] + 1)):
                lo_, hi_ = [0, 0], [0, 0]
                lo_[d] = hi_[d] = pos
                lo_[t], hi_[t] = -1, hi[t] + 1
                K[kern](phi=fab(g, (-1, -1)), bcbox=T.Box(lo_, hi_), dir=d, hilo=hilo)
        out[f"ghost_{name}"] = g[0]

    # ---- 9. AMRProlongS_2 (src/AMRNonLinearPoissonOp.cpp:1141-1206): PROLONG_2_NL over a fine box that refines coarse cells
    # [2..7] x [2..6] of the periodic coarse box; `coarse` is the coarsened-fine scratch with one ghost cell, here the coarse
    # correction itself on [1..8] x [1..7] (copyTo with ghost cells; no physical boundary, no second fine box); m = 2
    p2c = 1e-2 * (rng.rand(NY, NX) - 0.5)
    flo, fhi = (4, 4), (15, 13)
    p2f = (1e-3 * rng.rand(fhi[1] - flo[1] + 1, fhi[0] - flo[0] + 1))[None]
    out.update(prolong2_coarse=p2c, prolong2_fine_in=p2f[0].copy(), prolong2_box=np.array(flo + fhi, dtype=np.int32))
    temp = p2c[1:8, 1:9].copy()
    K["PROLONG_2_NL"](phi=fab(p2f, flo), coarse=fab(temp, (1, 1)), region=T.Box(flo, fhi), m=2)
    out.update(prolong2_out=p2f[0])

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "chf_kernels.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, "with", len(out), "arrays")


if __name__ == "__main__":
    main()

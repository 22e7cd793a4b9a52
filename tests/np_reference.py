"""An independent numpy restatement of the in-tree kernel formulas (whole-level arrays, vectorised), used to cross-check
the C oracle.  Written from the .ChF / .cpp sources, not from oracle/suhmo_oracle.c; arrays are [j, i] with one ghost ring.
Citations: src/AmrHydroF.ChF:23-68,81-112,199-231; src/VCAMRNonLinearPoissonOpF.ChF:46-168,201-284,319-406,419-561,574-601;
src/AMRNonLinearPoissonOpF.ChF:607-632; src/AmrHydro.cpp:248-309; util/GradientF.ChF:55-70; src/HydroIBC.cpp:138-184."""
import numpy as np


def fill_ghosts(phi, cfg, bc_vals, dx, homogeneous):
    """exchange (periodic wrap on the whole level) + mixBCValues on a [ny+2, nx+2] array; returns a copy"""
    p = phi.copy()
    (lo_val, hi_val) = bc_vals
    if cfg.periodic[0]:
        p[:, 0], p[:, -1] = p[:, -2], p[:, 1]
    if cfg.periodic[1]:
        p[0, :], p[-1, :] = p[-2, :], p[1, :]
    for d in range(2):
        if cfg.periodic[d]:
            continue
        for side in range(2):
            typ = (cfg.bc_hi if side else cfg.bc_lo)[d]
            val = 0.0 if homogeneous else (hi_val if side else lo_val)[d]
            sign = 1.0 if side else -1.0
            if d == 0:
                near = p[1:-1, -2] if side else p[1:-1, 1]
                g = 2 * val - near if typ == 0 else near + sign * dx[0] * val
                if side:
                    p[1:-1, -1] = g
                else:
                    p[1:-1, 0] = g
            else:
                near = p[-2, 1:-1] if side else p[1, 1:-1]
                g = 2 * val - near if typ == 0 else near + sign * dx[1] * val
                if side:
                    p[-1, 1:-1] = g
                else:
                    p[0, 1:-1] = g
    return p


def nl_terms(prm, phi, B, mask, Pi, zb):
    P = Pi - 1000.0 * 9.8 * (phi - zb)
    nl = -prm["A"] * B * P * P * P
    dnl = 3.0 * prm["A"] * B * 1000.0 * 9.8 * P * P
    c = prm["cutOffbr"] > B
    nl = np.where(c, nl * (1.0 - (prm["cutOffbr"] - B) / prm["cutOffbr"]) if prm["cutOffbr"] != 0 else nl, nl)
    dnl = np.where(c, dnl * B / prm["cutOffbr"] if prm["cutOffbr"] != 0 else dnl, dnl)
    m = prm["maxOffbr"] < B
    nl = np.where(m, nl * (1.0 - (prm["maxOffbr"] - B) / prm["maxOffbr"]), nl)
    dnl = np.where(m, dnl * B / prm["maxOffbr"], dnl)
    off = mask < 0.0
    return np.where(off, 0.0, nl), np.where(off, 0.0, dnl)


def lofphi(p, bX, bY, dx, beta, nl):
    """p: ghosted phi; bX [ny, nx+1], bY [ny+1, nx]; alpha = 0"""
    c = p[1:-1, 1:-1]
    d0, d1 = 1.0 / (dx[0] * dx[0]), 1.0 / (dx[1] * dx[1])
    return 0.0 * c - beta * (bX[:, 1:] * (p[1:-1, 2:] - c) * d0 - bX[:, :-1] * (c - p[1:-1, :-2]) * d0
                             + bY[1:, :] * (p[2:, 1:-1] - c) * d1 - bY[:-1, :] * (c - p[:-2, 1:-1]) * d1) + nl


def lam(bX, bY, dx, beta):
    d0, d1 = 1.0 / (dx[0] * dx[0]), 1.0 / (dx[1] * dx[1])
    l = np.zeros_like(bX[:, 1:]) * 0.0
    l = l + d0 * beta * (bX[:, 1:] + bX[:, :-1])
    l = l + d1 * beta * (bY[1:, :] + bY[:-1, :])
    return l


def gsrb(phi, rhs, F, cfg, prm, bc_vals, beta=-1.0):
    """one levelGSRB iteration on the ghosted array phi (returns the new ghosted array, trailing homogeneous fill included)"""
    dx = cfg.dx
    jj, ii = np.meshgrid(np.arange(cfg.ny), np.arange(cfg.nx), indexing="ij")
    for color in (0, 1):
        p = fill_ghosts(phi, cfg, bc_vals, dx, False)
        nl, dnl = nl_terms(prm, p[1:-1, 1:-1], F["B"], F["mask"], F["Pi"], F["zb"])
        lo = lofphi(p, F["bX"], F["bY"], dx, beta, nl)
        denom = 1.0e-16 + lam(F["bX"], F["bY"], dx, beta) + dnl
        new = p[1:-1, 1:-1] + (rhs - lo) / denom
        sel = ((ii + jj + color) % 2) == 0
        phi = p.copy()
        phi[1:-1, 1:-1] = np.where(sel, new, p[1:-1, 1:-1])
    return fill_ghosts(phi, cfg, bc_vals, dx, True)


def residual(phi, rhs, F, cfg, prm, bc_vals, beta=-1.0):
    p = fill_ghosts(phi, cfg, bc_vals, cfg.dx, False)
    nl, _ = nl_terms(prm, p[1:-1, 1:-1], F["B"], F["mask"], F["Pi"], F["zb"])
    return rhs - lofphi(p, F["bX"], F["bY"], cfg.dx, beta, nl)


def restrict4(a):
    """RESTRICTVCNL / RESTRICTRESVCNL accumulation order: 0 + a00/4 + a10/4 + a01/4 + a11/4 (i fastest)"""
    s = 0.0 + a[0::2, 0::2] / 4.0
    s = s + a[0::2, 1::2] / 4.0
    s = s + a[1::2, 0::2] / 4.0
    return s + a[1::2, 1::2] / 4.0


def prolong_pc(fine, corr):
    return fine + np.repeat(np.repeat(corr, 2, axis=0), 2, axis=1)


def bcoef_from_head(phi, Bg, maskg, cfg, prm, bc_vals):
    """UpdateOperator/WFlx_level on a single level: phi, Bg, maskg ghosted [ny+2, nx+2] -> bX, bY"""
    dx = cfg.dx
    p = fill_ghosts(phi, cfg, bc_vals, dx, False)
    fx, fy = 1.0 / dx[0], 1.0 / dx[1]
    gxf = fx * (p[1:-1, 1:] - p[1:-1, :-1])          # x-faces of valid cells [ny, nx+1]
    gyf = fy * (p[1:, 1:-1] - p[:-1, 1:-1])          # y-faces [ny+1, nx]
    if prm["use_mask_grad"]:
        gxf = np.where((maskg[1:-1, 1:] < 1e-6) | (maskg[1:-1, :-1] < 1e-6), 0.0, gxf)
        gyf = np.where((maskg[1:, 1:-1] < 1e-6) | (maskg[:-1, 1:-1] < 1e-6), 0.0, gyf)
    g = np.zeros((2,) + p.shape)
    g[0, 1:-1, 1:-1] = 0.5 * (gxf[:, :-1] + gxf[:, 1:])
    g[1, 1:-1, 1:-1] = 0.5 * (gyf[:-1, :] + gyf[1:, :])
    for c in range(2):                                # exchange (periodic) then ExtrapGhostCells (non-periodic), x then y
        a = g[c]
        if cfg.periodic[0]:
            a[:, 0], a[:, -1] = a[:, -2], a[:, 1]
        if cfg.periodic[1]:
            a[0, :], a[-1, :] = a[-2, :], a[1, :]
        if not cfg.periodic[0]:
            a[:, 0] = 2.0 * a[:, 1] - a[:, 2]
            a[:, -1] = 2.0 * a[:, -2] - a[:, -3]
        if not cfg.periodic[1]:
            a[0, :] = 2.0 * a[1, :] - a[2, :]
            a[-1, :] = 2.0 * a[-2, :] - a[-3, :]
    sq = np.sqrt(g[0] * g[0] + g[1] * g[1])
    discr = 1.0 + 4.0 * prm["omega"] * (Bg * Bg * Bg * 9.8 * sq) / (12.0 * prm["nu"] * prm["nu"])
    Re = (-1.0 + np.sqrt(discr)) / (2.0 * prm["omega"])

    def face(a, d):
        return 0.5 * (a[1:-1, 1:] + a[1:-1, :-1]) if d == 0 else 0.5 * (a[1:, 1:-1] + a[:-1, 1:-1])

    out = []
    for d in range(2):
        Bec, Rec = face(Bg, d), face(Re, d)
        ma, mb = (maskg[1:-1, 1:], maskg[1:-1, :-1]) if d == 0 else (maskg[1:, 1:-1], maskg[:-1, 1:-1])
        im = np.where(np.abs(ma - mb) < 1e-10, np.where(ma > 0.0, 1.0, -1.0), 0.0)
        if d == 0:
            im[:, 0] = 0.0
            im[:, -1] = 0.0
        else:
            im[0, :] = 0.0
            im[-1, :] = 0.0
        num = -(Bec * Bec * Bec * 9.8)
        den = 12.0 * prm["nu"] * (1.0 + prm["omega"] * Rec)
        out.append(np.where((im < 0.0) & (prm["cutOffBcoef"] > 0), 0.0, num / den))
    return out[0], out[1]

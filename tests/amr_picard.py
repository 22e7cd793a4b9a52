"""Builds the same multi-level time-step state for the oracle (oracle/picard_amr.py) and the device (suhmo_b200/timestep_amr.py)."""
import numpy as np

from oracle import binding as ob
from oracle import picard_amr as opa
from suhmo_b200 import synthetic as syn
from suhmo_b200.timestep import picard_params


def level_arrays(cfg, l, seed=12345):
    """global [j, i] arrays (one ghost cell) of every input field on level l"""
    r = 2 ** l
    g = syn.fields(cfg, ng=1, seed=seed, level_ratio=r)
    ny, nx = g["head"].shape
    jj, ii = np.meshgrid(np.arange(ny) - 1, np.arange(nx) - 1, indexing="ij")
    x, y = (ii + 0.5) / r, (jj + 0.5) / r          # level-0 index units: the same smooth functions on every level
    out = {k: g[k] for k in ("head", "B", "Pi", "zb", "mask")}
    out["MV"] = np.full((ny, nx), 1e-6)
    out["BH"] = 0.012 + 0.002 * np.sin(0.37 * x) * np.cos(0.21 * y)
    out["BL"] = np.full((ny, nx), 2.0)
    out["mR"] = 1e-7 * (1.0 + 0.3 * np.cos(0.11 * x + 0.05 * y))
    ms = np.zeros((ny, nx))
    ms[1:-1, 1:-1] = g["rhs"]
    out["MS"] = ms
    return out


def build_oracle(cfg, level_boxes, **picard_over):
    layouts = [ob.Layout(np.asarray(b, dtype=np.int32), (0, 0, cfg.nx * 2 ** l - 1, cfg.ny * 2 ** l - 1), cfg.periodic) for l, b in enumerate(level_boxes)]
    dx = [(cfg.dx[0] / 2 ** l, cfg.dx[1] / 2 ** l) for l in range(len(level_boxes))]
    prm = ob.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr, cutOffBcoef=cfg.cutOffBcoef,
                         use_mask_grad=cfg.use_mask_grad)
    bc = ob.make_bc(cfg.bc_lo, cfg.bc_hi, cfg.bc_lo_val, cfg.bc_hi_val)
    q = picard_params(ob.PicardParams, cfg, **picard_over)
    H = opa.Hierarchy(cfg, layouts, dx, prm, bc, q, moulins=cfg.moulins)
    for l in range(H.nlev):
        for k, a in level_arrays(cfg, l).items():
            H.S[l][k].set_global(a, (-1, -1))
    return H


def build_device(ctx, cfg, level_boxes, **picard_over):
    from suhmo_b200 import amr
    from suhmo_b200.timestep_amr import AmrState
    layouts = [amr.DisjointBoxLayout(ctx, b, (0, 0, cfg.nx * 2 ** l - 1, cfg.ny * 2 ** l - 1), cfg.periodic) for l, b in enumerate(level_boxes)]
    st = AmrState(ctx, cfg, layouts, **picard_over)
    for l in range(st.nlev):
        for k, a in level_arrays(cfg, l).items():
            st.S[l][k].set_global(a, (-1, -1))
    return st

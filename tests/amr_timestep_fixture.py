"""Fixtures for tests/cpp/timestep_host.cpp: whole time steps of AmrHydro::timeStepFAS run by the CPU oracle's independent restatement
(oracle/picard_amr.py) and written, inputs and outcomes, to one binary file that the C++ host program replays on the GPU through
suhmo_b200/host/suhmo_amrhydro.hpp.  Test infrastructure: the oracle is the checker, the C++ driver is what is checked.

Layout (little endian; int = int32, real = float64; FABs in Fortran order, i fastest):
  int   magic 0x53474832, nlev, nsteps, cur_step (of the first step, as the reference counts AFTER its increment), impl_diff, nx, ny,
        periodic[2]
  real  dx0[2], dt
  int   sizeof(sg_params), bytes; int sizeof(sg_bc), bytes; int sizeof(sg_picard_params), bytes
  per level: int nbox; per box: int lo0 lo1 hi0 hi1; 10 FABs with one ghost cell: head B Pi zb mask MV BH BL mR MS
  per step:  int regrid (0/1); if 1: int tag variable (0 meltingRate, 1 Pi, 2 GapHeight), real val_min val_max fill_ratio,
             int tags_grow grow_dir[2] block_factor nesting_radius max_box_size max_level, then what the regrid must produce:
             int new_nlev; per level 1..new_nlev-1: int nbox; per box: int lo0 lo1 hi0 hi1
             int picard_iterations; int head_cycles[picard_iterations]; real x_h[picard_iterations]; int gap_cycles (-1: explicit);
             per level, per box (of the hierarchy as it is then): FABs of the valid cells: head, B
"""
import ctypes
import struct

import numpy as np

from oracle import binding as ob
from oracle import br_regrid
from oracle import picard_amr as opa
from suhmo_b200 import amr
from suhmo_b200.capi import PicardParams
from suhmo_b200.timestep import picard_params
from tests.amr_picard import build_oracle

INPUTS = ("head", "B", "Pi", "zb", "mask", "MV", "BH", "BL", "mR", "MS")
MAGIC = 0x53474832


class ImplicitTimeStep(opa.TimeStep):
    """one level, solver.use_ImplDiff: the gap height from SolveForGap_nl (src/AmrHydro.cpp:3378-3391, 3425-3455, 594-662) in place of
    the explicit update; the pieces are the oracle's own"""

    def update_gap(self, dt, cur_step=0):
        H, L = self.H, self.L
        assert H.nlev == 1
        S = H.S[0]
        self.re_and_qw(0, True)
        self.melt_rate(0)
        L.orc_rhs_gap(ctypes.byref(H.q), S["RHSb"].h, S["Pi"].h, S["Pw"].h, S["mR"].h, S["B"].h, S["Dterm"].h, S["mask"].h, S["BH"].h, S["BL"].h,
                      S["MV"].h, dt)
        cur, ones = ob.Field(H.layouts[0], 1, 1), ob.Field(H.layouts[0], 1, 0)
        ones.setval(1.0)
        cur.copy_from(S["B"])
        s = ob.LinSolver(H.layouts[0], H.dx[0][0], 1.0, dt * H.q.DiffFactor, ones, S["Dc"][0], S["Dc"][1])
        sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10 if cur_step < 50 else 5, iter_min=2, eps=1e-7, hang=1e-6,
                                   norm_thresh=1e-7)
        it, hist = s.solve(cur, S["RHSb"], sp)
        s.free()
        S["B"].copy_from(cur)
        self.exchange(S["B"])
        L.orc_copy_ghost(S["B"].h)
        return len(hist) - 1


TAG_VARS = {"meltingRate": (0, "mR"), "Pi": (1, "Pi"), "GapHeight": (2, "B")}
PERSISTENT = ("head", "B", "BH", "BL", "zb", "Pi", "MV", "mR", "Pw", "mask", "MS")   # what destructiveRegrid carries (src/AmrHydro.cpp:4335-4345)


def oracle_regrid(H, cfg, rg):
    """AmrHydro::regrid (src/AmrHydro.cpp:4227-4511) on the oracle: tagCells, Berger-Rigoutsos (oracle/br_regrid.py), destructiveRegrid
    of every persistent field and the ghost fills after it; the IBC re-initialisation is left out on both sides (no-op hook).
    Returns the new Hierarchy and its box lists."""
    L = ob.lib()
    name = TAG_VARS[rg["var"]][1]
    top = min(H.nlev - 1, rg["max_level"] - 1)
    tags = [ob.tag_cells_level(H.S[l][name], rg["val_min"], rg["val_max"], rg["tags_grow"], rg["grow_dir"]) for l in range(top + 1)]
    dom0 = (0, 0, cfg.nx - 1, cfg.ny - 1)
    base = np.array([H.layouts[0].boxes[b] for b in range(len(H.layouts[0].boxes))], dtype=np.int32)
    levels = br_regrid.regrid(dom0, base, tags, rg["fill_ratio"], rg["block_factor"], rg["nesting_radius"], rg["max_box_size"])
    layouts = [H.layouts[0]] + [ob.Layout(np.asarray(levels[l], dtype=np.int32), (0, 0, cfg.nx * 2 ** l - 1, cfg.ny * 2 ** l - 1), cfg.periodic)
                                for l in range(1, len(levels))]
    N = opa.Hierarchy(cfg, layouts, [(cfg.dx[0] / 2 ** l, cfg.dx[1] / 2 ** l) for l in range(len(layouts))], H.prm, H.bc, H.q, moulins=H.moulins)
    N.S[0] = H.S[0]
    for l in range(1, N.nlev):
        for k in PERSISTENT:
            L.orc_regrid_transfer(N.S[l][k].h, H.S[l][k].h if l < H.nlev else None, N.S[l - 1][k].h, 2)
        for k in ("BH", "BL", "MV", "mR", "Pw"):
            L.orc_extrap_ghost(N.S[l][k].h)
        for k in ("zb", "Pi"):
            L.orc_pwl_fill_patch(N.S[l][k].h, N.S[l - 1][k].h, 2)
        L.orc_exchange_full(N.S[l]["zb"].h)
        L.orc_exchange_full(N.S[l]["Pi"].h)
        L.orc_copy_ghost(N.S[l]["zb"].h)
        L.orc_extrap_ghost(N.S[l]["Pi"].h)
        L.orc_pwl_fill_patch(N.S[l]["mask"].h, N.S[l - 1]["mask"].h, 2)
        L.orc_exchange_full(N.S[l]["mask"].h)
        L.orc_copy_ghost(N.S[l]["mask"].h)
    return N, levels


def write_fixture(path, cfg, level_boxes, nsteps=2, cur_step=1, dt=1800.0, impl_diff=False, regrid_before=None):
    """run nsteps time steps on the oracle and write the fixture; returns the per-step reports.  regrid_before: {step: dict(var, val_min,
    val_max, fill_ratio, tags_grow, grow_dir, block_factor, nesting_radius, max_box_size, max_level)} -- regrid before those steps"""
    regrid_before = regrid_before or {}
    level_boxes = [np.asarray(b, dtype=np.int32) for b in level_boxes]
    over = dict(use_ImplDiff=1) if impl_diff else {}
    H = build_oracle(cfg, level_boxes, **over)
    ts = ImplicitTimeStep(H) if impl_diff else opa.TimeStep(H)
    prm = amr.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr, cutOffBcoef=cfg.cutOffBcoef,
                          use_mask_grad=cfg.use_mask_grad)
    bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi, cfg.bc_lo_val, cfg.bc_hi_val)
    q = picard_params(PicardParams, cfg, **over)
    out = bytearray()
    out += struct.pack("<9i", MAGIC, H.nlev, nsteps, cur_step, int(impl_diff), cfg.nx, cfg.ny, int(cfg.periodic[0]), int(cfg.periodic[1]))
    out += struct.pack("<3d", cfg.dx[0], cfg.dx[1], dt)
    for s in (prm, bc, q):
        b = bytes(s)
        out += struct.pack("<i", len(b)) + b
    for l in range(H.nlev):
        boxes = np.asarray(level_boxes[l], dtype=np.int32)
        out += struct.pack("<i", len(boxes))
        for b, bx in enumerate(boxes):
            out += struct.pack("<4i", *[int(v) for v in bx])
            for k in INPUTS:
                out += np.ascontiguousarray(H.S[l][k].fab(b)[0][0], dtype=np.float64).tobytes()
    reports = []
    for step in range(nsteps):
        rg = regrid_before.get(step)
        out += struct.pack("<i", int(rg is not None))
        if rg is not None:
            H, level_boxes = oracle_regrid(H, cfg, rg)
            ts = opa.TimeStep(H)
            out += struct.pack("<i3d7i", TAG_VARS[rg["var"]][0], rg["val_min"], rg["val_max"], rg["fill_ratio"], rg["tags_grow"], rg["grow_dir"][0],
                               rg["grow_dir"][1], rg["block_factor"], rg["nesting_radius"], rg["max_box_size"], rg["max_level"])
            out += struct.pack("<i", len(level_boxes))
            for l in range(1, len(level_boxes)):
                out += struct.pack("<i", len(level_boxes[l])) + np.ascontiguousarray(level_boxes[l], dtype=np.int32).tobytes()
        if impl_diff:
            # the same loop as opa.TimeStep.time_step with the implicit gap update at its end
            r = _implicit_step(ts, dt, cur_step + step)
        else:
            r = ts.time_step(dt, cur_step + step)
            r["gap_cycles"] = -1
        r["boxes"] = [len(b) for b in level_boxes]
        reports.append(r)
        n = r["picard_iterations"]
        out += struct.pack("<i", n) + struct.pack(f"<{n}i", *r["head_cycles"]) + struct.pack(f"<{n}d", *r["x_h"]) + struct.pack("<i", r["gap_cycles"])
        for l in range(H.nlev):
            for b in range(len(level_boxes[l])):
                for k in ("head", "B"):
                    out += np.ascontiguousarray(H.S[l][k].fab(b)[0][0][1:-1, 1:-1], dtype=np.float64).tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))
    return reports


def _implicit_step(ts, dt, cur_step):
    saved = ts.update_gap
    cycles = []
    ts.update_gap = lambda dt_: cycles.append(saved(dt_, cur_step))
    try:
        r = opa.TimeStep.time_step(ts, dt, cur_step)
    finally:
        ts.update_gap = saved
    r["gap_cycles"] = cycles[0]
    return r

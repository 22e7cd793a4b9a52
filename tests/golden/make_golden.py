#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle (NOT from the reference, which cannot be built here and ships no
vectors for this path -- SURVEY.md 8c): frozen inputs and outputs that guard the oracle and the CUDA path against drift.
  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import binding as ob  # noqa: E402
from suhmo_b200 import synthetic as syn  # noqa: E402
from tests.problem import AmrOracleSide, OracleSide, amr_hierarchy  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def single_level():
    cfg = syn.config("C1", 1)  # exec/0_convergence_channelized/1lev: 32 x 8 cells, two 16 x 8 boxes
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes)
    out = {"in_" + k: orc.F[k].get_global() for k in ("head", "B", "Pi", "zb", "mask", "rhs")}
    orc.init_bcoef()
    out["bX0"], out["bY0"] = orc.F["bX"].get_global(), orc.F["bY"].get_global()
    it, hist = orc.solver().solve(orc.F["head"], orc.F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=5))
    out["resnorm"], out["head5"] = hist, orc.F["head"].get_global()
    np.savez_compressed(os.path.join(HERE, "c1_1lev_vcycles.npz"), **out)


def three_levels():
    cfg, lv = amr_hierarchy()
    orc = AmrOracleSide(cfg, lv)
    orc.average_down("head")
    orc.init_bcoef()
    it, hist = orc.solver().solve(orc.fields("head"), orc.fields("rhs"), 2, ob.make_solver_params(bottom=10, fixed_cycles=3))
    out = {"resnorm": hist}
    for l in range(3):
        out[f"head3_L{l}"] = np.nan_to_num(orc.F[l]["head"].get_global(), nan=0.0)
    np.savez_compressed(os.path.join(HERE, "amr_3lev_vcycles.npz"), **out)


def gap_solve():
    """implicit gap-height solve (SolveForGap_nl) on the C2 level-0 grid with the reference's solver constants"""
    from tests import gapsolve as gs
    cfg = syn.config("C2", 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    o = gs.OracleGap(cfg, boxes)
    sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10, iter_min=2, eps=1e-7, hang=1e-6, norm_thresh=1e-7)
    it, hist = o.solver.solve(o.F["b"], o.F["rhs"], sp)
    np.savez_compressed(os.path.join(HERE, "gap_c2_solve.npz"), resnorm=hist, gap=o.F["b"].get_global(), bottom_iters=o.solver.bottom_iters)


def host_smoke_problem():
    """the two-level problem of tests/cpp/host_smoke.cpp: 32^2 base grid in four 16^2 boxes, one 32^2-cell refined patch across them"""
    cfg = syn.config("C5", 1)
    cfg.nx, cfg.ny, cfg.max_box_size = 32, 32, 16
    base = syn.domain_split(32, 32, 16, 2)
    lev1 = np.array([(24, 24, 55, 55)], dtype=np.int32)
    return cfg, [base, lev1]


def host_smoke_fixture(cycles=3):
    """flat little-endian binary for the C++ host program (no npz reader there): int32 magic, nlev, cycles; float64 dx0[2];
    per level: int32 nbox; per box: int32 lo0 lo1 hi0 hi1, then the FArrayBoxes head, B, Pi, zb, mask (1 ghost cell), rhs (none)
    as the oracle holds them BEFORE the solve and the oracle's head (no ghost cells) AFTER `cycles` FAS V-cycles; finally the
    residual-norm history (cycles + 1 doubles)."""
    cfg, lv = host_smoke_problem()
    orc = AmrOracleSide(cfg, lv)
    orc.average_down("head")
    pre = [{k: [orc.F[l][k].fab(b)[0].copy() for b in range(len(lv[l]))] for k in ("head", "B", "Pi", "zb", "mask", "rhs")} for l in range(2)]
    orc.init_bcoef()
    it, hist = orc.solver().solve(orc.fields("head"), orc.fields("rhs"), 1, ob.make_solver_params(bottom=10, fixed_cycles=cycles))
    with open(os.path.join(HERE, "host_smoke_2lev.bin"), "wb") as f:
        np.array([0x53474831, 2, cycles], dtype="<i4").tofile(f)
        np.array(cfg.dx, dtype="<f8").tofile(f)
        for l in range(2):
            np.array([len(lv[l])], dtype="<i4").tofile(f)
            for b, bx in enumerate(lv[l]):
                np.asarray(bx, dtype="<i4").tofile(f)
                for k in ("head", "B", "Pi", "zb", "mask", "rhs"):
                    pre[l][k][b].astype("<f8").tofile(f)
                orc.F[l]["head"].fab(b)[0][:, 1:-1, 1:-1].astype("<f8").tofile(f)
        hist.astype("<f8").tofile(f)


if __name__ == "__main__":
    single_level()
    three_levels()
    gap_solve()
    host_smoke_fixture()
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))

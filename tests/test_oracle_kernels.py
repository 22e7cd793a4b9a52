"""The C oracle against an INDEPENDENT numpy restatement of the in-tree kernel formulas (tests/np_reference.py) and against
the frozen fixture under tests/golden/.  The reference itself holds no golden vectors for this path: these
tests pin the oracle's arithmetic to a second reading of the sources, bit for bit, and guard it against drift."""
import os

import numpy as np
import pytest

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests import np_reference as npr
from tests.problem import OracleSide

CASES = [("C1", 1), ("C1", 2), ("C2", 1), ("C3", 1), ("C4", 1), ("C5", 1)]


def ghosted(field, cfg):
    """whole-level [ny+2, nx+2] array from the oracle's boxes (valid data; ring left zero)"""
    g = np.zeros((cfg.ny + 2, cfg.nx + 2))
    g[1:-1, 1:-1] = field.get_global()
    return g


def setup(name, scale, bc_vals=None):
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes, bc_vals=bc_vals)
    orc.init_bcoef()
    F = {k: orc.F[k].get_global() for k in ("B", "Pi", "zb", "mask", "bX", "bY", "rhs")}
    prm = dict(orc.prm_kw, use_NL=1)
    return cfg, orc, F, prm


@pytest.mark.parametrize("name,scale", CASES)
def test_update_operator_matches_numpy(name, scale):
    cfg, orc, F, prm = setup(name, scale)
    # whole-level ghosted B and mask as AmrHydro hands them over: exchange + CopyGhostCells
    Bg, Mg = ghosted(orc.F["B"], cfg), ghosted(orc.F["mask"], cfg)
    for a in (Bg, Mg):
        if cfg.periodic[0]:
            a[:, 0], a[:, -1] = a[:, -2], a[:, 1]
        else:
            a[:, 0], a[:, -1] = a[:, 1], a[:, -2]
        if cfg.periodic[1]:
            a[0, :], a[-1, :] = a[-2, :], a[1, :]
        else:
            a[0, :], a[-1, :] = a[1, :], a[-2, :]
    bX, bY = npr.bcoef_from_head(ghosted(orc.F["head"], cfg), Bg, Mg, cfg, prm, orc.bc_vals)
    assert np.array_equal(bX, F["bX"]) and np.array_equal(bY, F["bY"])


@pytest.mark.parametrize("name,scale", CASES)
def test_relax_residual_restrict_prolong_match_numpy(name, scale):
    cfg, orc, F, prm = setup(name, scale, bc_vals=((120.0, 0.0), (1e-3, 0.0)))
    op = orc.op()
    phi = ghosted(orc.F["head"], cfg)
    # residual
    ores = ob.Field(orc.layout, 1, 0)
    op.residual(ores, orc.F["head"], orc.F["rhs"])
    assert np.array_equal(ores.get_global(), npr.residual(phi, F["rhs"], F, cfg, prm, orc.bc_vals))
    # lambda
    assert np.array_equal(op.lambda_field().get_global(), npr.lam(F["bX"], F["bY"], cfg.dx, -1.0))
    # two GSRB iterations
    for _ in range(2):
        op.relax(orc.F["head"], orc.F["rhs"], 1)
        phi = npr.gsrb(phi, F["rhs"], F, cfg, prm, orc.bc_vals)
        assert np.array_equal(orc.F["head"].get_global(), phi[1:-1, 1:-1])
    # restriction of the solution and of the residual, prolongation
    olc = orc.layout.coarsen(2)
    oresc, ophic = ob.Field(olc, 1, 0), ob.Field(olc, 1, 1)
    op.restrict_r(ophic, orc.F["head"])
    assert np.array_equal(ophic.get_global(), npr.restrict4(phi[1:-1, 1:-1]))
    op.restrict_residual(oresc, orc.F["head"], orc.F["rhs"])
    assert np.array_equal(oresc.get_global(), npr.restrict4(npr.residual(phi, F["rhs"], F, cfg, prm, orc.bc_vals)))
    before = orc.F["head"].get_global().copy()
    op.prolong_increment(orc.F["head"], ophic)
    assert np.array_equal(orc.F["head"].get_global(), npr.prolong_pc(before, ophic.get_global()))


def test_nonlinear_terms_cutoffs_and_mask():
    """COMPUTENONLINEARTERMS incl. the b-cutoff scalings and the mask<0 branch (src/AmrHydroF.ChF:40-62)"""
    cfg = syn.config("C4", 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes, prm_over=dict(cutOffbr=0.012, maxOffbr=0.0135))
    prm = dict(orc.prm_kw, use_NL=1)
    nl, dnl = ob.Field(orc.layout, 1, 0), ob.Field(orc.layout, 1, 0)
    F = orc.F
    ob.lib().orc_compute_nl(orc.prm, F["head"].h, F["B"].h, F["mask"].h, F["Pi"].h, F["zb"].h, nl.h, dnl.h)
    g = {k: F[k].get_global() for k in ("head", "B", "mask", "Pi", "zb")}
    enl, ednl = npr.nl_terms(prm, g["head"], g["B"], g["mask"], g["Pi"], g["zb"])
    assert (g["B"] < 0.012).any() and (g["B"] > 0.0135).any() and (g["mask"] < 0).any()
    assert np.array_equal(nl.get_global(), enl) and np.array_equal(dnl.get_global(), ednl)


def test_golden_fixture_quickstart_1lev():
    """frozen inputs/outputs made by tests/golden/make_golden.py (QuickStart 1lev, 32x8, 5 FAS V-cycles)"""
    path = os.path.join(os.path.dirname(__file__), "golden", "c1_1lev_vcycles.npz")
    z = np.load(path)
    cfg = syn.config("C1", 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes)
    for k in ("head", "B", "Pi", "zb", "mask", "rhs"):
        assert np.array_equal(orc.F[k].get_global(), z["in_" + k]), k
    orc.init_bcoef()
    assert np.array_equal(orc.F["bX"].get_global(), z["bX0"]) and np.array_equal(orc.F["bY"].get_global(), z["bY0"])
    it, hist = orc.solver().solve(orc.F["head"], orc.F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=5))
    assert np.array_equal(hist, z["resnorm"])
    assert np.array_equal(orc.F["head"].get_global(), z["head5"])

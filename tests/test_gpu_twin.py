"""k_gsrb_twin (two GSRB iterations per launch, branch-free point update with nvcc's division fast path inlined and the slow
path deferred) against the oracle: the default mode's use of it (levels above 2 M cells), every multigrid depth through it
(tune key 19), and the paths that leave the fast update -- the cut-off branches of COMPUTENONLINEARTERMS
(src/AmrHydroF.ChF:52-64), a zero dividend (rhs = L(phi) exactly), boundary warps.  Everything bit-exact."""
import dataclasses

import numpy as np
import pytest

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests.problem import GpuSide, OracleSide, fields_equal

pytestmark = pytest.mark.gpu


def same(g, o, what):
    d, eq = fields_equal(g, o)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


def build(ctx, cfg, box=64, **kw):
    boxes = syn.domain_split(cfg.nx, cfg.ny, box, cfg.block_factor)
    orc = OracleSide(cfg, boxes, **kw)
    orc.init_bcoef()
    return orc, GpuSide(ctx, orc)


@pytest.fixture
def twin_everywhere(gpu_ctx):
    gpu_ctx.set_relax_mode(1)
    gpu_ctx.set_tuning(19, 1)
    yield gpu_ctx
    gpu_ctx.set_tuning(19, 0)


def test_default_mode_above_2m_cells(gpu_ctx):
    """2304 x 1024 cells: the default relax mode takes k_gsrb_twin on depth 0 (interior warps on the lean update, boundary
    warps on the exact one), k_gsrb_tile below; relax counts 1..5 cover pairs and the odd remainder; then V-cycles"""
    cfg = dataclasses.replace(syn.config("C5", 1), nx=2304, ny=1024, domain_size=(9000.0, 4000.0))
    orc, gpu = build(gpu_ctx, cfg)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    for n in (2, 1, 4, 5):
        oop.relax(orc.F["head"], orc.F["rhs"], n)
        gop.relax(gpu.F["head"], gpu.F["rhs"], n)
        same(gpu.F["head"], orc.F["head"], f"relax x{n}")
    # the default really is the two-iterations-per-launch kernel here: tune key 18 = 1 (one iteration per launch) needs two more launches for four iterations
    counts = []
    for key18 in (0, 1):
        gpu_ctx.set_tuning(18, key18)
        n0 = gpu_ctx.kernel_launches()
        oop.relax(orc.F["head"], orc.F["rhs"], 4)
        gop.relax(gpu.F["head"], gpu.F["rhs"], 4)
        counts.append(gpu_ctx.kernel_launches() - n0)
        same(gpu.F["head"], orc.F["head"], f"relax x4, tune key 18 = {key18}")
    gpu_ctx.set_tuning(18, 0)
    assert counts[1] - counts[0] == 2, counts
    it, ohist = orc.solver().solve(orc.F["head"], orc.F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=2))
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, _ = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=2)
    assert np.array_equal(ghist, ohist), (ghist, ohist)
    same(gpu.F["head"], orc.F["head"], "head after 2 V-cycles")


@pytest.mark.parametrize("name,scale", [("C1", 4), ("C2", 2), ("C4", 1), ("C5", 2)])
def test_vcycles_every_depth(twin_everywhere, name, scale):
    """fixed V-cycles with k_gsrb_twin on every multigrid depth (C4: the valley's ice mask < 0 streams through the masked variant)"""
    ctx = twin_everywhere
    cfg = syn.config(name, scale)
    orc, gpu = build(ctx, cfg, box=cfg.max_box_size)
    it, ohist = orc.solver().solve(orc.F["head"], orc.F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=4))
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    git, ghist, _ = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=4)
    assert np.array_equal(ghist, ohist), (ghist, ohist)
    same(gpu.F["head"], orc.F["head"], f"{name}: head after 4 V-cycles")


@pytest.mark.parametrize("periodic", [(1, 0), (1, 1)])
def test_periodic_sides(twin_everywhere, periodic):
    """no physical boundary in x (and y): every warp runs the lean update, the ring cells are periodic images"""
    ctx = twin_everywhere
    kw = dict(nx=256, ny=192, periodic=periodic, max_box_size=64, domain_size=(8000.0, 6000.0))
    if periodic == (1, 0):
        kw.update(bc_lo=(0, 0), bc_hi=(0, 1))
    cfg = dataclasses.replace(syn.config("C5", 1), **kw)
    orc, gpu = build(ctx, cfg)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    for n in (2, 4, 3):
        oop.relax(orc.F["head"], orc.F["rhs"], n)
        gop.relax(gpu.F["head"], gpu.F["rhs"], n)
        same(gpu.F["head"], orc.F["head"], f"periodic {periodic}: relax x{n}")


@pytest.mark.parametrize("which", ["cutOffbr", "maxOffbr", "both"])
def test_cutoff_branches_leave_the_fast_update(twin_everywhere, which):
    """cutOffbr above / maxOffbr below part of the gap heights: those cells take COMPUTENONLINEARTERMS' cut-off branches, the lane
    raises the flag and the step is redone with the exact update; cells elsewhere stay on the lean update"""
    ctx = twin_everywhere
    base = dataclasses.replace(syn.config("C5", 1), nx=512, ny=384, domain_size=(8000.0, 6000.0))
    B = syn.fields(base, ng=1)["B"]
    lo, hi = float(np.quantile(B, 0.3)), float(np.quantile(B, 0.7))
    assert lo < hi
    cut, mx = {"cutOffbr": (lo, 10000.0), "maxOffbr": (0.0, hi), "both": (lo, hi)}[which]
    cfg = dataclasses.replace(base, cutOffbr=cut, maxOffbr=mx)
    orc, gpu = build(ctx, cfg)
    Bv = orc.F["B"].get_global()
    Bv = Bv[~np.isnan(Bv)]
    assert (cut > Bv).any() or (mx < Bv).any(), "the case must reach a cut-off branch"
    assert ((cut <= Bv) & (mx >= Bv)).any(), "and leave cells outside it"
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    for n in (2, 4):
        oop.relax(orc.F["head"], orc.F["rhs"], n)
        gop.relax(gpu.F["head"], gpu.F["rhs"], n)
        same(gpu.F["head"], orc.F["head"], f"cut-offs ({cut}, {mx}): relax x{n}")


def test_zero_dividend_takes_the_slow_path(twin_everywhere):
    """rhs = L(phi) bit for bit: the dividend of every point update is exactly zero, which the fast division does not handle
    (sign of zero) -- every step falls back, and the head must come back unchanged bit for bit as it does from the oracle"""
    ctx = twin_everywhere
    cfg = dataclasses.replace(syn.config("C5", 1), nx=512, ny=256, domain_size=(8000.0, 4000.0))
    orc, gpu = build(ctx, cfg)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    olof = ob.Field(orc.layout, 1, 0)
    oop.apply(olof, orc.F["head"], False)
    glof = gpu.new_like("rhs")
    gop.applyOp(glof, gpu.F["head"], False)
    same(glof, olof, "L(phi)")
    before = gpu.F["head"].get_global().copy()
    oop.relax(orc.F["head"], olof, 2)
    gop.relax(gpu.F["head"], glof, 2)
    same(gpu.F["head"], orc.F["head"], "relax with rhs = L(phi)")
    # red cells see a zero dividend exactly; black cells see the operator re-evaluated around unchanged red cells: zero as well
    after = gpu.F["head"].get_global()
    ok = ~np.isnan(before)
    assert np.mean(after[ok] == before[ok]) > 0.99, "the dividend should have been exactly zero nearly everywhere"

"""GPU parity for the implicit gap-height solve (AmrHydro::SolveForGap_nl, src/AmrHydro.cpp:594-662): every
VCAMRPoissonOp2 entry point, the RelaxSolver bottom solve, one correction-form V-cycle and the whole solve with the reference's
parameters, bit for bit against the oracle through the C ABI."""
import numpy as np
import pytest

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests import gapsolve as gs
from tests.problem import fields_equal

pytestmark = pytest.mark.gpu

CASES = [("C1", 2, None), ("C1", 8, None), ("C2", 1, None), ("C2", 4, 32), ("C4", 1, None), ("C5", 1, None)]


def make(ctx, name, scale, mb=None, **kw):
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, mb or cfg.max_box_size, cfg.block_factor)
    orc = gs.OracleGap(cfg, boxes, **kw)
    return cfg, orc, gs.GpuGap(ctx, orc)


def same(gpu_ld, orc_f, what):
    d, eq = fields_equal(gpu_ld, orc_f)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


@pytest.mark.parametrize("name,scale,mb", CASES)
def test_operator_entry_points(gpu_ctx, name, scale, mb):
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    S, G = orc.solver, gpu.solver
    assert G.depth == S.depth
    rng = np.random.RandomState(11)
    for d in range(S.depth):
        L = S.layout_at(d)
        dom = L.domain
        ny, nx = dom[3] + 1, dom[2] + 1
        ophi, orhs, ores = ob.Field(L, 1, 1), ob.Field(L, 1, 0), ob.Field(L, 1, 0)
        ophi.set_global(rng.rand(ny + 2, nx + 2), (-1, -1))
        orhs.set_global(rng.rand(ny, nx), (0, 0))
        gphi, grhs, gres = gpu.new(d, 1, ophi), gpu.new(d, 0, orhs), gpu.new(d, 0)
        glam = gpu.new(d, 0)
        G.lambda_(glam, depth=d)
        same(glam, S.lambda_field(d), f"lambda depth {d}")
        S.relax(ophi, orhs, 3, depth=d)
        G.relax(gphi, grhs, 3, depth=d)
        same(gphi, ophi, f"relax depth {d}")
        S.residual(ores, ophi, orhs, depth=d)
        G.residual(gres, gphi, grhs, depth=d)
        same(gres, ores, f"residual depth {d}")
        S.applyOp(ores, ophi, depth=d)
        G.applyOp(gres, gphi, depth=d)
        same(gres, ores, f"applyOp depth {d}")
        S.preCond(ophi, orhs, depth=d)
        G.preCond(gphi, grhs, depth=d)
        same(gphi, ophi, f"preCond depth {d}")
        if d + 1 < S.depth:
            Lc = S.layout_at(d + 1)
            orc_c, gc = ob.Field(Lc, 1, 0), gpu.new(d + 1, 0)
            S.restrictResidual(orc_c, ophi, orhs, depth=d)
            G.restrictResidual(gc, gphi, grhs, depth=d)
            same(gc, orc_c, f"restrictResidual depth {d}")
            S.prolongIncrement(ophi, orc_c, depth=d)
            G.prolongIncrement(gphi, gc, depth=d)
            same(gphi, ophi, f"prolongIncrement depth {d}")


@pytest.mark.parametrize("name,scale,mb", CASES)
def test_bottom_solver_and_vcycle(gpu_ctx, name, scale, mb):
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    S, G = orc.solver, gpu.solver
    d = S.depth - 1
    Lb = S.layout_at(d)
    rng = np.random.RandomState(3)
    orhs, oe = ob.Field(Lb, 1, 0), ob.Field(Lb, 1, 1)
    orhs.set_global(rng.rand(Lb.domain[3] + 1, Lb.domain[2] + 1), (0, 0))
    oe.setval(0.0)
    grhs, ge = gpu.new(d, 0, orhs), gpu.new(d, 1, oe)
    oi = S.bottom_solve(oe, orhs)
    gi = G.bottom_solve(ge, grhs)
    assert gi == oi
    same(ge, oe, "RelaxSolver::solve")
    # one V-cycle on the initial residual
    sp = ob.make_solver_params(pre=2, post=2, bottom=4)
    ores, ocorr = ob.Field(orc.layout, 1, 0), ob.Field(orc.layout, 1, 1)
    ocorr.setval(0.0)
    S.residual(ores, orc.F["b"], orc.F["rhs"])
    gres, gcorr = gpu.new(0, 0), gpu.new(0, 1)
    G.residual(gres, gpu.F["b"], gpu.F["rhs"])
    same(gres, ores, "initial residual")
    S.vcycle(ocorr, ores, sp)
    G.vcycle(gcorr, gres)
    same(gcorr, ocorr, "V-cycle correction")


@pytest.mark.parametrize("name,scale,mb", CASES)
@pytest.mark.parametrize("cur_step", [0, 100])
def test_solve_for_gap(gpu_ctx, name, scale, mb, cur_step):
    """SolveForGap_nl end to end with the reference's constants: same iteration count, same norms, same gap height"""
    cfg, orc, gpu = make(gpu_ctx, name, scale, mb)
    sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10 if cur_step < 50 else 5, iter_min=2, eps=1e-7, hang=1e-6,
                               norm_thresh=1e-7)
    oit, ohist = orc.solver.solve(orc.F["b"], orc.F["rhs"], sp)
    dt_df = orc.beta
    git, ghist, stats = gpu.amr.SolveForGap_nl(gpu_ctx, [gpu.layout], [gpu.F["a"]], [gpu.F["bX"]], [gpu.F["bY"]], [], (orc.dx, orc.dx),
                                               [gpu.F["b"]], [gpu.F["rhs"]], dt_df, 1.0, cur_step)
    assert git == oit and np.array_equal(ghist, ohist), (git, oit, ghist, ohist)
    same(gpu.F["b"], orc.F["b"], "gap height after the implicit solve")
    assert stats.kernel_launches > 0 and oit >= 2


def test_unsupported_shapes(gpu_ctx):
    from suhmo_b200.capi import ERR_INVALID, ERR_UNSUPPORTED, SuhmoGpuError
    cfg, orc, gpu = make(gpu_ctx, "C2", 1)
    amr = gpu.amr
    with pytest.raises(SuhmoGpuError) as e:
        amr.GapHeightSolver().define(gpu_ctx, [gpu.layout, gpu.layout], [2], (orc.dx, orc.dx), 1.0, [gpu.F["a"]] * 2, 1.0, [gpu.F["bX"]] * 2,
                                     [gpu.F["bY"]] * 2)
    assert e.value.code == ERR_UNSUPPORTED
    with pytest.raises(SuhmoGpuError) as e:
        gpu.solver.relax(gpu.new(1, 1), gpu.F["rhs"], 1)   # coarse field handed to depth 0
    assert e.value.code == ERR_INVALID


def test_repeated_calls_reuse_the_solver(gpu_ctx):
    """sg_solve_for_gap keeps factory + solver per (grids, coefficient fields) and refreshes them: a second call with another dt
    and changed diffusion coefficients must equal a fresh oracle solve"""
    cfg, orc, gpu = make(gpu_ctx, "C2", 2)
    sp = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=5, iter_min=2, eps=1e-7, hang=1e-6, norm_thresh=1e-7)
    for rep, (scale, dtf) in enumerate([(1.0, 1.0), (0.5, 3.0), (2.0, 0.25)]):
        # change D on both sides, keep the same device fields (same handles -> cache hit)
        for k in ("bX", "bY"):
            a = orc.g[k] * scale
            orc.F[k].set_global(a, (0, 0))
            gpu.push(gpu.F[k], orc.F[k])
        beta = orc.beta * dtf
        osol = ob.LinSolver(orc.layout, orc.dx, 1.0, beta, orc.F["a"], orc.F["bX"], orc.F["bY"])
        oit, ohist = osol.solve(orc.F["b"], orc.F["rhs"], sp)
        osol.free()
        git, ghist, st = gpu.amr.SolveForGap_nl(gpu_ctx, [gpu.layout], [gpu.F["a"]], [gpu.F["bX"]], [gpu.F["bY"]], [], (orc.dx, orc.dx),
                                                [gpu.F["b"]], [gpu.F["rhs"]], beta, 1.0, 100)
        assert git == oit and np.array_equal(ghist, ohist), (rep, git, oit)
        same(gpu.F["b"], orc.F["b"], f"gap height, call {rep}")


def test_against_golden_fixture(gpu_ctx):
    """the CUDA path against the committed fixture tests/golden/gap_c2_solve.npz"""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gap_c2_solve.npz"))
    cfg, orc, gpu = make(gpu_ctx, "C2", 1)
    git, ghist, st = gpu.amr.SolveForGap_nl(gpu_ctx, [gpu.layout], [gpu.F["a"]], [gpu.F["bX"]], [gpu.F["bY"]], [], (orc.dx, orc.dx),
                                            [gpu.F["b"]], [gpu.F["rhs"]], orc.beta, 1.0, 0)
    assert np.array_equal(ghist, z["resnorm"]) and np.array_equal(gpu.F["b"].get_global(), z["gap"])

"""CPU: the oracle's restatement of the implicit gap-height solve (stock VCAMRPoissonOp2 + linear AMRMultiGrid + RelaxSolver,
src/AmrHydro.cpp:594-662) against an independent numpy restatement of the kernels and a sparse direct solve of the system."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spl

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests import gapsolve as gs


def make(name, scale, **kw):
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    return cfg, gs.OracleGap(cfg, boxes, **kw)


@pytest.mark.parametrize("name,scale", [("C1", 2), ("C2", 1), ("C4", 1)])
def test_kernels_match_numpy(name, scale):
    cfg, o = make(name, scale)
    g, S = o.g, o.solver
    assert np.array_equal(S.lambda_field().get_global(), gs.np_lambda(g, o.dx, o.alpha, o.beta))
    p = g["b"].copy()
    for it in range(2):
        S.relax(o.F["b"], o.F["rhs"], 1)
        p = gs.np_gsrb(p, g["rhs"], g, cfg, o.dx, o.alpha, o.beta)
        assert np.array_equal(o.F["b"].get_global(), p[1:-1, 1:-1]), it
    res = ob.Field(o.layout, 1, 0)
    S.residual(res, o.F["b"], o.F["rhs"])
    pr = gs.np_ghosts(p, cfg)
    want = g["rhs"] - gs.np_lofphi(pr, g, o.dx, o.alpha, o.beta)
    assert np.array_equal(res.get_global(), want)
    lhs = ob.Field(o.layout, 1, 0)
    S.applyOp(lhs, o.F["b"])
    assert np.array_equal(lhs.get_global(), gs.np_lofphi(pr, g, o.dx, o.alpha, o.beta))
    if S.depth > 1:
        Lc = o.layout.coarsen(2)
        rc = ob.Field(Lc, 1, 0)
        S.restrictResidual(rc, o.F["b"], o.F["rhs"])
        s = 0.0 + want[0::2, 0::2] / 4.0
        s = s + want[0::2, 1::2] / 4.0
        s = s + want[1::2, 0::2] / 4.0
        s = s + want[1::2, 1::2] / 4.0
        assert np.array_equal(rc.get_global(), s)
        before = o.F["b"].get_global().copy()
        S.prolongIncrement(o.F["b"], rc)
        assert np.array_equal(o.F["b"].get_global(), before + np.repeat(np.repeat(s, 2, axis=0), 2, axis=1))


def assemble(g, cfg, dx, alpha, beta):
    ny, nx = g["a"].shape
    idx = np.arange(nx * ny).reshape(ny, nx)
    A = sp.lil_matrix((nx * ny, nx * ny))
    s = beta / (dx * dx)
    for j in range(ny):
        for i in range(nx):
            k = idx[j, i]
            A[k, k] += alpha * g["a"][j, i]
            for (dj, di, b) in ((0, 1, g["bX"][j, i + 1]), (0, -1, g["bX"][j, i]), (1, 0, g["bY"][j + 1, i]), (-1, 0, g["bY"][j, i])):
                jj, ii = j + dj, i + di
                if not (0 <= ii < nx):
                    if not cfg.periodic[0]:
                        continue          # ghost = near: the face carries no flux
                    ii %= nx
                if not (0 <= jj < ny):
                    if not cfg.periodic[1]:
                        continue
                    jj %= ny
                A[k, k] += s * b
                A[k, idx[jj, ii]] -= s * b
    return A.tocsr()


@pytest.mark.parametrize("name,scale", [("C1", 2), ("C2", 1), ("C4", 1)])
def test_solve_converges_to_direct_solution(name, scale):
    cfg, o = make(name, scale)
    sp_ = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10, iter_min=2, eps=1e-7, hang=1e-6, norm_thresh=1e-7)
    sp_.eps, sp_.norm_thresh = 1e-12, 1e-14
    it, hist = o.solver.solve(o.F["b"], o.F["rhs"], sp_)
    assert 2 <= it <= 100
    assert hist[-1] < 1e-9 * hist[0]
    x = spl.spsolve(assemble(o.g, cfg, o.dx, o.alpha, o.beta).tocsc(), o.g["rhs"].ravel()).reshape(cfg.ny, cfg.nx)
    got = o.F["b"].get_global()
    assert np.max(np.abs(got - x)) < 1e-9 * np.max(np.abs(x))
    assert 0 <= o.solver.bottom_iters <= 40


def test_stop_logic_and_bottom_solver():
    cfg, o = make("C2", 2)
    # reference parameters: setSolverParameters(2,2,4,1,100,1e-7,1e-6,1e-7), m_imin = 10 (step < 50), m_iterMin = 2
    sp_ = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10, iter_min=2, eps=1e-7, hang=1e-6, norm_thresh=1e-7)
    it, hist = o.solver.solve(o.F["b"], o.F["rhs"], sp_)
    assert it >= 2 and len(hist) == it + 1
    assert hist[-1] <= max(1e-7 * hist[0], 1e-7) or it == 100 or hist[-1] >= (1 - 1e-6) * hist[-2]
    assert np.all(np.diff(hist[:3]) < 0)
    # fixed-cycle protocol and the V-cycle entry point agree
    cfg, a = make("C2", 2)
    cfg, b = make("C2", 2)
    fx = ob.make_solver_params(pre=2, post=2, bottom=4, fixed_cycles=1)
    a.solver.solve(a.F["b"], a.F["rhs"], fx)
    res = ob.Field(b.layout, 1, 0)
    corr = ob.Field(b.layout, 1, 1)
    corr.setval(0.0)
    b.solver.residual(res, b.F["b"], b.F["rhs"])
    b.solver.vcycle(corr, res, fx)
    assert np.array_equal(a.F["b"].get_global(), b.F["b"].get_global() + corr.get_global())
    # bottom solver alone: drives the coarsest-level residual down by 1e-6 in the 2-norm or stops at 40
    d = b.solver.depth - 1
    Lb = b.solver.layout_at(d)
    rng = np.random.RandomState(3)
    rhs, e, r = ob.Field(Lb, 1, 0), ob.Field(Lb, 1, 1), ob.Field(Lb, 1, 0)
    dom = Lb.domain
    rhs.set_global(rng.rand(dom[3] + 1, dom[2] + 1), (0, 0))
    e.setval(0.0)
    b.solver.residual(r, e, rhs, depth=d)
    n0 = np.sqrt(np.sum(r.get_global() ** 2))
    its = b.solver.bottom_solve(e, rhs)
    b.solver.residual(r, e, rhs, depth=d)
    n1 = np.sqrt(np.sum(r.get_global() ** 2))
    assert 1 <= its <= 40 and (n1 < 1e-6 * n0 * (1 + 1e-12) or its == 40)


def test_golden_fixture_gap_solve():
    """frozen output of tests/golden/make_golden.py:gap_solve (guards the oracle against drift; it does not pin the reference)"""
    import os
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "gap_c2_solve.npz"))
    cfg, o = make("C2", 1)
    sp_ = ob.make_solver_params(pre=2, post=2, bottom=4, max_iter=100, imin=10, iter_min=2, eps=1e-7, hang=1e-6, norm_thresh=1e-7)
    it, hist = o.solver.solve(o.F["b"], o.F["rhs"], sp_)
    assert np.array_equal(hist, z["resnorm"]) and np.array_equal(o.F["b"].get_global(), z["gap"])
    assert o.solver.bottom_iters == int(z["bottom_iters"])


# ---- size-independent properties of the implicit gap-height operator and its multigrid cycle ----
def _rand_field(layout, ng, rng, ghosts=False):
    f = ob.Field(layout, 1, ng)
    d = layout.domain
    ny, nx = d[3] + 1, d[2] + 1
    f.set_global(rng.rand(ny + 2 * ng, nx + 2 * ng), (-ng, -ng))
    return f


@pytest.mark.parametrize("name", ["C2", "C4", "C5"])
def test_operator_is_self_adjoint_and_positive(name):
    """L = alpha*a - beta*div(D grad) with FixedNeumBCFill / periodic sides is symmetric positive definite: <u, Lv> = <Lu, v>,
    <u, Lu> > 0 (what lets a linear multigrid with GSRB smoothing converge at all)"""
    cfg, o = make(name, 1)
    rng = np.random.RandomState(21)
    u, v = _rand_field(o.layout, 1, rng), _rand_field(o.layout, 1, rng)
    Lu, Lv = ob.Field(o.layout, 1, 0), ob.Field(o.layout, 1, 0)
    o.solver.applyOp(Lu, u)
    o.solver.applyOp(Lv, v)
    ug, vg, Lug, Lvg = u.get_global(), v.get_global(), Lu.get_global(), Lv.get_global()
    a, b = float(np.sum(ug * Lvg)), float(np.sum(Lug * vg))
    assert abs(a - b) <= 1e-11 * max(abs(a), abs(b))
    assert float(np.sum(ug * Lug)) > 0.0


@pytest.mark.parametrize("name", ["C2", "C5"])
def test_vcycle_commutes_with_power_of_two_scaling(name):
    """The correction-form cycle is linear in the residual, and scaling by a power of two is exact in binary floating point
    (the bottom solver's stop test is relative): V(4 r) == 4 V(r) bit for bit, with the same bottom pass count."""
    cfg, o = make(name, 1)
    rng = np.random.RandomState(8)
    r1 = _rand_field(o.layout, 0, rng)
    r4 = ob.Field(o.layout, 1, 0)
    r4.set_global(4.0 * r1.get_global(), (0, 0))
    sp_ = ob.make_solver_params(pre=2, post=2, bottom=4)
    c1, c4 = ob.Field(o.layout, 1, 1), ob.Field(o.layout, 1, 1)
    c1.setval(0.0)
    c4.setval(0.0)
    o.solver.vcycle(c1, r1, sp_)
    n1 = o.solver.bottom_iters
    o.solver.vcycle(c4, r4, sp_)
    assert o.solver.bottom_iters == n1
    assert np.array_equal(c4.get_global(), 4.0 * c1.get_global())


def test_restriction_is_the_scaled_adjoint_of_prolongation():
    """<P c, r>_fine = 4 <c, R r>_coarse for piecewise-constant prolongation P and 2x2 averaging R (restrictResidual with phi = 0)"""
    cfg, o = make("C2", 1)
    S = o.solver
    rng = np.random.RandomState(4)
    Lc = S.layout_at(1)
    r, c = _rand_field(o.layout, 0, rng), _rand_field(Lc, 0, rng)
    zero = ob.Field(o.layout, 1, 1)
    zero.setval(0.0)
    Rr = ob.Field(Lc, 1, 0)
    S.restrictResidual(Rr, zero, r)            # residual of phi = 0 is r itself
    Pc = ob.Field(o.layout, 1, 1)
    Pc.setval(0.0)
    S.prolongIncrement(Pc, c)
    lhs, rhs = float(np.sum(Pc.get_global() * r.get_global())), 4.0 * float(np.sum(c.get_global() * Rr.get_global()))
    assert abs(lhs - rhs) <= 1e-12 * abs(lhs)

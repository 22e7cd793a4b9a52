"""Shared set-up for the implicit gap-height solve tests (src/AmrHydro.cpp:594-662, 3378-3455): seeded
diffusion coefficients, gap height and right-hand side on a level, plus an independent numpy restatement of the stock
VCAMRPoissonOp2 formulas (whole-level arrays [j, i] with one ghost ring) used to cross-check the C oracle."""
import numpy as np

from oracle import binding as ob

XFACE, YFACE = 1, 2


def gap_problem(cfg, boxes, seed=5, dt=3600.0, diff_factor=1.0, dscale=None):
    """aCoef = 1 (aCoeff_GH), bCoef = D on faces (positive, rough), b = current gap height, rhs = b + dt * smooth forcing"""
    rng = np.random.RandomState(seed)
    nx, ny = cfg.nx, cfg.ny
    dx = cfg.dx[0]
    if dscale is None:
        dscale = 4.0 * dx * dx / (dt * diff_factor)   # beta*D/dx^2 ~ 4: diffusion matters on the finest grid
    g = {
        "a": np.ones((ny, nx)),
        "bX": dscale * (0.2 + rng.rand(ny, nx + 1)),
        "bY": dscale * (0.2 + rng.rand(ny + 1, nx)),
        "b": 0.01 + 0.05 * rng.rand(ny + 2, nx + 2),
        "rhs": 0.01 + 0.05 * rng.rand(ny, nx),
    }
    if cfg.periodic[0]:
        g["bX"][:, -1] = g["bX"][:, 0]
    if cfg.periodic[1]:
        g["bY"][-1, :] = g["bY"][0, :]
    return g, dx, 1.0, dt * diff_factor


class OracleGap:
    def __init__(self, cfg, boxes, **kw):
        self.cfg = cfg
        self.boxes = np.asarray(boxes, dtype=np.int32)
        self.layout = ob.Layout(self.boxes, (0, 0, cfg.nx - 1, cfg.ny - 1), cfg.periodic)
        self.g, self.dx, self.alpha, self.beta = gap_problem(cfg, boxes, **kw)
        F = {"a": ob.Field(self.layout, 1, 0), "bX": ob.Field(self.layout, 1, 0, XFACE), "bY": ob.Field(self.layout, 1, 0, YFACE),
             "b": ob.Field(self.layout, 1, 1), "rhs": ob.Field(self.layout, 1, 0)}
        for k in ("a", "bX", "bY", "rhs"):
            F[k].set_global(self.g[k], (0, 0))
        F["b"].set_global(self.g["b"], (-1, -1))
        self.F = F
        self.solver = ob.LinSolver(self.layout, self.dx, self.alpha, self.beta, F["a"], F["bX"], F["bY"])


# ---- numpy restatement (stock VCAMRPoissonOp2F.ChF formulas; FixedNeumBCFill src/AmrHydro.cpp:404-436) ----
def np_ghosts(p, cfg):
    q = p.copy()
    if cfg.periodic[0]:
        q[:, 0], q[:, -1] = q[:, -2], q[:, 1]
    else:
        q[1:-1, 0], q[1:-1, -1] = q[1:-1, 1], q[1:-1, -2]
    if cfg.periodic[1]:
        q[0, :], q[-1, :] = q[-2, :], q[1, :]
    else:
        q[0, 1:-1], q[-1, 1:-1] = q[1, 1:-1], q[-2, 1:-1]
    return q


def np_lofphi(p, g, dx, alpha, beta):
    c = p[1:-1, 1:-1]
    dxinv = 1.0 / (dx * dx)
    t = g["bX"][:, 1:] * (p[1:-1, 2:] - c)
    t = t - g["bX"][:, :-1] * (c - p[1:-1, :-2])
    t = t + g["bY"][1:, :] * (p[2:, 1:-1] - c)
    t = t - g["bY"][:-1, :] * (c - p[:-2, 1:-1])
    return alpha * g["a"] * c - beta * t * dxinv


def np_lambda(g, dx, alpha, beta):
    scale = 1.0 / (dx * dx)
    lam = g["a"] * alpha
    lam = lam + scale * beta * (g["bX"][:, 1:] + g["bX"][:, :-1])
    lam = lam + scale * beta * (g["bY"][1:, :] + g["bY"][:-1, :])
    return 1.0 / lam


def np_gsrb(p, rhs, g, cfg, dx, alpha, beta):
    ny, nx = rhs.shape
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    lam = np_lambda(g, dx, alpha, beta)
    for color in (0, 1):
        p = np_ghosts(p, cfg)
        lo = np_lofphi(p, g, dx, alpha, beta)
        new = p[1:-1, 1:-1] - lam * (lo - rhs)
        sel = ((ii + jj + color) % 2) == 0
        p[1:-1, 1:-1][sel] = new[sel]
    return p


class GpuGap:
    """Device twin of an OracleGap: same bits uploaded, GapHeightSolver on top."""

    def __init__(self, ctx, orc, owner=None):
        from suhmo_b200 import amr
        self.amr, self.ctx, self.orc = amr, ctx, orc
        cfg = orc.cfg
        self.layout = amr.DisjointBoxLayout(ctx, orc.boxes, (0, 0, cfg.nx - 1, cfg.ny - 1), cfg.periodic, owner)
        spec = dict(a=(0, 0), bX=(0, XFACE), bY=(0, YFACE), b=(1, 0), rhs=(0, 0))
        self.F = {}
        for k, (ng, cent) in spec.items():
            self.F[k] = amr.LevelData(self.layout, 1, ng, cent)
            self.push(self.F[k], orc.F[k])
        self.solver = amr.GapHeightSolver().define(ctx, [self.layout], [], (orc.dx, orc.dx), orc.alpha, [self.F["a"]], orc.beta,
                                                   [self.F["bX"]], [self.F["bY"]])

    def push(self, ld, of):
        ld.upload([of.fab(b)[0].copy() if ld.layout.owned(b) else None for b in range(len(ld.layout.boxes))])

    def new(self, depth=0, ng=0, src=None):
        L = self.layout if depth == 0 else self.layout.coarsen(1 << depth)
        f = self.amr.LevelData(L, 1, ng, 0)
        if src is not None:
            self.push(f, src)
        return f

"""Builds the same head-solve problem for the CPU oracle and (optionally) the GPU library from one set of
seeded synthetic fields, so parity tests feed both sides identical bits."""
import numpy as np

from oracle import binding as ob
from suhmo_b200 import synthetic as syn

CELL, XFACE, YFACE = 0, 1, 2
NAMES = ("head", "B", "Pi", "zb", "mask", "rhs", "a", "bX", "bY")


class OracleSide:
    def __init__(self, cfg, boxes, seed=12345, perturb=True, bc_vals=None, prm_over=None):
        self.cfg = cfg
        self.boxes = np.asarray(boxes, dtype=np.int32)
        self.domain = (0, 0, cfg.nx - 1, cfg.ny - 1)
        self.layout = ob.Layout(self.boxes, self.domain, cfg.periodic)
        g = syn.fields(cfg, ng=1, seed=seed, perturb=perturb)
        F = {}
        for k in ("head", "B", "Pi", "zb", "mask"):
            F[k] = ob.Field(self.layout, 1, 1)
            F[k].set_global(g[k], (-1, -1))
            # what AmrHydro does before the solve: exchange, and copy into domain ghosts (src/AmrHydro.cpp:2360-2445)
            ob.lib().orc_exchange_full(F[k].h)
            if k != "head":
                ob.lib().orc_copy_ghost(F[k].h)
        F["rhs"] = ob.Field(self.layout, 1, 0)
        F["rhs"].set_global(g["rhs"], (0, 0))
        F["a"] = ob.Field(self.layout, 1, 0)
        F["bX"] = ob.Field(self.layout, 1, 0, XFACE)
        F["bY"] = ob.Field(self.layout, 1, 0, YFACE)
        self.F = F
        lo_val, hi_val = bc_vals if bc_vals else ((0.0, 0.0), (0.0, 0.0))
        self.bc_vals = (lo_val, hi_val)
        self.bc = ob.make_bc(cfg.bc_lo, cfg.bc_hi, lo_val, hi_val)
        self.prm_kw = dict(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr,
                           cutOffBcoef=cfg.cutOffBcoef, use_mask_grad=cfg.use_mask_grad)
        if prm_over:
            self.prm_kw.update(prm_over)
        self.prm = ob.make_params(**self.prm_kw)
        self.alpha, self.beta = 0.0, -1.0

    def op(self):
        F = self.F
        return ob.Op(self.layout, self.cfg.dx, self.alpha, self.beta, self.bc, self.prm,
                     F["a"], F["bX"], F["bY"], F["B"], F["Pi"], F["zb"], F["mask"])

    def solver(self):
        F = self.F
        return ob.Solver(self.layout, self.cfg.dx, self.alpha, self.beta, self.bc, self.prm,
                         F["a"], F["bX"], F["bY"], F["B"], F["Pi"], F["zb"], F["mask"])

    def init_bcoef(self):
        """bCoef as the Picard body would hand it over: B(h) of the current head (aCoeff_bCoeff)."""
        o = self.op()
        o.update_operator(self.F["head"])
        return o


class GpuSide:
    """Device twin of an OracleSide: every box's FArrayBox is uploaded from the oracle's arrays."""

    def __init__(self, ctx, orc, owner=None):
        from suhmo_b200 import amr
        self.amr, self.ctx, self.orc = amr, ctx, orc
        cfg = orc.cfg
        self.layout = amr.DisjointBoxLayout(ctx, orc.boxes, orc.domain, cfg.periodic, owner)
        spec = dict(head=(1, CELL), B=(1, CELL), Pi=(1, CELL), zb=(1, CELL), mask=(1, CELL), rhs=(0, CELL),
                    a=(0, CELL), bX=(0, XFACE), bY=(0, YFACE))
        self.F = {}
        for k, (ng, cent) in spec.items():
            self.F[k] = amr.LevelData(self.layout, 1, ng, cent)
            self.push(k)
        self.bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi, *orc.bc_vals)
        self.prm = amr.make_params(**orc.prm_kw)
        self.factory = amr.VCAMRNonLinearPoissonOpFactory().define(
            ctx, [self.layout], [], cfg.dx, self.bc, orc.alpha, [self.F["a"]], orc.beta, [self.F["bX"]], [self.F["bY"]],
            self.prm, [self.F["B"]], [self.F["Pi"]], [self.F["zb"]], [self.F["mask"]])

    def push(self, name, src=None):
        """upload oracle field `name` (or the given oracle field) into the device field `name`"""
        of = src if src is not None else self.orc.F[name]
        fabs = []
        for b in range(len(self.layout.boxes)):
            fabs.append(of.fab(b)[0].copy() if self.layout.owned(b) else None)
        self.F[name].upload(fabs)

    def new_like(self, name):
        f = self.F[name]
        return self.amr.LevelData(self.layout, f.ncomp, f.ng, f.cent)


def fields_equal(gpu_ld, orc_field, valid_only=True):
    """max abs difference and exact-equality flag between a device field and an oracle field (valid cells)."""
    g = gpu_ld.get_global()
    o = orc_field.get_global()
    # NaN marks cells no box of the level holds (get_global's fill) -- and a diverged computation.  The two patterns must agree
    # before the finite values are compared: a NaN on the device where the oracle is finite is a failure, not a hole.
    if not np.array_equal(np.isnan(g), np.isnan(o)):
        return float("inf"), False
    m = ~np.isnan(o)
    diff = np.abs(g[m] - o[m])
    return (float(diff.max()) if diff.size else 0.0), bool(np.array_equal(g[m], o[m]))


def rel_l2(a, b):
    if not np.array_equal(np.isnan(a), np.isnan(b)):
        return float("inf")
    m = ~np.isnan(b)
    den = np.sqrt(np.sum(b[m] ** 2))
    return float(np.sqrt(np.sum((a[m] - b[m]) ** 2)) / (den if den > 0 else 1.0))


# ---------------------------------------------------------------------------------------------------------------
# multi-level (AMR) problems: explicit box lists per level (what Chombo reads from AmrHydro.grids_file), ref ratio 2
# ---------------------------------------------------------------------------------------------------------------
def amr_hierarchy(name="C5"):
    """Explicit test hierarchies (boxes are lo0 lo1 hi0 hi1, block-factor 8 aligned, refinement ratio 2).
    "C5":     3 levels on a 64x64 base grid: L-shaped level 1 (concave corner, box-box exchange, a box on the domain boundary),
              level 2 properly nested inside it.
    "C4":     2 levels on the valley geometry (256x64, exec/E_SHMIP): ice mask < 0 inside the refined boxes AND across the
              coarse-fine interface, masked gradients (solver.use_mask_for_gradients), cut_solve_outside_domain -- the BASELINE
              config where the masked coarse-fine gradient and the mask-streaming smoother matter.
    "C5_256": 3 levels on the AMR_multiMoulins base grid at its native 256x256 size (16 boxes of 64^2), the same shapes scaled.
    "C5_BR":  3 levels on the 64x64 base grid as AmrHydro::regrid leaves them (tags on Pi > 5e6, fill ratio 0.7, block factor 8, nesting
              radius 2, max box 32): Berger-Rigoutsos shapes -- 16x8 and 8x16 boxes, re-entrant corners one box wide, refined boxes
              along two physical boundaries on both levels."""
    if name == "C4":
        cfg = syn.config("C4", 1)
        base = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        lev1 = np.array([(16, 16, 79, 79), (80, 16, 143, 79), (16, 80, 79, 111), (448, 32, 511, 95)], dtype=np.int32)
        return cfg, [base, lev1]
    if name == "C5_256":
        cfg = syn.config("C5", 1)
        base = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        lev1 = np.array([(64, 64, 127, 127), (128, 64, 191, 127), (64, 128, 127, 191), (448, 0, 511, 63)], dtype=np.int32)
        lev2 = np.array([(160, 160, 223, 223), (224, 160, 287, 223), (160, 224, 223, 287)], dtype=np.int32)
        return cfg, [base, lev1, lev2]
    if name == "C5_BR":
        cfg, lv = amr_hierarchy("C5")
        lev1 = np.array([(64, 0, 95, 15), (96, 0, 127, 31), (80, 16, 95, 23), (88, 24, 95, 39), (96, 32, 127, 63)], dtype=np.int32)
        lev2 = np.array([(184, 0, 215, 31), (216, 0, 231, 31), (232, 0, 255, 31), (184, 32, 215, 47), (216, 32, 231, 47), (232, 32, 255, 47),
                         (184, 48, 215, 71), (216, 48, 231, 71), (232, 48, 255, 71)], dtype=np.int32)
        return cfg, [lv[0], lev1, lev2]
    cfg = syn.config(name, 1)
    cfg.nx, cfg.ny = 64, 64
    cfg.max_box_size = 32
    base = syn.domain_split(64, 64, 32, 2)
    lev1 = np.array([(16, 16, 47, 47), (48, 16, 79, 47), (16, 48, 47, 79), (96, 0, 127, 31)], dtype=np.int32)
    lev2 = np.array([(48, 48, 79, 79), (80, 48, 111, 79), (48, 80, 79, 111)], dtype=np.int32)
    return cfg, [base, lev1, lev2]


class AmrOracleSide:
    """Per-level oracle fields of a head-solve problem on an explicit hierarchy.  boxwise: evaluate the synthetic fields box
    group by box group (syn.box_fields) instead of on the level's whole index space -- same bits, and the only way at sizes where
    a global fine-level array would not fit."""

    def __init__(self, cfg, level_boxes, seed=12345, bc_vals=None, prm_over=None, boxwise=False, periodic_ghosts=False):
        self.cfg, self.level_boxes = cfg, [np.asarray(b, dtype=np.int32) for b in level_boxes]
        self.nlev = len(level_boxes)
        self.layouts, self.F, self.dx = [], [], []
        for l, boxes in enumerate(self.level_boxes):
            r = 2 ** l
            dom = (0, 0, cfg.nx * r - 1, cfg.ny * r - 1)
            lay = ob.Layout(boxes, dom, cfg.periodic)
            F = {k: ob.Field(lay, 1, 1) for k in ("head", "B", "Pi", "zb", "mask")}
            F["rhs"] = ob.Field(lay, 1, 0)
            if boxwise:
                for b, f in syn.box_fields(cfg, boxes, level_ratio=r, ng=1, seed=seed, periodic_ghosts=periodic_ghosts):
                    for k in ("head", "B", "Pi", "zb", "mask", "rhs"):
                        F[k].fab(b)[0][0][...] = f[k]
            else:
                g = syn.fields(cfg, ng=1, seed=seed, level_ratio=r)
                for k in ("head", "B", "Pi", "zb", "mask"):
                    F[k].set_global(g[k], (-1, -1))   # ghosts (CF ghosts included) from the analytic fields
                F["rhs"].set_global(g["rhs"], (0, 0))
            for k in ("B", "Pi", "zb", "mask"):
                ob.lib().orc_copy_ghost(F[k].h)
            F["a"] = ob.Field(lay, 1, 0)
            F["bX"] = ob.Field(lay, 1, 0, XFACE)
            F["bY"] = ob.Field(lay, 1, 0, YFACE)
            self.layouts.append(lay)
            self.F.append(F)
            self.dx.append((cfg.dx[0] / r, cfg.dx[1] / r))
        lo_val, hi_val = bc_vals if bc_vals else ((0.0, 0.0), (0.0, 0.0))
        self.bc_vals = (lo_val, hi_val)
        self.bc = ob.make_bc(cfg.bc_lo, cfg.bc_hi, lo_val, hi_val)
        self.prm_kw = dict(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr,
                           cutOffBcoef=cfg.cutOffBcoef, use_mask_grad=cfg.use_mask_grad)
        if prm_over:
            self.prm_kw.update(prm_over)
        self.prm = ob.make_params(**self.prm_kw)
        self.alpha, self.beta = 0.0, -1.0

    def fields(self, name):
        return [F[name] for F in self.F]

    def level_op(self, l):
        F = self.F[l]
        return ob.Op(self.layouts[l], self.dx[l], self.alpha, self.beta, self.bc, self.prm,
                     F["a"], F["bX"], F["bY"], F["B"], F["Pi"], F["zb"], F["mask"])

    def average_down(self, name="head"):
        """fine -> coarse average of the covered region (AmrHydro does this before the solve, src/AmrHydro.cpp:3139)"""
        for l in range(self.nlev - 1, 0, -1):
            clay = self.layouts[l].coarsen(2)
            tmp = ob.Field(clay, 1, 0)
            ob.lib().orc_coarse_average(self.F[l][name].h, tmp.h, 2)
            ob.copy_to(self.F[l - 1][name], tmp)

    def init_bcoef(self):
        """bCoef = B(h) of the current head on every level (what aCoeff_bCoeff hands to the solver)"""
        for l in range(self.nlev):
            op = self.level_op(l)
            if l == 0:
                op.update_operator(self.F[0]["head"])
            else:
                ob.cf_interp(self.F[l]["head"], self.F[l - 1]["head"], 2, self.dx[l][0])
                op.update_operator_amr(self.F[l]["head"], self.F[l - 1]["head"], self.F[l - 1]["mask"])

    def solver(self):
        f = self.fields
        return ob.AmrSolver(self.layouts, self.cfg.dx, self.alpha, self.beta, self.bc, self.prm,
                            f("a"), f("bX"), f("bY"), f("B"), f("Pi"), f("zb"), f("mask"))


class AmrGpuSide:
    """Device twin of an AmrOracleSide.  owners: per level, the rank owning each box (None: everything on this rank)."""

    def __init__(self, ctx, orc, owners=None):
        from suhmo_b200 import amr
        self.amr, self.ctx, self.orc = amr, ctx, orc
        cfg = orc.cfg
        spec = dict(head=(1, CELL), B=(1, CELL), Pi=(1, CELL), zb=(1, CELL), mask=(1, CELL), rhs=(0, CELL),
                    a=(0, CELL), bX=(0, XFACE), bY=(0, YFACE))
        self.layouts, self.F = [], []
        for l in range(orc.nlev):
            r = 2 ** l
            lay = amr.DisjointBoxLayout(ctx, orc.level_boxes[l], (0, 0, cfg.nx * r - 1, cfg.ny * r - 1), cfg.periodic,
                                        None if owners is None else owners[l])
            F = {k: amr.LevelData(lay, 1, ng, cent) for k, (ng, cent) in spec.items()}
            self.layouts.append(lay)
            self.F.append(F)
            for k in spec:
                self.push(l, k)
        self.bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi, *orc.bc_vals)
        self.prm = amr.make_params(**orc.prm_kw)
        f = self.fields
        self.factory = amr.VCAMRNonLinearPoissonOpFactory().define(
            ctx, self.layouts, [2] * (orc.nlev - 1), cfg.dx, self.bc, orc.alpha, f("a"), orc.beta, f("bX"), f("bY"),
            self.prm, f("B"), f("Pi"), f("zb"), f("mask"))

    def fields(self, name):
        return [F[name] for F in self.F]

    def push(self, l, name, src=None):
        of = src if src is not None else self.orc.F[l][name]
        lay = self.layouts[l]
        self.F[l][name].upload([of.fab(b)[0].copy() if lay.owned(b) else None for b in range(len(lay.boxes))])

    def new_like(self, l, name):
        f = self.F[l][name]
        return self.amr.LevelData(self.layouts[l], f.ncomp, f.ng, f.cent)


def fabs_equal(gpu_ld, orc_field, ghosts=True):
    """compare whole FArrayBoxes box by box (ghost cells included when both sides carry them)"""
    worst, same = 0.0, True
    outs = gpu_ld.download()
    for b in range(len(orc_field.layout.boxes)):
        o = orc_field.fab(b)[0]
        g = outs[b]
        if not ghosts and gpu_ld.ng:
            n = gpu_ld.ng
            o, g = o[:, n:-n, n:-n], g[:, n:-n, n:-n]
        same &= bool(np.array_equal(o, g))
        worst = max(worst, float(np.abs(o - g).max()))
    return worst, same

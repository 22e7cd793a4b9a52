"""Deterministic synthetic inputs for the head solve.

Host-side input generators only (numpy): the closed-form initial fields of the reference's IBC
classes, a seedable perturbation so that the solve has something to do, a plausible right-hand
side, and Chombo's `domainSplit` box generation.  Nothing here is on the hot path.

Field conventions: every array is [j, i] (x fastest), carries `ng` ghost cells on every side and
is indexed so that element [ng, ng] is cell (0, 0).
"""
import math
from dataclasses import dataclass, field

import numpy as np

RHO_I, RHO_W, GRAV = 910.0, 1000.0, 9.8


@dataclass
class Config:
    """One head-solve problem: geometry, BCs, physics constants (input.hydro keys)."""
    name: str
    ibc: str                      # basic | sqrt | valley | dino
    nx: int
    ny: int
    domain_size: tuple
    periodic: tuple
    bc_lo: tuple
    bc_hi: tuple
    max_box_size: int
    block_factor: int = 1
    A: float = 2.5e-25
    omega: float = 1e-3
    nu: float = 1.787e-6
    cutOffbr: float = 0.0
    maxOffbr: float = 10000.0
    cutOffBcoef: int = 0
    use_mask_grad: int = 0
    H: float = 500.0
    slope: float = 0.02
    gap_init: float = 0.01
    gamma: float = 0.05
    distributed_input: float = 1e-11
    moulins: list = field(default_factory=list)  # (x, y, flux, sigma)
    bc_lo_val: tuple = (0.0, 0.0)
    bc_hi_val: tuple = (0.0, 0.0)

    @property
    def dx(self):
        return (self.domain_size[0] / self.nx, self.domain_size[1] / self.ny)


def config(name, scale=1):
    """The five BASELINE.json configs (level-0 grids).  `scale` multiplies the resolution."""
    if name == "C1":  # exec/0_convergence_channelized/1lev/input.hydro
        return Config("C1", "basic", 32 * scale, 8 * scale, (64.0, 16.0), (0, 1), (0, 0), (1, 0), 16,
                      block_factor=8, moulins=[(16.015625, 8.015625, 30.0, 1.0)])
    if name == "C2":  # exec/1_convergence_distributed
        return Config("C2", "sqrt", 64 * scale, 16 * scale, (80000.0, 20000.0), (0, 1), (0, 1), (1, 1), 64,
                      block_factor=2, A=2.5e-25, H=5000.0, slope=0.0, distributed_input=5.79e-9)
    if name == "C3":  # exec/A_SHMIP/A3
        return Config("C3", "sqrt", 320 * scale, 64 * scale, (100000.0, 20000.0), (0, 0), (0, 1), (1, 1), 64,
                      block_factor=2, A=5e-25, H=5000.0, slope=0.0, distributed_input=5.79e-9)
    if name == "C4":  # exec/E_SHMIP/E1
        return Config("C4", "valley", 256 * scale, 64 * scale, (6000.0, 1500.0), (0, 0), (0, 1), (1, 1), 64,
                      block_factor=8, A=5e-25, slope=0.0, cutOffBcoef=1, use_mask_grad=1, distributed_input=1.158e-6, gamma=0.05)
    if name == "C5":  # exec/AMR_multiMoulins/run_C_3lev (base level)
        n = 256 * scale
        rng = np.random.RandomState(63)
        mo = [(float(x), float(y), 80.0, 200.0)
              for x, y in rng.uniform(10000.0, 90000.0, size=(63, 2))]
        return Config("C5", "dino", n, n, (100000.0, 100000.0), (0, 0), (0, 1), (1, 0), 64,
                      block_factor=2, H=400.0, slope=0.0, distributed_input=9e-10, moulins=mo)
    raise ValueError(name)


def domain_split(nx, ny, max_size, block_factor=1):
    """Chombo domainSplit (absent): bisect per direction until <= max_size."""
    def split1(n):
        segs = [(0, n // block_factor - 1)]
        lim = max(1, max_size // block_factor)
        changed = True
        while changed:
            changed = False
            out = []
            for lo, hi in segs:
                if hi - lo + 1 > lim:
                    mid = (lo + hi) // 2 + 1  # low part keeps [lo, (lo+hi)/2], high part starts at (lo+hi)/2+1
                    out += [(lo, mid - 1), (mid, hi)]
                    changed = True
                else:
                    out.append((lo, hi))
            segs = out
        return [(lo * block_factor, (hi + 1) * block_factor - 1) for lo, hi in segs]
    xs, ys = split1(nx), split1(ny)
    # DisjointBoxLayout sorts boxes lexicographically; ordering affects ownership only, never arithmetic.
    boxes = [(x0, y0, x1, y1) for (y0, y1) in ys for (x0, x1) in xs]
    return np.array(boxes, dtype=np.int32)


def _hash_uniform(i, j, seed):
    """counter-based hash -> uniform[-1, 1): independent of traversal order (replaces std::normal_distribution noise)."""
    h = (i.astype(np.uint64) * np.uint64(0x9E3779B97F4A7C15)) ^ (j.astype(np.uint64) * np.uint64(0xC2B2AE3D27D4EB4F)) \
        ^ np.uint64(seed * 0x165667B19E3779F9 & 0xFFFFFFFFFFFFFFFF)
    h ^= h >> np.uint64(33)
    h *= np.uint64(0xFF51AFD7ED558CCD)
    h ^= h >> np.uint64(33)
    h *= np.uint64(0xC4CEB9FE1A85EC53)
    h ^= h >> np.uint64(33)
    return (h >> np.uint64(11)).astype(np.float64) * (2.0 / float(1 << 53)) - 1.0


def fields(cfg, ng=1, seed=12345, perturb=True, lo=(0, 0), shape=None, level_ratio=1, moulin_cutoff=None, rhs_only=False, periodic_ghosts=False):
    """IBC fields on index window [lo, lo+shape) (default: the whole level-0 domain), with ng ghosts.

    level_ratio refines dx (AMR level = base dx / level_ratio).  Returns dict of [j, i] arrays:
    head, B (gap height), Pi, zb, mask, rhs (no ghosts).
    moulin_cutoff (in units of a moulin's sigma): evaluate each Gaussian moulin source only where it is not negligible; with
    cutoff >= 12 the omitted terms are below one ulp of the background recharge, i.e. the sum is unchanged.
    rhs_only: only the right-hand side (same bits as the full call), for geometries without an ice mask.
    periodic_ghosts: in periodic directions, cells outside the domain take the values of their periodic images (indices wrapped
    before the formulas are evaluated) instead of the formulas' continuation -- the state LevelData::exchange leaves behind.
    """
    nx, ny = shape if shape is not None else (cfg.nx * level_ratio, cfg.ny * level_ratio)
    dx, dy = cfg.dx[0] / level_ratio, cfg.dx[1] / level_ratio
    with np.errstate(over="ignore"):
        ii = np.arange(lo[0] - ng, lo[0] + nx + ng, dtype=np.int64)
        jj = np.arange(lo[1] - ng, lo[1] + ny + ng, dtype=np.int64)
        i_org, j_org = int(ii[0]), int(jj[0])
        if periodic_ghosts:
            if cfg.periodic[0]:
                ii = np.mod(ii, cfg.nx * level_ratio)
            if cfg.periodic[1]:
                jj = np.mod(jj, cfg.ny * level_ratio)
        I, J = np.meshgrid(ii, jj)
        x, y = (I + 0.5) * dx, (J + 0.5) * dy
        mask = np.ones_like(x)
        if rhs_only:
            if cfg.ibc == "valley":
                raise ValueError("rhs_only: the valley geometry masks the right-hand side")
            head = B = Pi = zb = None
        elif cfg.ibc == "basic":      # src/HydroIBC.cpp:230-271
            zb = cfg.slope * x
            B = np.full_like(x, cfg.gap_init)
            Pi = RHO_I * GRAV * np.full_like(x, cfg.H)
            head = Pi * 0.5 * (1.0 / (RHO_W * GRAV)) + zb
        elif cfg.ibc == "sqrt":     # src/SqrtIBC.cpp:219-281
            zb = cfg.slope * x
            Hh = np.maximum(6.0 * (np.sqrt(np.maximum(x + cfg.H, 0.0)) - math.sqrt(cfg.H)) + 1.0, 0.0)
            Pi = np.maximum(RHO_I * GRAV * Hh, 0.0)
            B = np.where(Pi < 2.0, 1e-16, cfg.gap_init)
            head = 101325 * (1.0 / (RHO_W * GRAV)) + zb
        elif cfg.ibc == "valley":   # src/ValleyIBC.cpp:225-342
            yl = y - 750.0
            xs = np.maximum(x, -199.0)
            Hs = 100.0 * np.power(xs + 200.0, 0.25) + x / 60.0 - 2.0e10 ** 0.25 + 1.0
            H6 = 100.0 * (6000.0 + 200.0) ** 0.25 + 6000.0 / 60.0 - 2.0e10 ** 0.25 + 1.0
            fx = (H6 - 6000.0 * cfg.gamma) * x * x / (6000.0 * 6000.0) + cfg.gamma * x
            fxg = (H6 - 6000.0 * 0.05) * x * x / (6000.0 * 6000.0) + 0.05 * x
            gy = 0.5e-6 * np.abs(yl * yl * yl)
            hx = (-4.5 * x / 6000.0 + 5.0) * (Hs - fx) / (Hs - fxg + 1e-16)
            zb = fx + gy * hx
            Pi = RHO_I * GRAV * np.maximum(Hs - zb, 0.0)
            B = np.where(Pi < 1e-10, 1e-16, cfg.gap_init)
            head = Pi * 0.5 * (1.0 / (RHO_W * GRAV)) + zb
            mask = np.where(Pi > 0.0, 1.0, -1.0)
        elif cfg.ibc == "dino":     # src/MountainSetupIBC.cpp:153-317 minus the traversal-order-dependent noise
            def bump(x0, y0, ang, amax, sy, sig):
                a = ang * 3.14159 / 180.0
                xb, yb = x - x0, y - y0
                xt = xb * math.cos(a) + yb * math.sin(a)
                yt = -xb * math.sin(a) + yb * math.cos(a)
                s = sig(yt) if callable(sig) else sig
                return amax * np.exp(-0.5 / (sy * sy) * yt * yt) * np.exp(-0.5 / (s * s) * xt * xt)
            s1 = np.maximum(1.5e-3 * x - 1.5e-3 * y + 100.0, 0.0)
            Hh = 2.0 * (1.5e-3 * x - 1.5e-3 * y + 100.0 + 100.0)
            zb = (s1 + bump(100000.0, 0.0, 35.0, 250.0, 50000.0,
                            lambda yt: 12000.0 - 3000.0 * np.minimum(1.0 - (50000.0 - yt) / 50000.0, 1.0))
                  + bump(85000.0, 0.0, 90.0, 250.0, 30000.0, 10000.0)
                  + bump(55000.0, 0.0, 60.0, 100.0, 20000.0, 5000.0)
                  + bump(100000.0, 20000.0, 2.0, 300.0, 35000.0, 6000.0))
            zb = np.maximum(zb + _hash_uniform(I, J, seed + 7), 0.0)
            Pi = RHO_I * GRAV * np.maximum(Hh, 0.0)
            B = np.where(Pi == 0.0, 1e-16, cfg.gap_init)
            head = Pi * 0.5 * (1.0 / (RHO_W * GRAV)) + zb
        else:
            raise ValueError(cfg.ibc)
        if perturb and not rhs_only:
            Lx, Ly = cfg.domain_size
            head = head * (1.0 + 1e-3 * np.sin(2 * np.pi * 3 * x / Lx) * np.cos(2 * np.pi * 2 * y / Ly))
            B = B * (1.0 + 0.5 * _hash_uniform(I, J, seed))
        # right-hand side of the head equation: recharge + moulins (Gaussian), cf. src/AmrHydro.cpp:2801-3079
        rhs = np.full_like(x, cfg.distributed_input)
        for (mx, my, flux, sig) in cfg.moulins:
            if moulin_cutoff is None:
                rhs += flux / (2.0 * np.pi * sig * sig) * np.exp(-0.5 * ((x - mx) ** 2 + (y - my) ** 2) / (sig * sig))
                continue
            w = moulin_cutoff * sig
            i0 = max(0, int(math.floor((mx - w) / dx - 0.5)) - i_org); i1 = min(len(ii), int(math.ceil((mx + w) / dx - 0.5)) + 1 - i_org)
            j0 = max(0, int(math.floor((my - w) / dy - 0.5)) - j_org); j1 = min(len(jj), int(math.ceil((my + w) / dy - 0.5)) + 1 - j_org)
            if i1 <= i0 or j1 <= j0:
                continue
            xs, ys = x[j0:j1, i0:i1], y[j0:j1, i0:i1]
            rhs[j0:j1, i0:i1] += flux / (2.0 * np.pi * sig * sig) * np.exp(-0.5 * ((xs - mx) ** 2 + (ys - my) ** 2) / (sig * sig))
        rhs = np.where(mask < 0.0, 0.0, rhs)  # no recharge outside the ice (rhs/1e-16 would blow up where lambda = 0)
    core = (slice(ng, ng + ny), slice(ng, ng + nx))
    if rhs_only:
        return dict(rhs=np.ascontiguousarray(rhs[core]))
    return dict(head=head, B=B, Pi=Pi, zb=zb, mask=mask, rhs=np.ascontiguousarray(rhs[core]))


def box_fields(cfg, boxes, level_ratio=1, ng=1, seed=12345, bin_cells=512, moulin_cutoff=12.0, rhs_only=False, periodic_ghosts=False):
    """fields() for a list of boxes of one AMR level without ever forming the level's global arrays: boxes are grouped by the
    bin (bin_cells x bin_cells fine cells) their low corner falls into, the fields are evaluated once on each group's bounding
    window and cut into the boxes' FArrayBoxes.  Yields (box index, dict name -> [j, i] array with ng ghosts; rhs without)."""
    boxes = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
    groups = {}
    for b, bx in enumerate(boxes):
        groups.setdefault((int(bx[1]) // bin_cells, int(bx[0]) // bin_cells), []).append(b)
    for key in sorted(groups):
        ids = groups[key]
        sub = boxes[ids]
        x0, y0, x1, y1 = int(sub[:, 0].min()), int(sub[:, 1].min()), int(sub[:, 2].max()), int(sub[:, 3].max())
        g = fields(cfg, ng=ng, seed=seed, lo=(x0, y0), shape=(x1 - x0 + 1, y1 - y0 + 1), level_ratio=level_ratio, moulin_cutoff=moulin_cutoff,
                   rhs_only=rhs_only, periodic_ghosts=periodic_ghosts)
        for b in ids:
            bx = boxes[b]
            i0, j0 = int(bx[0]) - x0, int(bx[1]) - y0
            nx, ny = int(bx[2] - bx[0] + 1), int(bx[3] - bx[1] + 1)
            out = {} if rhs_only else {k: g[k][j0:j0 + ny + 2 * ng, i0:i0 + nx + 2 * ng] for k in ("head", "B", "Pi", "zb", "mask")}
            out["rhs"] = g["rhs"][j0:j0 + ny, i0:i0 + nx]
            yield b, out

"""The bench workload: BASELINE.json configs[4], "AMR_multiMoulins 3-level AMR scaled to a synthetic 8192^2 base grid".

One TILE = the AMR_multiMoulins problem (exec/AMR_multiMoulins/run_C_3lev/input.hydro: MountainIBC-like geometry, 63 moulins,
Dirichlet / Neumann sides) on a size x size base grid of 64^2 boxes, with two refined levels built from tags exactly as the input file asks
(fill_ratio 0.5, block_factor 2, nestingRadius 4, max_box_size 64, tags_grow 4).

  weak scaling   : N tiles stacked in y, one per GPU, every tile built from the same tile-local formulas (so every GPU has the same
                   grids and the same amount of work); the stack is one problem with the physical boundary conditions of the input
                   file on its outer sides (the configuration is not periodic) and ordinary box-to-box coupling between tiles.
  strong scaling : one tile, base level cut into N y-strips, refined boxes dealt out cluster by cluster (connected groups of
                   boxes stay on one GPU; clusters are balanced by cell count) -- Chombo's LoadBalance with the constraint that
                   same-level neighbours share a rank.

Both bench arms (GPU library, CPU oracle) build their fields from this module, so they solve the same problem on the same grids.
Nothing here is on the hot path."""
import numpy as np

from suhmo_b200 import synthetic as syn

BOX = 64
TAG_LEVEL0 = 20.0     # tag where rhs > TAG_LEVEL0 x background recharge (the moulins' footprint) ...
TAG_LEVEL1 = 200.0    # ... and their cores on level 1
REGRID = dict(fill_ratio=0.5, block_factor=2, nesting_radius=4, max_box_size=64, tags_grow=4)


def tile_config(size):
    cfg = syn.config("C5", 1)
    cfg.nx = cfg.ny = size
    return cfg


def global_config(size, ntiles):
    cfg = tile_config(size)
    cfg.ny = size * ntiles
    cfg.domain_size = (cfg.domain_size[0], cfg.domain_size[1] * ntiles)
    return cfg


def level_rhs_fabs(cfg, boxes, level):
    """right-hand side FArrayBoxes of `boxes` (tile-local indices) on AMR level `level`, in box order"""
    out = [None] * len(boxes)
    for b, f in syn.box_fields(cfg, boxes, level_ratio=2 ** level, ng=1, rhs_only=True):
        out[b] = f["rhs"]
    return out


def build_tile_hierarchy(cfg, tag_level, regrid, nlevels=3):
    """Box lists of levels 0..nlevels-1 of one tile.  tag_level(level, boxes, rhs_fabs, vmin) -> uint8 tag map of that level
    (the arm's own tagging: sg_tag_cells_level on the device, orc_tag_cells_level in the oracle); regrid(base, [tags...]) -> levels."""
    base = syn.domain_split(cfg.nx, cfg.ny, BOX, cfg.block_factor)
    bg = cfg.distributed_input
    levels, tags = [base], []
    for l in range(nlevels - 1):
        thr = (TAG_LEVEL0 if l == 0 else TAG_LEVEL1) * bg
        tags.append(tag_level(l, levels[l], level_rhs_fabs(cfg, levels[l], l) if l > 0 else None, thr))
        new = regrid(base, tags)
        if len(new) < l + 2:
            break
        levels = new
    return levels


def components(boxes):
    """connected groups of boxes (sharing an edge or a corner), as a label per box"""
    boxes = np.asarray(boxes, dtype=np.int64).reshape(-1, 4)
    n = len(boxes)
    parent = np.arange(n)

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a
    bins = {}
    for b, (x0, y0, x1, y1) in enumerate(boxes):
        for by in range((y0 - 1) // BOX, (y1 + 1) // BOX + 1):
            for bx in range((x0 - 1) // BOX, (x1 + 1) // BOX + 1):
                bins.setdefault((bx, by), []).append(b)
    for ids in bins.values():
        for p in range(len(ids)):
            a = boxes[ids[p]]
            for q in range(p + 1, len(ids)):
                c = boxes[ids[q]]
                if a[0] <= c[2] + 1 and c[0] <= a[2] + 1 and a[1] <= c[3] + 1 and c[1] <= a[3] + 1:
                    ra, rc = find(ids[p]), find(ids[q])
                    if ra != rc:
                        parent[ra] = rc
    return np.array([find(b) for b in range(n)])


def cluster_balance(levels, nranks, rows0):
    """owner per box for every level: level 0 by y-strips of rows0 rows; refined levels cluster-wise.  A level-1 cluster and the
    finer boxes nested in it form one unit (inter-level copies stay local); units go, largest first, to the rank under them when
    that rank is not yet over its share, else to the least loaded rank."""
    owners = [(np.asarray(levels[0])[:, 1] // rows0).astype(np.int32)]
    if len(levels) == 1:
        return owners
    if nranks == 1:
        return owners + [np.zeros(len(b), dtype=np.int32) for b in levels[1:]]
    l1 = np.asarray(levels[1], dtype=np.int64)
    lab1 = components(l1)
    cells = {}
    for b, lab in enumerate(lab1):
        cells[lab] = cells.get(lab, 0) + int((l1[b, 2] - l1[b, 0] + 1) * (l1[b, 3] - l1[b, 1] + 1))
    unit_of_finer = []
    for l in range(2, len(levels)):
        bl = np.asarray(levels[l], dtype=np.int64)
        sh = l - 1
        # the level-1 box under each finer box's low corner names its unit
        order = np.lexsort((l1[:, 0], l1[:, 1]))
        keys = {}
        for b in order:
            keys.setdefault((int(l1[b, 0]) // BOX, int(l1[b, 1]) // BOX), []).append(b)
        u = np.zeros(len(bl), dtype=np.int64)
        for k, bx in enumerate(bl):
            x, y = int(bx[0]) >> sh, int(bx[1]) >> sh
            found = -1
            for by in (y // BOX, y // BOX - 1):
                for bxx in (x // BOX, x // BOX - 1):
                    for b in keys.get((bxx, by), []):
                        if l1[b, 0] <= x <= l1[b, 2] and l1[b, 1] <= y <= l1[b, 3]:
                            found = b
                            break
                    if found >= 0:
                        break
                if found >= 0:
                    break
            if found < 0:
                raise RuntimeError("cluster_balance: a level-%d box is not nested in level 1" % l)
            u[k] = lab1[found]
            cells[u[k]] += int((bx[2] - bx[0] + 1) * (bx[3] - bx[1] + 1))
        unit_of_finer.append(u)
    share = sum(cells.values()) / nranks
    load = np.zeros(nranks)
    rank_of = {}
    ymid = {}
    for b, lab in enumerate(lab1):
        ymid.setdefault(lab, []).append((l1[b, 1] + l1[b, 3]) // 4)  # level-0 row
    for lab in sorted(cells, key=lambda k: (-cells[k], k)):
        home = min(nranks - 1, int(np.median(ymid[lab])) // rows0)
        r = home if load[home] + cells[lab] <= 1.05 * share else int(np.argmin(load))
        rank_of[lab] = r
        load[r] += cells[lab]
    owners.append(np.array([rank_of[lab] for lab in lab1], dtype=np.int32))
    for u in unit_of_finer:
        owners.append(np.array([rank_of[k] for k in u], dtype=np.int32))
    return owners


class Problem:
    """global grids + owners of a bench run"""

    def __init__(self, size, nranks, scaling, tile_levels):
        self.size, self.nranks, self.scaling = size, nranks, scaling
        self.tile_cfg = tile_config(size)
        self.ntiles = nranks if scaling == "weak" else 1
        self.cfg = global_config(size, self.ntiles)
        self.tile_levels = [np.asarray(b, dtype=np.int32) for b in tile_levels]
        self.nlev = len(tile_levels)
        if scaling == "weak":
            self.levels, self.owners = [], []
            for l, boxes in enumerate(self.tile_levels):
                sh = size << l
                self.levels.append(np.concatenate([boxes + np.array([0, k * sh, 0, k * sh], dtype=np.int32) for k in range(nranks)]).astype(np.int32))
                self.owners.append(np.repeat(np.arange(nranks, dtype=np.int32), len(boxes)))
        else:
            if size % nranks or (size // nranks) % BOX:
                raise ValueError("strong scaling: the base grid must split into strips of whole boxes")
            self.levels = self.tile_levels
            self.owners = cluster_balance(self.levels, nranks, size // nranks)

    def domain(self, l):
        return (0, 0, (self.cfg.nx << l) - 1, (self.cfg.ny << l) - 1)

    def owned(self, l, rank):
        return np.flatnonzero(self.owners[l] == rank)

    def to_tile(self, l, boxes):
        """global box indices -> tile-local (weak scaling: every tile is the same problem)"""
        if self.scaling != "weak":
            return boxes
        sh = self.size << l
        b = np.array(boxes, dtype=np.int64, copy=True)
        k = b[:, 1] // sh
        b[:, 1] -= k * sh
        b[:, 3] -= k * sh
        return b

    def level_fabs(self, l, rank, names=("head", "rhs", "B", "Pi", "zb", "mask")):
        """FArrayBoxes of the boxes `rank` owns on level l: dict name -> list over ALL boxes of the level (None where not owned)"""
        ids = self.owned(l, rank)
        out = {k: [None] * len(self.levels[l]) for k in names}
        if len(ids) == 0:
            return out
        local = self.to_tile(l, self.levels[l][ids])
        for b, f in syn.box_fields(self.tile_cfg, local, level_ratio=2 ** l, ng=1, periodic_ghosts=True):
            for k in names:
                out[k][ids[b]] = f[k]
        return out

    def cells(self):
        return [int(sum((b[2] - b[0] + 1) * (b[3] - b[1] + 1) for b in boxes)) for boxes in self.levels]

    def describe(self):
        c = self.cells()
        per_rank = [[int(sum((b[2] - b[0] + 1) * (b[3] - b[1] + 1) for b in self.levels[l][self.owned(l, r)])) for r in range(self.nranks)]
                    for l in range(self.nlev)]
        return {"levels": self.nlev, "boxes": [int(len(b)) for b in self.levels], "cells": c, "cells_per_rank": per_rank}

#!/usr/bin/env python
"""Micro-benchmark of the finest-level smoother (levelGSRB) under the library's tuning knobs.
CUDA events on the library's stream; inputs larger than L2.  Usage: python tools/relax_bench.py [size]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from suhmo_b200 import amr, synthetic as syn  # noqa: E402


def single_level_setup(size):
    """the bench workload's base level alone (AMR_multiMoulins, size^2 cells, 64^2 boxes, one GPU)"""
    from tools import workload as wl
    cfg = wl.tile_config(size)
    boxes = syn.domain_split(cfg.nx, cfg.ny, wl.BOX, cfg.block_factor)
    owner = [0] * len(boxes)
    return cfg, boxes, owner


def main():
    size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    ctx = amr.Context(device=0)
    cfg, boxes, owner = single_level_setup(size)
    g = syn.fields(cfg, ng=1)
    layout = amr.DisjointBoxLayout(ctx, boxes, (0, 0, cfg.nx - 1, cfg.ny - 1), cfg.periodic, owner)
    spec = dict(head=(1, 0), rhs=(0, 0), B=(1, 0), Pi=(1, 0), zb=(1, 0), mask=(1, 0), a=(0, 0), bX=(0, 1), bY=(0, 2))
    F = {k: amr.LevelData(layout, 1, ng, cent) for k, (ng, cent) in spec.items()}
    for k in ("head", "rhs", "B", "Pi", "zb", "mask"):
        F[k].set_global(g[k], (-spec[k][0], -spec[k][0]))
    del g
    bc = amr.make_bc(cfg.bc_lo, cfg.bc_hi)
    prm = amr.make_params(A=cfg.A, omega=cfg.omega, nu=cfg.nu, cutOffbr=cfg.cutOffbr, maxOffbr=cfg.maxOffbr)
    fac = amr.VCAMRNonLinearPoissonOpFactory().define(ctx, [layout], [], cfg.dx, bc, 0.0, [F["a"]], -1.0, [F["bX"]], [F["bY"]],
                                                      prm, [F["B"]], [F["Pi"]], [F["zb"]], [F["mask"]])
    op0 = fac.AMRnewOp(0)
    op0.UpdateOperator(F["head"], None, 0, 0, False)
    res = []
    variants = [] if os.environ.get("SG_ONLY") else [(1, 0, 3), (1, 32, 3), (1, 64, 3), (4, 0, 3)]
    for extra in os.environ.get("SG_VARIANTS", "").split(";"):
        if extra:
            variants.append(tuple(int(x) for x in extra.split(",")))
    for v in variants:
        mode, rows, minb = v[:3]
        ctx.set_relax_mode(mode)
        ctx.set_tuning(0, rows)
        ctx.set_tuning(1, minb)
        op0.relax(F["head"], F["rhs"], 3)
        n = int(os.environ.get("SG_ITERS", "48"))
        ctx.event_record(0)
        op0.relax(F["head"], F["rhs"], n)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / n
        gbs = 72.0 * size * size / (ms * 1e-3) / 1e9
        res.append(dict(mode=mode, rows_per_warp=rows, ctas_per_sm=minb, ms=ms, gbs_at_72B=gbs))
        print(json.dumps(res[-1]), flush=True)
    ctx.destroy()


if __name__ == "__main__":
    main()

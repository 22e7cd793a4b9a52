#!/usr/bin/env python
"""Micro-benchmark of the finest-level residual kernel (k_apply<1>: out = rhs - L(phi), 64 B per cell) under the block-shape
knob (sg_set_tuning key 4 = 100*bx + rows).  CUDA events on the library's stream.  Usage: python tools/apply_bench.py [size]"""
import json
import os
import sys

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
    res = amr.LevelData(layout, 1, 0, 0)
    knobs = [0, 1, 3204, 6416]
    for extra in os.environ.get("SG_KNOBS", "").split(","):
        if extra:
            knobs.append(int(extra))
    for k in knobs:
        ctx.set_tuning(4, k)
        for _ in range(3):
            op0.residual(res, F["head"], F["rhs"])
        n = 10
        ctx.event_record(0)
        for _ in range(n):
            op0.residual(res, F["head"], F["rhs"])
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / n
        print(json.dumps(dict(knob=k, ms=ms, gbs_at_64B=64.0 * size * size / (ms * 1e-3) / 1e9)), flush=True)
    # restriction (k_restrict<0>: residual restricted to the next MG depth), knob 5 = coarse rows per thread
    resC = op0.createCoarser(F["rhs"])
    for k in (1, 4, 8, 16):
        ctx.set_tuning(5, k)
        for _ in range(3):
            op0.restrictResidual(resC, F["head"], None, F["rhs"], False)
        n = 10
        ctx.event_record(0)
        for _ in range(n):
            op0.restrictResidual(resC, F["head"], None, F["rhs"], False)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / n
        print(json.dumps(dict(restrict_rows=k, ms=ms, gbs_at_58B=58.0 * size * size / (ms * 1e-3) / 1e9)), flush=True)
    ctx.destroy()


if __name__ == "__main__":
    main()

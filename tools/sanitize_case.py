#!/usr/bin/env python
"""A small end-to-end exercise of every kernel family (single-level solve in all relax modes, 3-level AMR solve, Picard-body
kernels) meant to run under `compute-sanitizer --tool memcheck`: exits 0 iff the library calls succeed."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from suhmo_b200 import amr, synthetic as syn  # noqa: E402
from tests import picard  # noqa: E402
from tests.problem import AmrGpuSide, AmrOracleSide, GpuSide, OracleSide, amr_hierarchy  # noqa: E402


def main():
    ctx = amr.Context(device=0)
    for name, scale in (("C1", 2), ("C4", 1)):
        cfg = syn.config(name, scale)
        boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        orc = OracleSide(cfg, boxes)
        orc.init_bcoef()
        gpu = GpuSide(ctx, orc)
        for mode in (0, 1, 2, 3):
            ctx.set_relax_mode(mode)
            mg = amr.AMRFASMultiGrid().define(gpu.factory, 1)
            mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
            mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=2)
            mg.destroy()
        ctx.set_relax_mode(1)
        gbe = picard.GpuBackend(gpu)
        GX = picard.extra_fields(gbe, lambda f, g: f.set_global(g, (-1, -1)))
        picard.picard_step(gbe, gpu.F, GX, npicard=1, ncyc=1)
    cfg, lv = amr_hierarchy()
    orc = AmrOracleSide(cfg, lv)
    orc.average_down("head")
    orc.init_bcoef()
    gpu = AmrGpuSide(ctx, orc)
    mg = amr.AMRFASMultiGrid().define(gpu.factory, 3)
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    mg.solve(gpu.fields("head"), gpu.fields("rhs"), fixed_cycles=2)
    ctx.sync()
    print("sanitize_case: OK")


if __name__ == "__main__":
    main()

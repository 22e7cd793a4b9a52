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
    # multi-level Picard body, moulin recharge, explicit gap update, regrid transfer (round-2 kernels)
    from suhmo_b200.timestep_amr import AmrTimeStep
    from tests.amr_picard import build_device
    cfg.moulins = [(30000.0, 40000.0, 80.0, 3000.0), (70000.0, 20000.0, 40.0, 2500.0)]
    st = build_device(ctx, cfg, lv)
    ts = AmrTimeStep(st)
    ts.begin_step()
    ts.moulin_sources()
    ts.picard_iteration(fixed_cycles=2)
    ts.update_gap(1800.0)
    st.ops[1].regridTransfer(st.S[1]["work"], st.S[1]["head"], st.S[0]["head"])
    # implicit gap-height solve (tile smoother of the linear operator) on a level without periodic sides
    cfg5 = syn.config("C5", 1)
    boxes5 = syn.domain_split(cfg5.nx, cfg5.ny, cfg5.max_box_size, cfg5.block_factor)
    from tests import gapsolve as gs
    ogap = gs.OracleGap(cfg5, boxes5)
    ggap = gs.GpuGap(ctx, ogap, None)
    amr.SolveForGap_nl(ctx, [ggap.layout], [ggap.F["a"]], [ggap.F["bX"]], [ggap.F["bY"]], [], (ogap.dx, ogap.dx), [ggap.F["b"]], [ggap.F["rhs"]],
                       ogap.beta, 1.0, 0)
    ctx.sync()
    print("sanitize_case: OK")


if __name__ == "__main__":
    main()

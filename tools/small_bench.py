#!/usr/bin/env python
"""Time per FAS V-cycle on the reference's own (small, launch-bound) grid sizes, with and without CUDA-graph replay.
  python tools/small_bench.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from suhmo_b200 import amr, synthetic as syn  # noqa: E402
from tests.problem import GpuSide, OracleSide  # noqa: E402


def main():
    ctx = amr.Context(device=0)
    for name, scale in (("C1", 1), ("C1", 16), ("C2", 4), ("C3", 1), ("C4", 1), ("C5", 1), ("C5", 4)):
        cfg = syn.config(name, scale)
        boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
        orc = OracleSide(cfg, boxes)
        orc.init_bcoef()
        gpu = GpuSide(ctx, orc)
        out = {"config": name, "grid": [cfg.nx, cfg.ny]}
        for graphs in (0, 1):
            ctx.set_tuning(2, 0 if graphs else 1)
            mg = amr.AMRFASMultiGrid().define(gpu.factory, 1)
            mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
            mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=4)
            it, hist, st = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=20)
            out["graph_ms_per_cycle" if graphs else "plain_ms_per_cycle"] = st.device_ms / 20
            out["launches_per_cycle"] = st.kernel_launches / 20
            mg.destroy()
        print(json.dumps(out), flush=True)
    ctx.set_tuning(2, 0)


if __name__ == "__main__":
    main()

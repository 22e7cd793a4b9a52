#!/usr/bin/env python
"""Runs SUHMO-style time steps on one GPU from an input.hydro: reads the file (suhmo_b200.inputs), builds the level-0 problem
from the IBC's closed-form fields (suhmo_b200.synthetic), and advances it with suhmo_b200.timestep.time_step (Picard loop with
the reference's convergence test, FAS head solves, explicit or implicit gap update).  The remaining Picard-body inputs (bump
height/spacing, sliding speed, moulin source) are the deterministic fillers of timestep.extra_fields, not a spun-up state, so
keep dt short.  Usage: python examples/run_timesteps.py tests/data/input.sample.hydro --steps 3 --dt 5"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from suhmo_b200 import amr, inputs, timestep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("input")
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--dt", type=float, default=5.0)
    ap.add_argument("--scale", type=int, default=1, help="multiply the resolution of AmrHydro.num_cells")
    args = ap.parse_args()
    inp = inputs.read(args.input)
    cfg = inputs.to_config(inp, os.path.basename(args.input))
    cfg.nx *= args.scale
    cfg.ny *= args.scale
    ctx = amr.Context(device=0)
    lev, F = timestep.Level0.from_config(ctx, cfg, use_NL=inp["params"]["use_NL"], bcoeff_otf=inp["params"]["bcoeff_otf"])
    be = timestep.GpuBackend(lev, impl_diff=bool(inp["picard"]["use_ImplDiff"]))
    X = timestep.extra_fields(be, lambda f, g: f.set_global(g, (-1, -1)))
    op0 = be.op0()
    for step in range(args.steps):
        info = timestep.time_step(be, F, X, args.dt, step)
        info.update(step=step, max_head=op0.norm(F["head"], 0), max_gap=op0.norm(F["B"], 0))
        print(json.dumps(info), flush=True)
    ctx.destroy()


if __name__ == "__main__":
    main()

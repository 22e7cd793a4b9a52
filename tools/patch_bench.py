"""Times the refined-level smoother (k_gsrb_patch) of the bench workload under tuning keys 8 (cells per thread and batch), 10
(shared-memory carve-out in percent) and 11 (CTAs per SM the registers are capped for).  python tools/patch_bench.py [size]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from suhmo_b200 import amr  # noqa: E402
from tools import workload as wl  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = amr.Context(device=0)
levels = bench.gpu_tile_hierarchy(0, wl.tile_config(size), 3)
prob = wl.Problem(size, 1, "weak", levels)
gp = bench.GpuProblem(ctx, prob, 0)
cells = prob.cells()
for k8, k10, k11 in ((0, 0, 0), (0, 0, 3), (0, 50, 0), (0, 100, 0), (0, 75, 0), (2, 0, 0), (4, 0, 0), (0, 0, 0)):
    ctx.set_tuning(8, k8); ctx.set_tuning(10, k10); ctx.set_tuning(11, k11)
    out = {"batch": k8 or 1, "carveout_pct": k10, "ctas_per_sm": k11 or 4}
    for l in (1, 2):
        op = gp.ops[l]
        op.relax(gp.F[l]["head"], gp.F[l]["rhs"], 4)
        ctx.event_record(0)
        op.relax(gp.F[l]["head"], gp.F[l]["rhs"], 16)
        ctx.event_record(1)
        ms = ctx.event_elapsed_ms(0, 1) / 16
        out[f"L{l}_ms"] = round(ms, 4)
        out[f"L{l}_gbs_at_64B"] = round(64.0 * cells[l] / (ms * 1e-3) / 1e9, 1)
    print(json.dumps(out), flush=True)

"""UpdateOperator on the 8192^2 base level: the gradient / extrapolation / face kernels (default) vs the one-pass kernel (tune key 12 = 1)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from suhmo_b200 import amr  # noqa: E402
from tools import workload as wl  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
ctx = amr.Context(device=0)
prob = wl.Problem(size, 1, "weak", [bench.syn.domain_split(size, size, 64, 2)])
gp = bench.GpuProblem(ctx, prob, 0)
op = gp.ops[0]
for key in (0, 1, 0, 1):
    ctx.set_tuning(12, key)
    op.UpdateOperator(gp.F[0]["head"], None, 0, 0, False)
    ctx.event_record(0)
    for _ in range(10):
        op.UpdateOperator(gp.F[0]["head"], None, 0, 0, False)
    ctx.event_record(1)
    print(json.dumps({"tune12": key, "ms_per_UpdateOperator": ctx.event_elapsed_ms(0, 1) / 10}), flush=True)

"""The bench workload end to end at a size the oracle finishes in seconds: each arm builds its own grids (device tagging vs oracle
tagging, same Berger-Rigoutsos) and its own fields from tools/workload.py; grids, residual-norm history and head must agree bit for
bit.  This is the small-size twin of the "parity_full_size" key of the bench line."""
import numpy as np
import pytest

import bench
from tools import workload as wl

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("size", [256, 512])
def test_bench_workload_gpu_vs_oracle(gpu_ctx, size):
    cfg = wl.tile_config(size)
    glv = bench.gpu_tile_hierarchy(0, cfg, 3)
    clv = bench.cpu_tile_hierarchy(cfg, 3)
    assert len(glv) == len(clv) == 3
    for a, b in zip(glv, clv):
        assert np.array_equal(a, b)
    prob = wl.Problem(size, 1, "weak", glv)
    gp = bench.GpuProblem(gpu_ctx, prob, 0, pinned=False)
    gp.mg.setSolverParameters(4, 4, 16, 1, 100, 1e-10, 1e-4, 1e-7)
    _, ghist, _ = gp.mg.solve(gp.fields("head"), gp.fields("rhs"), fixed_cycles=3)
    cp = bench.CpuProblem(size, 3, 4, levels=clv)
    _, _, ohist = cp.solve(3, 16)
    assert np.array_equal(ghist, ohist), (ghist, ohist)
    for l in range(3):
        g, o = gp.F[l]["head"].get_global(), cp.orc.F[l]["head"].get_global()
        assert np.array_equal(np.isnan(g), np.isnan(o))
        m = ~np.isnan(o)
        assert np.array_equal(g[m], o[m]), f"level {l}"
    gp.mg.destroy()

"""GPU parity of the Picard-body field kernels and of head + gap after Picard iterations: the same
orchestration (tests/picard.py) runs on the oracle and through the C ABI; every field must agree bit for bit, and head / gap
within the north-star tolerance 1e-10 relative L2."""
import numpy as np
import pytest

from suhmo_b200 import synthetic as syn
from tests import picard
from tests.problem import GpuSide, OracleSide, fields_equal, rel_l2

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("impl_diff", [False, True])
@pytest.mark.parametrize("name,scale", [("C1", 2), ("C2", 1), ("C4", 1), ("C5", 1)])
def test_picard_iterations_head_and_gap_parity(gpu_ctx, name, scale, impl_diff):
    cfg = syn.config(name, scale)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes)
    gpu = GpuSide(gpu_ctx, orc)
    obe, gbe = picard.OracleBackend(orc, impl_diff), picard.GpuBackend(gpu, impl_diff)
    OX = picard.extra_fields(obe, lambda f, g: f.set_global(g, (-1, -1)))
    GX = picard.extra_fields(gbe, lambda f, g: f.set_global(g, (-1, -1)))
    oh = picard.picard_step(obe, orc.F, OX, npicard=2, ncyc=3)
    gh = picard.picard_step(gbe, gpu.F, GX, npicard=2, ncyc=3)
    for a, b in zip(oh, gh):
        assert np.array_equal(a, b), (a, b)
    for k in ("head", "B", "rhs", "bX", "bY"):
        d, eq = fields_equal(gpu.F[k], orc.F[k])
        assert eq, f"{k}: max abs diff {d:g}"
    assert rel_l2(gpu.F["head"].get_global(), orc.F["head"].get_global()) <= 1e-10
    assert rel_l2(gpu.F["B"].get_global(), orc.F["B"].get_global()) <= 1e-10
    for k in ("mR", "Pw", "Re", "gradH", "qgh", "qgz", "Dterm", "RHSb"):
        d, eq = fields_equal(GX[k], OX[k])
        assert eq, f"{k}: max abs diff {d:g}"
    for k in ("Bec", "mRec", "gH", "gZ", "Dc", "Reec", "Qw", "t1", "t2", "IMec"):
        for dd in range(2):
            d, eq = fields_equal(GX[k][dd], OX[k][dd])
            assert eq, f"{k}[{dd}]: max abs diff {d:g}"


@pytest.mark.parametrize("name,impl_diff", [("C2", True), ("C4", False), ("C5", True)])
def test_time_steps_with_picard_convergence(gpu_ctx, name, impl_diff):
    """suhmo_b200.timestep.time_step -- Picard loop under the reference's lagged convergence test, head solves under its stop
    logic, explicit or implicit gap update -- over four time steps: same iteration counts, same convergence measures, head and gap
    height bit for bit"""
    cfg = syn.config(name, 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes)
    gpu = GpuSide(gpu_ctx, orc)
    obe, gbe = picard.OracleBackend(orc, impl_diff), picard.GpuBackend(gpu, impl_diff)
    OX = picard.extra_fields(obe, lambda f, g: f.set_global(g, (-1, -1)))
    GX = picard.extra_fields(gbe, lambda f, g: f.set_global(g, (-1, -1)))
    for step in (0, 1, 2, 60):
        oi = picard.time_step(obe, orc.F, OX, 5.0, step, eps_picard=1e-3)
        gi = picard.time_step(gbe, gpu.F, GX, 5.0, step, eps_picard=1e-3)
        assert gi == oi, (step, gi, oi)
        for k in ("head", "B"):
            d, eq = fields_equal(gpu.F[k], orc.F[k])
            assert eq, f"step {step} {k}: max abs diff {d:g}"


def test_example_driver_runs_from_an_input_file():
    """examples/run_timesteps.py: input.hydro -> device problem -> time steps, through the package only (no oracle)"""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "examples", "run_timesteps.py"), os.path.join(root, "tests", "data", "input.sample.hydro"),
                        "--steps", "3", "--dt", "5", "--scale", "2"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-1500:] + r.stderr[-1500:]
    lines = [json.loads(l) for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 3 and all(np.isfinite(l["max_head"]) and np.isfinite(l["max_gap"]) and l["picard_iterations"] >= 1 for l in lines)
    assert lines[0]["picard_iterations"] >= 4

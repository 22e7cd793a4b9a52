"""CPU: the Picard-step orchestration runs on the oracle and produces finite, changing fields (the GPU twin is compared
against it bit for bit in tests/test_gpu_picard.py)."""
import numpy as np
import pytest

from suhmo_b200 import synthetic as syn
from tests import picard
from tests.problem import OracleSide


@pytest.mark.parametrize("impl_diff", [False, True])
@pytest.mark.parametrize("name", ["C1", "C4"])
def test_picard_step_on_oracle(name, impl_diff):
    cfg = syn.config(name, 2 if name == "C1" else 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes)
    be = picard.OracleBackend(orc, impl_diff)
    X = picard.extra_fields(be, lambda f, g: f.set_global(g, (-1, -1)))
    b0 = orc.F["B"].get_global().copy()
    h0 = orc.F["head"].get_global().copy()
    hists = picard.picard_step(be, orc.F, X)
    b1, h1 = orc.F["B"].get_global(), orc.F["head"].get_global()
    assert np.all(np.isfinite(b1)) and np.all(np.isfinite(h1))
    assert not np.array_equal(b0, b1) and not np.array_equal(h0, h1)
    assert all(np.all(np.isfinite(h)) for h in hists)
    mr = X["mR"].get_global()
    assert np.all(mr >= 0.0)


@pytest.mark.parametrize("name,impl_diff", [("C2", True), ("C4", False)])
def test_time_steps_with_picard_convergence_on_oracle(name, impl_diff):
    """suhmo_b200.timestep.time_step (Picard loop with the reference's lagged convergence test and solver stop logic) on the oracle.
    The extra fields of tests/picard.py are deterministic fillers, not a spun-up SUHMO state: a short dt keeps the gap update tame."""
    cfg = syn.config(name, 1)
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes)
    be = picard.OracleBackend(orc, impl_diff)
    X = picard.extra_fields(be, lambda f, g: f.set_global(g, (-1, -1)))
    infos = []
    for step in (0, 1, 2, 60):
        infos.append(picard.time_step(be, orc.F, X, 5.0, step, eps_picard=1e-3))
    assert infos[0]["picard_iterations"] >= 4 and infos[1]["picard_iterations"] >= 4     # "m_cur_PicardIte > 2" while step < 2
    assert all(i["x_h"][-1] < 0.05 for i in infos[:3]) and infos[3]["x_h"][-1] < 1e-3
    assert all(c >= 2 for i in infos for c in i["head_cycles"])
    assert (infos[0]["gap_cycles"] is not None) == impl_diff
    assert np.all(np.isfinite(orc.F["head"].get_global())) and np.all(np.isfinite(orc.F["B"].get_global()))

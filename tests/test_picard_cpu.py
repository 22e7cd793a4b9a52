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

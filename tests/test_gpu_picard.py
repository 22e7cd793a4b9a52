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

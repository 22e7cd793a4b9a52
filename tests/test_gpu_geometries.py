"""GPU parity on geometries the five named configurations do not exercise: x-periodic and doubly periodic domains, sizes that
are not powers of two (odd row/column counts on the coarser MG depths), strips narrower than one warp strip, a single tiny box,
anisotropic mesh spacing, non-zero boundary values on every side.  Every relax mode, residual, restriction/prolongation chain
and fixed V-cycles stay bit-exact against the oracle."""
import dataclasses

import numpy as np
import pytest

from oracle import binding as ob
from suhmo_b200 import synthetic as syn
from tests.problem import GpuSide, OracleSide, fields_equal

pytestmark = pytest.mark.gpu


def variant(base, **kw):
    return dataclasses.replace(syn.config(base, 1), **kw)


GEOMS = {
    "x-periodic": lambda: variant("C5", nx=128, ny=96, periodic=(1, 0), bc_lo=(0, 0), bc_hi=(0, 1), max_box_size=32, domain_size=(4000.0, 3000.0)),
    "doubly-periodic": lambda: variant("C5", nx=64, ny=64, periodic=(1, 1), max_box_size=16, domain_size=(2000.0, 2000.0)),
    "non-pow2": lambda: variant("C2", nx=72, ny=120, max_box_size=24, block_factor=1, domain_size=(72000.0, 120000.0)),
    "non-pow2-wide": lambda: variant("C4", nx=200, ny=56, max_box_size=8, block_factor=1, domain_size=(5000.0, 1400.0)),
    "narrow": lambda: variant("C1", nx=24, ny=160, max_box_size=8, block_factor=1, domain_size=(24.0, 160.0)),
    "one-box": lambda: variant("C1", nx=8, ny=8, max_box_size=8, block_factor=1, domain_size=(16.0, 16.0)),
    "anisotropic": lambda: variant("C3", nx=96, ny=64, max_box_size=32, domain_size=(100000.0, 20000.0)),
}


def same(g, o, what):
    d, eq = fields_equal(g, o)
    assert eq, f"{what}: max abs diff {d:g} (expected bit-exact)"


@pytest.mark.parametrize("geom", sorted(GEOMS))
def test_geometry_parity(gpu_ctx, geom):
    cfg = GEOMS[geom]()
    boxes = syn.domain_split(cfg.nx, cfg.ny, cfg.max_box_size, cfg.block_factor)
    orc = OracleSide(cfg, boxes, bc_vals=((0.3, -0.2), (0.1, 0.4)))
    orc.init_bcoef()
    gpu = GpuSide(gpu_ctx, orc)
    oop, gop = orc.op(), gpu.factory.AMRnewOp(0)
    try:
        for mode in (0, 1, 2, 3, 4, 5):
            gpu_ctx.set_relax_mode(mode)
            for n in (1, 2):
                oop.relax(orc.F["head"], orc.F["rhs"], n)
                gop.relax(gpu.F["head"], gpu.F["rhs"], n)
                same(gpu.F["head"], orc.F["head"], f"{geom}: relax x{n} mode {mode}")
    finally:
        gpu_ctx.set_relax_mode(1)
    ores, gres = ob.Field(orc.layout, 1, 0), gpu.new_like("rhs")
    oop.residual(ores, orc.F["head"], orc.F["rhs"])
    gop.residual(gres, gpu.F["head"], gpu.F["rhs"])
    same(gres, ores, f"{geom}: residual")
    oop.update_operator(orc.F["head"])
    for key12, flow in ((0, "gradient / exchange / extrapolation / face kernels"), (1, "one pass")):   # tune key 12 = 1: k_update_op_fused
        gpu_ctx.set_tuning(12, key12)
        try:
            for f in (gpu.F["bX"], gpu.F["bY"]):
                gop.setToZero(f)
            gop.UpdateOperator(gpu.F["head"], None, 0, 0, False)
        finally:
            gpu_ctx.set_tuning(12, 0)
        same(gpu.F["bX"], orc.F["bX"], f"{geom}: bX after UpdateOperator ({flow})")
        same(gpu.F["bY"], orc.F["bY"], f"{geom}: bY after UpdateOperator ({flow})")
    osolver = orc.solver()
    mg = gpu.amr.AMRFASMultiGrid().define(gpu.factory, 1)
    assert mg.depth == osolver.depth
    mg.setSolverParameters(4, 4, 10, 1, 100, 1e-10, 1e-4, 1e-7)
    for mode in (1, 0, 3, 5):
        gpu_ctx.set_relax_mode(mode)
        it, ohist = osolver.solve(orc.F["head"], orc.F["rhs"], ob.make_solver_params(bottom=10, fixed_cycles=3))
        git, ghist, st = mg.solve([gpu.F["head"]], [gpu.F["rhs"]], fixed_cycles=3)
        assert np.array_equal(ghist, ohist), (geom, mode, ghist, ohist)
        same(gpu.F["head"], orc.F["head"], f"{geom}: head after 3 V-cycles, mode {mode}")
    gpu_ctx.set_relax_mode(1)

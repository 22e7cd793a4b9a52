"""Properties of the restated inter-level transfers and of the moulin recharge (oracle side, no GPU): PiecewiseLinearFillPatch and
FineInterp reproduce linear fields and keep bounds, FineInterp conserves the coarse mean, the moulin source integrates to the
prescribed fluxes over the composite grid; the multi-level Picard restatement runs and couples the levels."""
import ctypes as C

import numpy as np

from oracle import binding as ob
from tests.amr_picard import build_oracle
from tests.problem import amr_hierarchy


def _two_levels():
    cfg, lv = amr_hierarchy("C5")
    lc = ob.Layout(lv[0], (0, 0, 63, 63), (0, 0))
    lf = ob.Layout(lv[1], (0, 0, 127, 127), (0, 0))
    return lc, lf


def _lin(ny, nx, r, a=0.3, b=-0.7, c=2.0):
    jj, ii = np.meshgrid(np.arange(ny) - 1, np.arange(nx) - 1, indexing="ij")
    return c + a * (ii + 0.5) / r + b * (jj + 0.5) / r


def test_pwl_fill_patch_linear_and_bounded():
    lc, lf = _two_levels()
    crse, fine = ob.Field(lc, 1, 1), ob.Field(lf, 1, 1)
    crse.set_global(_lin(66, 66, 1), (-1, -1))
    fine.setval(-99.0)
    ob.lib().orc_pwl_fill_patch(fine.h, crse.h, 2)
    exact = _lin(130, 130, 2)
    n = 0
    for b in range(len(lf.boxes)):
        fab = fine.fab(b)[0][0]
        x0, y0, x1, y1 = lf.boxes[b]
        ref = exact[y0:y1 + 3, x0:x1 + 3]
        filled = fab != -99.0
        filled[1:-1, 1:-1] = False
        n += int(filled.sum())
        assert np.allclose(fab[filled], ref[filled], rtol=0, atol=1e-13)
    assert n > 100
    # a step function must not overshoot (van Leer limiter)
    g = np.where(np.arange(66)[None, :] > 20, 1.0, 0.0) * np.ones((66, 1))
    crse.set_global(g, (-1, -1))
    fine.setval(0.5)
    ob.lib().orc_pwl_fill_patch(fine.h, crse.h, 2)
    for b in range(len(lf.boxes)):
        fab = fine.fab(b)[0][0]
        assert fab.min() >= 0.0 and fab.max() <= 1.0


def test_fine_interp_linear_conservative_and_regrid_transfer():
    lc, lf = _two_levels()
    crse, fine = ob.Field(lc, 1, 1), ob.Field(lf, 1, 1)
    crse.set_global(_lin(66, 66, 1), (-1, -1))
    ob.lib().orc_fine_interp(fine.h, crse.h, 2)
    g = fine.get_global()
    exact = _lin(130, 130, 2)[1:-1, 1:-1]
    m = ~np.isnan(g)
    assert m.any() and np.allclose(g[m], exact[m], rtol=0, atol=1e-13)
    rng = np.random.RandomState(5)
    crse.set_global(rng.rand(66, 66), (-1, -1))
    ob.lib().orc_fine_interp(fine.h, crse.h, 2)
    g, c = fine.get_global(), crse.get_global()
    for bx in lf.boxes:
        blk = g[bx[1]:bx[3] + 1, bx[0]:bx[2] + 1]
        mean = blk.reshape(blk.shape[0] // 2, 2, blk.shape[1] // 2, 2).mean(axis=(1, 3))
        assert np.allclose(mean, c[bx[1] // 2:bx[3] // 2 + 1, bx[0] // 2:bx[2] // 2 + 1], rtol=0, atol=1e-14)
        if bx[0] > 0 and bx[1] > 0 and bx[2] < 127 and bx[3] < 127:        # limited (the boundary-normal one-sided slope is not: type 3)
            assert blk.min() >= 0.0 and blk.max() <= 1.0
    # destructiveRegrid: old data wins where it exists
    old = ob.Field(ob.Layout(lf.boxes[:2], (0, 0, 127, 127), (0, 0)), 1, 1)
    old.setval(7.0)
    ob.lib().orc_regrid_transfer(fine.h, old.h, crse.h, 2)
    g = fine.get_global()
    for k, bx in enumerate(lf.boxes):
        blk = g[bx[1]:bx[3] + 1, bx[0]:bx[2] + 1]
        assert (blk == 7.0).all() if k < 2 else (blk != 7.0).all()


def test_moulin_source_integrates_to_the_fluxes_and_picard_runs():
    from oracle import picard_amr as opa
    cfg, lv = amr_hierarchy("C5")
    cfg.moulins = [(30000.0, 40000.0, 80.0, 3000.0), (70000.0, 20000.0, 40.0, 2500.0)]
    H = build_oracle(cfg, lv)
    ts = opa.TimeStep(H)
    ts.begin_step()
    integ = ts.moulin_sources()
    assert (integ > 0).all()
    total = 0.0
    for l in range(H.nlev):
        d = ob.Field(H.layouts[l], 1, 1)
        d.copy_from(H.S[l]["MS"])
        if l + 1 < H.nlev:
            ob.zero_covered(d, H.layouts[l + 1])
        g = d.get_global()
        total += np.nansum(g) * H.dx[l][0] * H.dx[l][1]
    assert total == np.float64(total) and abs(total - 120.0) < 1e-6 * 120.0
    h0 = [S["head"].get_global().copy() for S in H.S]
    hist = ts.picard_iteration(ob.make_solver_params(bottom=10, fixed_cycles=2))
    assert len(hist) == 3 and np.isfinite(hist).all()
    ts.update_gap(3600.0)
    for l in range(H.nlev):
        g = H.S[l]["head"].get_global()
        assert np.isfinite(g[~np.isnan(g)]).all() and not np.array_equal(np.nan_to_num(g), np.nan_to_num(h0[l]))
        b = H.S[l]["B"].get_global()
        assert np.isfinite(b[~np.isnan(b)]).all() and (b[~np.isnan(b)] > 0).all()
    assert 0.0 <= ts.picard_change() < 1.0

"""AMR hierarchy generation.  Grid generation (Berger-Rigoutsos, host-only code of the library) is checked
by the properties BRMeshRefine guarantees -- the fork's tie-breaking rules cannot be recovered from SUHMO, so box-for-box
parity with the reference is UNPINNED: every tag covered, boxes disjoint, inside the domain, block-factor aligned, no longer
than max_box_size, properly nested with the buffer, and clustering efficiency not below the fill ratio where it can be met."""
import numpy as np
import pytest

from suhmo_b200 import amr, synthetic as syn


def covered(boxes, shape):
    m = np.zeros(shape, dtype=np.int32)
    for b in boxes:
        m[b[1]:b[3] + 1, b[0]:b[2] + 1] += 1
    return m


def blob_tags(n, centres, rad):
    jj, ii = np.meshgrid(np.arange(n), np.arange(n), indexing="ij")
    t = np.zeros((n, n), dtype=np.uint8)
    for (cx, cy) in centres:
        t |= (((ii - cx) ** 2 + (jj - cy) ** 2) <= rad * rad).astype(np.uint8)
    return t


@pytest.mark.parametrize("bf,maxbox,fill", [(2, 64, 0.5), (8, 32, 0.7), (4, 16, 0.9)])
def test_two_level_regrid_properties(bf, maxbox, fill):
    n = 128
    base = syn.domain_split(n, n, 64, 2)
    t0 = blob_tags(n, [(30, 40), (90, 85), (100, 20)], 9)
    mr = amr.BRMeshRefine((0, 0, n - 1, n - 1), fill, bf, 2, maxbox)
    levels = mr.regrid(base, [t0])
    assert len(levels) == 2
    b1 = levels[1]
    cov = covered(b1, (2 * n, 2 * n))
    assert cov.max() == 1                                         # disjoint
    fine_tags = np.repeat(np.repeat(t0, 2, axis=0), 2, axis=1)
    assert np.all(cov[fine_tags > 0] == 1)                        # every tag is refined
    assert np.all(b1[:, :2] % bf == 0) and np.all((b1[:, 2:] + 1) % bf == 0)      # block-factor aligned
    assert np.all(b1[:, 2] - b1[:, 0] + 1 <= maxbox) and np.all(b1[:, 3] - b1[:, 1] + 1 <= maxbox)
    assert b1.min() >= 0 and b1[:, 2].max() < 2 * n and b1[:, 3].max() < 2 * n
    # overall efficiency: tagged fine cells / refined cells (alignment can only lower it by the coarsening factor squared)
    eff = fine_tags.sum() / cov.sum()
    assert eff >= fill / max(1, bf // 2) ** 2 * 0.5, eff
    # sorted like a DisjointBoxLayout
    key = b1[:, 1].astype(np.int64) * 100000 + b1[:, 0]
    assert np.all(np.diff(key) > 0)


def test_three_level_regrid_is_properly_nested():
    n, nr = 64, 2
    base = syn.domain_split(n, n, 32, 2)
    t0 = blob_tags(n, [(20, 24), (45, 40)], 6)
    t1 = blob_tags(2 * n, [(40, 48), (92, 84)], 5)       # tags on level 1 (inside the level-1 grids the level-0 tags produce)
    mr = amr.BRMeshRefine((0, 0, n - 1, n - 1), 0.6, 2, nr, 32)
    levels = mr.regrid(base, [t0, t1])
    assert len(levels) == 3
    c1, c2 = covered(levels[1], (2 * n, 2 * n)), covered(levels[2], (4 * n, 4 * n))
    assert c1.max() == 1 and c2.max() == 1
    assert np.all(c2[np.repeat(np.repeat(t1, 2, axis=0), 2, axis=1) > 0] == 1)
    # proper nesting: every level-2 box, coarsened to level 1 and grown by the buffer (clipped to the domain), lies in level 1
    for b in levels[2]:
        x0, y0, x1, y1 = b[0] // 2 - nr, b[1] // 2 - nr, b[2] // 2 + nr, b[3] // 2 + nr
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, 2 * n - 1), min(y1, 2 * n - 1)
        assert np.all(c1[y0:y1 + 1, x0:x1 + 1] == 1), b
    # determinism
    again = mr.regrid(base, [t0, t1])
    assert all(np.array_equal(a, b) for a, b in zip(levels, again))


def test_no_tags_no_levels_and_full_tags_full_level():
    n = 32
    base = syn.domain_split(n, n, 16, 2)
    mr = amr.BRMeshRefine((0, 0, n - 1, n - 1), 0.5, 2, 1, 16)
    assert len(mr.regrid(base, [np.zeros((n, n), dtype=np.uint8)])) == 1
    lv = mr.regrid(base, [np.ones((n, n), dtype=np.uint8)])
    assert covered(lv[1], (2 * n, 2 * n)).min() == 1 and len(lv[1]) == 16


@pytest.mark.parametrize("seed", [1, 2, 3, 4, 5, 6])
def test_box_for_box_against_the_oracle_restatement(seed):
    """sg_br_regrid (C++, the product) against oracle/br_regrid.py (numpy, written separately from the same rule set):
    identical box lists, level by level, on random blob / noise tag maps, two and three levels, several parameter sets."""
    from oracle import br_regrid as obr
    rng = np.random.RandomState(seed)
    n = [64, 96, 128][seed % 3]
    nr = [1, 2, 4][seed % 3]
    bf = [2, 4, 8][(seed // 2) % 3]
    maxbox = [16, 32, 64][seed % 3]
    fill = [0.5, 0.7, 0.85][(seed // 3) % 3]
    base = syn.domain_split(n, n, 32, 2)
    cs = [(int(x), int(y)) for x, y in rng.randint(8, n - 8, size=(4, 2))]
    t0 = blob_tags(n, cs, int(rng.randint(3, 9)))
    t0 |= (rng.rand(n, n) < 0.002).astype(np.uint8)           # isolated tags
    mr = amr.BRMeshRefine((0, 0, n - 1, n - 1), fill, bf, nr, maxbox)
    got = mr.regrid(base, [t0])
    exp = obr.regrid((0, 0, n - 1, n - 1), base, [t0], fill, bf, nr, maxbox)
    assert len(got) == len(exp) == 2
    assert np.array_equal(got[1], exp[1]), (got[1][:5], exp[1][:5])
    # third level: tags inside the level-1 grids
    c1 = covered(got[1], (2 * n, 2 * n)) > 0
    t1 = (blob_tags(2 * n, [(2 * c[0], 2 * c[1]) for c in cs[:2]], 4) > 0) & c1
    got = mr.regrid(base, [t0, t1.astype(np.uint8)])
    exp = obr.regrid((0, 0, n - 1, n - 1), base, [t0, t1.astype(np.uint8)], fill, bf, nr, maxbox)
    assert len(got) == len(exp)
    for l in range(1, len(got)):
        assert np.array_equal(got[l], exp[l]), f"level {l}"

"""The bench workload module (tools/workload.py): grids, owners, fields -- host logic, no GPU."""
import numpy as np

from tools import workload as wl


def _levels():
    base = np.array([(x, y, x + 63, y + 63) for y in range(0, 256, 64) for x in range(0, 256, 64)], dtype=np.int32)
    # three clusters on level 1 (one L-shaped of 3 boxes, two single boxes), level 2 nested in two of them
    l1 = np.array([(32, 32, 95, 95), (96, 32, 159, 95), (32, 96, 95, 127), (320, 64, 383, 127), (64, 384, 127, 447)], dtype=np.int32)
    l2 = np.array([(96, 96, 159, 159), (160, 96, 223, 159), (672, 160, 735, 223)], dtype=np.int32)
    return [base, l1, l2]


def test_components_groups_touching_boxes():
    lab = wl.components(_levels()[1])
    assert lab[0] == lab[1] == lab[2]
    assert len({lab[0], lab[3], lab[4]}) == 3
    # corner contact counts (the fused smoother's ghost records look at diagonal neighbours' owners too)
    lab = wl.components(np.array([(0, 0, 7, 7), (8, 8, 15, 15), (32, 0, 39, 7)]))
    assert lab[0] == lab[1] != lab[2]


def test_cluster_balance_keeps_clusters_and_nesting_on_one_rank():
    lv = _levels()
    for n in (2, 4):
        own = wl.cluster_balance(lv, n, 256 // n)
        assert [len(o) for o in own] == [len(b) for b in lv]
        assert own[0].tolist() == [(b[1] // (256 // n)) for b in lv[0]]
        assert own[1][0] == own[1][1] == own[1][2]           # the L-shaped cluster stays together
        assert own[2][0] == own[2][1] == own[1][0]           # and so do the finer boxes nested in it
        assert own[2][2] == own[1][3]
        assert all(0 <= r < n for o in own for r in o)
    own = wl.cluster_balance(lv, 2, 128)
    loads = [sum(int((b[2] - b[0] + 1) * (b[3] - b[1] + 1)) for l in (1, 2) for b, r in zip(lv[l], own[l]) if r == k) for k in range(2)]
    assert min(loads) > 0


def test_weak_problem_is_the_tile_stacked():
    lv = _levels()
    p = wl.Problem(256, 4, "weak", lv)
    assert p.cfg.ny == 1024 and p.cfg.nx == 256 and p.cfg.dx == p.tile_cfg.dx
    for l in range(3):
        assert len(p.levels[l]) == 4 * len(lv[l])
        for k in range(4):
            ids = p.owned(l, k)
            assert np.array_equal(p.to_tile(l, p.levels[l][ids]), lv[l])
    # every tile's fields are the same bits
    a, b = p.level_fabs(1, 0), p.level_fabs(1, 3)
    n1 = len(lv[1])
    for k in a:
        for i in range(n1):
            assert np.array_equal(a[k][i], b[k][3 * n1 + i])
    d = p.describe()
    assert d["cells"][0] == 256 * 1024 and d["cells_per_rank"][0] == [256 * 256] * 4


def test_strong_problem_partitions_every_box_once():
    p = wl.Problem(256, 2, "strong", _levels())
    for l in range(3):
        got = np.sort(np.concatenate([p.owned(l, r) for r in range(2)]))
        assert np.array_equal(got, np.arange(len(p.levels[l])))


def test_cpu_hierarchy_small_tile():
    """the oracle-tagged hierarchy of a 256^2 tile: nested, aligned, non-trivial"""
    import bench
    cfg = wl.tile_config(256)
    lv = bench.cpu_tile_hierarchy(cfg, 3)
    assert len(lv) == 3 and len(lv[1]) > 0 and len(lv[2]) > 0
    for l in (1, 2):
        b = lv[l]
        assert (b[:, :2] % 2 == 0).all() and ((b[:, 2:] + 1) % 2 == 0).all()
        assert ((b[:, 2] - b[:, 0] + 1) <= 64).all() and ((b[:, 3] - b[:, 1] + 1) <= 64).all()
    # level 2 lies inside level 1 (coarsened)
    cover = np.zeros((512, 512), dtype=bool)
    for b in lv[1]:
        cover[b[1]:b[3] + 1, b[0]:b[2] + 1] = True
    for b in lv[2]:
        assert cover[b[1] // 2:b[3] // 2 + 1, b[0] // 2:b[2] // 2 + 1].all()

"""Host-only logic of the multi-GPU particle decomposition: the work-balanced layer cuts every rank derives
from the full particle set (p3m_balanced_cuts; csrc/dist.cu:dist_balance_cuts).  Needs no GPU."""
import numpy as np

from particlesimulation_b200 import capi, ics


def p3m_params(grid):
    f32 = np.float32
    p = capi.default_params()
    p.nx = p.ny = p.nz = grid
    p.box[:] = (60.0, 60.0, 60.0)
    p.H = f32(60.0 / (grid // 2))
    p.DT, p.G = 1.0, 4.5e-3
    p.particle_diameter = f32(3 * float(p.H))
    p.p3m = 1
    p.cutoff_radius = f32(0.7 * float(p.particle_diameter))
    p.softening = 0.5
    return p


def layer_weights(prm, pos, layers):
    """numpy restatement: n_cell * (300 + particles in the 27-cell neighbourhood), summed per z layer."""
    m = int(60.0 / float(prm.cutoff_radius))
    hc = 60.0 / m
    idx = np.clip(np.floor(pos.astype(np.float64) / hc).astype(np.int64), 0, m - 1)
    cnt = np.zeros((m, m, m), np.float64)  # [z, y, x]
    np.add.at(cnt, (idx[:, 2], idx[:, 1], idx[:, 0]), 1.0)
    s = cnt.copy()
    for ax in range(3):
        lo = np.roll(s, 1, axis=ax)
        hi = np.roll(s, -1, axis=ax)
        sl = [slice(None)] * 3
        sl[ax] = 0
        lo[tuple(sl)] = 0
        sl[ax] = -1
        hi[tuple(sl)] = 0
        s = s + lo + hi
    assert layers == m
    return (cnt * (300.0 + s)).sum(axis=(1, 2))


def test_balanced_cuts_p3m_even_out_the_pair_work():
    prm = p3m_params(128)
    pos, _, _ = ics.clustered_disk_halo(200000, seed=3)
    for nranks in (2, 4, 8):
        cuts, layers = capi.balanced_cuts(prm, nranks, pos)
        geo, layers2 = capi.slab_cuts(prm, nranks)
        assert layers == layers2 and cuts[0] == 0 and cuts[-1] == layers and np.all(np.diff(cuts) > 0)
        again, _ = capi.balanced_cuts(prm, nranks, pos.copy())
        assert np.array_equal(cuts, again), "every rank must derive the same cuts from the same set"
        w = layer_weights(prm, pos, layers)
        per_rank = lambda c: np.array([w[c[r]:c[r + 1]].sum() for r in range(nranks)])
        bal, eq = per_rank(cuts), per_rank(geo)
        assert bal.max() <= eq.max() * 1.0001, (nranks, bal, eq)
        # a single layer is never split: the heaviest rank carries at most its fair share + one layer
        assert bal.max() <= w.sum() / nranks + w.max() * 1.0001, (nranks, bal)


def test_balanced_cuts_pm_are_count_quantiles():
    p = capi.default_params()
    p.nx = p.ny = p.nz = 128
    p.box[:] = (60.0, 60.0, 60.0)
    p.H = 60.0 / 64
    p.p3m = 0
    rng = np.random.default_rng(1)
    pos = np.stack([rng.uniform(2, 58, 300000), rng.uniform(2, 58, 300000),
                    np.clip(30 + 6 * rng.standard_normal(300000), 2, 58)], axis=1).astype(np.float32)
    cuts, layers = capi.balanced_cuts(p, 4, pos)
    geo, _ = capi.slab_cuts(p, 4)
    lay = np.clip(np.floor(pos[:, 2] / np.float32(p.H) / 8).astype(int), 0, layers - 1)
    count = lambda c: np.array([((lay >= c[r]) & (lay < c[r + 1])).sum() for r in range(4)])
    assert count(cuts).sum() == len(pos)
    assert count(cuts).max() < count(geo).max()  # z-clustered set: equal layer counts are badly unbalanced
    assert np.all(np.diff(cuts) > 0) and cuts[-1] == layers

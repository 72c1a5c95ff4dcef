"""CPU models of two pieces of round-2 host / index logic (no GPU): the rank arithmetic of the incremental re-sort
(particlesimulation_b200/csrc/incsort.cu) and the plane ownership of the split FFT slabs
(particlesimulation_b200/csrc/dist_mesh.cu, SlabRuns).  The kernels themselves are checked on the GPU
(tests/test_gpu_parity.py::test_incremental_sort_equals_full_sort, tests/test_multi_gpu.py); these restatements pin
the formulas the kernels implement."""
import numpy as np
import pytest


# ---------------------------------------------------------------------------------------------- incremental sort
def merge_movers(old_keys, new_keys):
    """Destination of every element under the incsort.cu rules.  old_keys is sorted (the keys the array was sorted
    by); an element is a mover when its new key differs.  Stayer i -> (i - movers before i) + #sorted movers that
    order before (key, i); sorted mover j -> j + #stayers before the first old index i* with
    (old_keys[i*], i*) >= (key, index)."""
    n = len(old_keys)
    idx = np.arange(n)
    mover = new_keys != old_keys
    movers_before = np.concatenate(([0], np.cumsum(mover)))[:-1]            # exclusive prefix (moff + mask word)
    mk, mi = new_keys[mover], idx[mover]
    order = np.lexsort((mi, mk))                                            # stable radix sort of (key, index) pairs
    mk, mi = mk[order], mi[order]
    dest = np.empty(n, np.int64)
    # stayers: binary search in the sorted movers, ties by old index
    comp_m = mk.astype(np.int64) * (n + 1) + mi
    st = idx[~mover]
    comp_s = new_keys[st].astype(np.int64) * (n + 1) + st
    dest[st] = (st - movers_before[st]) + np.searchsorted(comp_m, comp_s, side="left")
    # movers: binary search in the OLD key array (sorted over all old indices), ties by old index
    comp_old = old_keys.astype(np.int64) * (n + 1) + idx
    istar = np.searchsorted(comp_old, comp_m, side="left")
    stayers_before = istar - np.concatenate((movers_before, [mover.sum()]))[istar]
    dest[mi] = np.arange(len(mi)) + stayers_before
    return dest


@pytest.mark.parametrize("n,nkeys,frac", [(1, 3, 1.0), (1000, 7, 0.0), (1000, 7, 0.1), (5000, 50, 0.5), (4096, 4096, 0.05),
                                          (3000, 5, 1.0)])
def test_mover_merge_equals_a_stable_sort(n, nkeys, frac):
    rng = np.random.default_rng(n + nkeys)
    old = np.sort(rng.integers(0, nkeys, n)).astype(np.uint32)
    new = old.copy()
    move = rng.random(n) < frac
    new[move] = rng.integers(0, nkeys, int(move.sum()))                    # may by chance equal the old key: a stayer
    dest = merge_movers(old, new)
    assert np.array_equal(np.sort(dest), np.arange(n))                     # a permutation
    out = np.empty(n, np.int64)
    out[dest] = np.arange(n)
    assert np.array_equal(out, np.argsort(new, kind="stable"))             # element for element the stable sort


def test_migration_arrivals_are_movers_and_keep_the_old_keys_sorted():
    """Arrivals of a migration are appended behind the stayers with old key 0xffffffff (k_migrate_skeys)."""
    rng = np.random.default_rng(3)
    stay = np.sort(rng.integers(0, 100, 500)).astype(np.uint32)
    old = np.concatenate((stay, np.full(60, 0xFFFFFFFF, np.uint32)))
    new = np.concatenate((stay, rng.integers(0, 100, 60).astype(np.uint32)))
    assert np.all(np.diff(old.astype(np.int64)) >= 0)
    dest = merge_movers(old, new)
    out = np.empty(len(old), np.int64)
    out[dest] = np.arange(len(old))
    assert np.array_equal(out, np.argsort(new, kind="stable"))


@pytest.mark.parametrize("keybits", [6, 15, 18, 27, 30, 32])
def test_radix_digit_plan_covers_the_key(keybits):
    passes = (keybits + 9) // 10                                           # sort_incremental()
    rb = (keybits + passes - 1) // passes
    assert 1 <= rb <= 10 and passes * rb >= keybits and (passes - 1) * rb < keybits


# ---------------------------------------------------------------------------------------------- split FFT slabs
class SlabRuns:
    def __init__(self, nz, nranks, split):
        self.nz, self.nzl = nz, nz // nranks
        self.nruns, self.len = (2, self.nzl // 2) if split else (1, self.nzl)

    def first(self, rank, run):
        if self.nruns == 1:
            return rank * self.nzl
        return rank * self.len if run == 0 else self.nz // 2 + rank * self.len

    def local_of(self, rank, run, z):
        return run * self.len + z - self.first(rank, run)


@pytest.mark.parametrize("nz,nranks", [(32, 2), (32, 4), (512, 8), (1024, 8), (256, 2)])
@pytest.mark.parametrize("split", [True, False])
def test_every_plane_has_one_owner_and_one_local_slot(nz, nranks, split):
    sr = SlabRuns(nz, nranks, split)
    owner = np.full(nz, -1)
    local = np.full(nz, -1)
    for r in range(nranks):
        for run in range(sr.nruns):
            z0 = sr.first(r, run)
            assert np.all(owner[z0:z0 + sr.len] == -1)
            owner[z0:z0 + sr.len] = r
            local[z0:z0 + sr.len] = [sr.local_of(r, run, z) for z in range(z0, z0 + sr.len)]
    assert np.all(owner >= 0)
    for r in range(nranks):
        assert sorted(local[owner == r]) == list(range(sr.nzl))           # the rank's slab is filled exactly once
    if split:                                                              # run 0 in the occupied half, run 1 in the padding
        assert np.all(owner[:nz // 2] == np.repeat(np.arange(nranks), sr.len))
        assert np.all(owner[nz // 2:] == np.repeat(np.arange(nranks), sr.len))


@pytest.mark.parametrize("nz,nranks", [(64, 2), (256, 4), (512, 8)])
def test_density_planes_reach_their_owners_exactly_once_and_travel_less_when_split(nz, nranks):
    """slab_reduce_density: the planes [z0, z0 + nzp) a particle rank deposited into are cut by the owners' runs."""
    rng = np.random.default_rng(nz)
    sent = {}
    for split in (True, False):
        sr = SlabRuns(nz, nranks, split)
        total = 0
        for src in range(nranks):
            # particle slab of `src`: an equal share of the occupied half, plus the stencil halo
            z0 = max(src * (nz // 2) // nranks - 2, 0)
            z1 = min((src + 1) * (nz // 2) // nranks + 3, nz)
            covered = np.zeros(nz, int)
            for dst in range(nranks):
                for run in range(sr.nruns):
                    lo, hi = max(z0, sr.first(dst, run)), min(z1, sr.first(dst, run) + sr.len)
                    if hi > lo:
                        covered[lo:hi] += 1
                        if dst != src:
                            total += hi - lo
            assert np.all(covered[z0:z1] == 1) and covered.sum() == z1 - z0
        sent[split] = total
    assert sent[True] < sent[False]                                        # that is the point of the split

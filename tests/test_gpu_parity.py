"""GPU parity tests: the CUDA path (through the C ABI of include/p3m_b200.h) against the CPU oracle
on the same seeded inputs.  Bars (BASELINE.json north_star):
  * cell assignment and sort order: bit-exact;
  * density, potential, accelerations: rel-L2 <= 1e-4 (fp32 path vs the fp32 oracle == reference),
    <= 1e-6 (fp64 path vs the fp64 restatement);
  * energy / momentum drift over 100 leapfrog steps matching the oracle's run.
"""
import numpy as np
import pytest

import refapi
from common import (degenerate_mask, disk_case, morton3, plummer_case, project_out_degenerate, to_p3m,
                    uniform_case)
from particlesimulation_b200 import capi
from refapi import Oracle, rel_l2

pytestmark = pytest.mark.gpu

TOL32 = 1e-4  # north_star: fp32 path
TOL64 = 1e-6  # north_star: fp64 path

_orc = {}


def oracle(prec):
    if prec not in _orc:
        _orc[prec] = Oracle(prec)
    return _orc[prec]


def code_units(p, pos, vel, mass, prec="f32"):
    return oracle(prec).to_code_units(p, pos, vel, mass)


# --------------------------------------------------------------------------------------------- A0
@pytest.mark.parametrize("p3m", [True, False])
@pytest.mark.parametrize("n", [1, 37, 5000])
def test_cells_and_sort_order_bit_exact(n, p3m):
    p, pos, vel, mass = plummer_case(n)
    pc, _, _ = code_units(p, pos, vel, mass)
    with capi.Context(to_p3m(p, p3m=p3m)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.bin_sort()
        mc, cc, order = ctx.cells()
        gpos, _, _ = ctx.get_particles(capi.UNITS_CODE, want=("pos",))
        gdims = ctx.chaining_dims() if p3m else None
        sbits = ctx.binning()["sbits"]
    assert np.array_equal(gpos, pc), "code-unit positions must be bit-identical to the reference's"
    # PM mesh cell: (int)pos, flat x + y*Nx + z*Nx*Ny  (source/pmMethod.cpp:250-252)
    t = pc.astype(np.int32)  # truncation, positions are positive
    assert np.array_equal(mc, t[:, 0] + t[:, 1] * p.nx + t[:, 2] * p.nx * p.ny)
    if p3m:
        dims, cell = oracle("f32").chaining_cells(p, pc)
        assert np.array_equal(gdims, dims)
        assert np.array_equal(cc, cell), "chaining-mesh cell ids must be bit-exact"
        cx, cy, cz = cell % dims[0], (cell // dims[0]) % dims[1], cell // (dims[0] * dims[1])
    else:
        assert np.all(cc == -1)
        cx, cy, cz = t[:, 0] >> 3, t[:, 1] >> 3, t[:, 2] >> 3
    key = morton3(cx, cy, cz)
    if sbits and not p3m:
        # PM: mesh cell inside the 8^3 tile, x fastest
        sub = ((t[:, 2] & 7).astype(np.uint64) << np.uint64(6)) | ((t[:, 1] & 7).astype(np.uint64) << np.uint64(3)) \
            | (t[:, 0] & 7).astype(np.uint64)
        key = (key << np.uint64(9)) | sub
    elif sbits:
        # sub-cell inside the chaining cell, fp32 arithmetic as on the device
        f32 = np.float32
        hc = [f32(f32(f32(p.box[d]) / f32(dims[d])) / f32(p.H)) for d in range(3)]
        S = 1 << sbits
        sub = [np.clip(((pc[:, d] / hc[d] - c.astype(f32)) * f32(S)).astype(np.int32), 0, S - 1)
               for d, c in enumerate((cx, cy, cz))]
        key = (key << np.uint64(3 * sbits)) | morton3(*sub)
    expect = np.lexsort((np.arange(n), key))  # stable: ties broken by particle id
    assert np.array_equal(order, expect.astype(np.int32)), "sort order must be (z-order cell, sub-cell, id)"


def test_sort_is_idempotent_and_deterministic():
    p, pos, vel, mass = plummer_case(4096)
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.bin_sort()
        o1 = ctx.cells()[2]
        ctx.bin_sort()
        o2 = ctx.cells()[2]
        g1 = ctx.get_particles(capi.UNITS_ORIGINAL, want=("pos",))[0]
    assert np.array_equal(o1, o2)
    assert np.allclose(g1, pos, rtol=2e-7, atol=0)  # H * (pos / H): one rounding pair


# --------------------------------------------------------------------------------------------- A1
@pytest.mark.parametrize("is_", [refapi.NGP, refapi.CIC, refapi.TSC])
@pytest.mark.parametrize("case", ["plummer", "uniform", "disk"])
def test_deposit_fp32(case, is_):
    mk = dict(plummer=plummer_case, uniform=uniform_case, disk=disk_case)[case]
    p, pos, vel, mass = mk(20000, is_=is_)
    pc, _, mcode = code_units(p, pos, vel, mass)
    rho_ref = oracle("f32").deposit(p, pc, mcode)
    rho64 = oracle("f64").deposit(p, pc.astype(np.float64), mcode.astype(np.float64))
    with capi.Context(to_p3m(p, p3m=(case != "uniform"))) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.bin_sort()
        ctx.deposit()
        rho = ctx.density()
    # the reference adds particle by particle in fp32 (biased rounding in dense cells: up to ~1e-4 for
    # NGP with equal masses); the bar is the north-star tolerance against it ...
    assert rel_l2(rho, rho_ref) < TOL32
    # ... and no further from the exact (fp64) assignment than the reference itself is
    assert rel_l2(rho, rho64) < 2 * rel_l2(rho_ref, rho64) + 1e-7
    assert abs(rho.sum(dtype=np.float64) - mcode.sum(dtype=np.float64)) < 1e-5 * mcode.sum(dtype=np.float64)


def test_deposit_fp64():
    p, pos, vel, mass = plummer_case(20000)
    o = oracle("f64")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    rho_ref = o.deposit(p, pc, mcode)
    with capi.Context(to_p3m(p, p3m=True, precision=capi.F64)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.bin_sort()
        ctx.deposit()
        rho = ctx.density(f64=True)
    assert rel_l2(rho, rho_ref) < 1e-13


def test_deposit_edge_particles():
    """Particles on cell faces, at the origin corner (x-1 aliasing) and at the box face."""
    p = refapi.make_params(8, (16, 16, 16), (8.0, 8.0, 8.0), H=1.0)
    pos = np.array([[0.0, 3.5, 3.5], [1.0, 1.0, 1.0], [7.999, 7.999, 7.999], [4.0, 4.0, 4.0],
                    [0.25, 0.25, 1.5], [3.0, 0.0, 2.0], [5.5, 5.5, 0.999], [2.0, 6.0, 6.0]], np.float32)
    mass = np.linspace(1, 2, 8).astype(np.float32)
    pc, _, mcode = code_units(p, pos, None, mass)
    rho_ref = oracle("f32").deposit(p, pc, mcode)
    for p3m in (False, True):
        with capi.Context(to_p3m(p, p3m=p3m)) as ctx:
            ctx.set_particles(pos, None, mass)
            ctx.bin_sort()
            ctx.deposit()
            rho = ctx.density()
        assert rel_l2(rho, rho_ref) < 1e-6


# --------------------------------------------------------------------------------------------- A4
@pytest.mark.parametrize("gfunc", [refapi.DISCRETE_LAPLACIAN, refapi.S1_OPTIMAL, refapi.S2_OPTIMAL,
                                   refapi.POOR_MAN])
@pytest.mark.parametrize("grid", [(16, 16, 16), (16, 8, 12)])
def test_green_table(gfunc, grid):
    p = refapi.make_params(1, grid, (60.0, 60.0, 60.0), gfunc=gfunc, zero_degenerate=True)
    g64 = oracle("f64").green(p)
    gsym = 0.5 * (g64 + np.roll(g64[::-1, ::-1, ::-1], 1, axis=(0, 1, 2)))  # G(k) + G(-k mod N)
    with capi.Context(to_p3m(p, p3m=False, precision=capi.F64)) as ctx:
        ctx.green_init()
        g = ctx.get_green_table()
    assert rel_l2(g, gsym) < 1e-10
    # the fp32 reference table agrees away from the degenerate modes
    p.greenZeroDegenerate = 0
    g32 = oracle("f32").green(p).astype(np.float64)
    g32sym = 0.5 * (g32 + np.roll(g32[::-1, ::-1, ::-1], 1, axis=(0, 1, 2)))
    m = ~degenerate_mask(g32.shape) if gfunc in (refapi.S1_OPTIMAL, refapi.S2_OPTIMAL) else np.ones_like(g32, bool)
    assert rel_l2(g[m], g32sym[m]) < 2e-5


# ---------------------------------------------------------------------------------------- A2 + A3
@pytest.mark.parametrize("grid", [(32, 32, 32), (32, 16, 8)])
def test_poisson_with_reference_table(grid):
    """FFT + multiply alone: density and Green table taken from the oracle."""
    box = (60.0, 60.0 * grid[1] / grid[0], 60.0 * grid[2] / grid[0])
    p, pos, vel, mass = plummer_case(20000, grid=grid, box=box, gfunc=refapi.DISCRETE_LAPLACIAN)
    o = oracle("f32")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    rho = o.deposit(p, pc, mcode)
    g = o.green(p)
    phi_ref = o.poisson(p, rho, g)
    phi64 = oracle("f64").poisson(p, rho.astype(np.float64), g.astype(np.float64))
    with capi.Context(to_p3m(p, p3m=False)) as ctx:
        ctx.set_green_table(g)
        ctx.set_density(rho)
        ctx.poisson()
        phi = ctx.potential()
    assert rel_l2(phi, phi_ref) < 5e-6
    assert rel_l2(phi, phi64) < 5e-6


def test_poisson_optimal_table_matches_c2c_real_part():
    """R2C/C2R with the symmetrised table == the reference's C2C + .real()."""
    p, pos, vel, mass = plummer_case(20000)
    o = oracle("f32")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    rho = o.deposit(p, pc, mcode)
    g = o.green(p)  # literal reference table, asymmetric, noise at the degenerate modes
    phi_ref = o.poisson(p, rho, g)
    with capi.Context(to_p3m(p, p3m=False)) as ctx:
        ctx.set_green_table(g)
        ctx.set_density(rho)
        ctx.poisson()
        phi = ctx.potential()
    assert rel_l2(phi, phi_ref) < 5e-6


# --------------------------------------------------------------------------------------------- A5
@pytest.mark.parametrize("fds", [refapi.TWO_POINT, refapi.FOUR_POINT])
def test_gradient_field(fds):
    p, pos, vel, mass = plummer_case(5000, grid=(32, 16, 24), box=(60.0, 30.0, 45.0), fds=fds,
                                     gfunc=refapi.DISCRETE_LAPLACIAN)
    rng = np.random.default_rng(3)
    phi = rng.standard_normal((p.nz, p.ny, p.nx)).astype(np.float32)
    f_ref = oracle("f32").field(p, phi)
    with capi.Context(to_p3m(p, p3m=False)) as ctx:
        ctx.set_potential(phi)
        ctx.gradient()
        f = ctx.field()
    assert rel_l2(f, f_ref) < 1e-6


# --------------------------------------------------------------------------------------------- A6
@pytest.mark.parametrize("is_,fds", [(refapi.TSC, refapi.TWO_POINT), (refapi.TSC, refapi.FOUR_POINT),
                                     (refapi.CIC, refapi.TWO_POINT), (refapi.NGP, refapi.TWO_POINT)])
@pytest.mark.parametrize("case", ["plummer", "disk"])
def test_gather(case, is_, fds):
    mk = dict(plummer=plummer_case, disk=disk_case)[case]
    p, pos, vel, mass = mk(20000, is_=is_, fds=fds, gfunc=refapi.DISCRETE_LAPLACIAN)
    o = oracle("f32")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    rng = np.random.default_rng(5)
    phi = rng.standard_normal((p.nz, p.ny, p.nx)).astype(np.float32)
    acc_ref = o.gather(p, pc, o.field(p, phi))
    for p3m in (False, True):
        with capi.Context(to_p3m(p, p3m=p3m)) as ctx:
            ctx.set_particles(pos, vel, mass)
            ctx.bin_sort()
            ctx.set_potential(phi)
            ctx.gather()
            acc = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        assert rel_l2(acc, acc_ref) < 5e-6, f"p3m={p3m}"


# ----------------------------------------------------------------------------------- A7, A8, A9
@pytest.mark.parametrize("use_table,cloud", [(True, refapi.S1), (True, refapi.S2), (False, refapi.S1),
                                             (False, refapi.S2)])
def test_short_range_forces(use_table, cloud):
    p, pos, vel, mass = plummer_case(6000, use_table=use_table, cloud=cloud,
                                     gfunc=refapi.S1_OPTIMAL if cloud == refapi.S1 else refapi.S2_OPTIMAL)
    o = oracle("f32")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    sr_ref = o.sr_forces(p, pc, mcode) / mcode[:, None]
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, vel, mass)
        if use_table:
            assert rel_l2(ctx.sr_table(), o.sr_table(p)) < 1e-7
        ctx.bin_sort()
        ctx.short_range()
        _, sr = ctx.acc_parts()
    assert rel_l2(sr, sr_ref) < 2e-5
    # Newton's third law: total short-range force vanishes
    f = sr * mcode[:, None].astype(np.float64)
    assert np.abs(f.sum(axis=0)).max() < 1e-5 * np.abs(f).sum()


def test_short_range_fp64():
    p, pos, vel, mass = plummer_case(6000)
    o = oracle("f64")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    sr_ref = o.sr_forces(p, pc, mcode) / mcode[:, None]
    with capi.Context(to_p3m(p, p3m=True, precision=capi.F64)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.bin_sort()
        ctx.short_range()
        _, sr = ctx.acc_parts()
    assert rel_l2(sr, sr_ref) < 1e-9


def test_short_range_dense_and_sparse_paths_agree_with_direct_sum():
    """A tight clump (dense-cell tiled kernel) inside a sparse halo (per-target kernel)."""
    rng = np.random.default_rng(11)
    n1, n2 = 3000, 3000
    clump = 30.0 + 0.8 * rng.standard_normal((n1, 3))
    halo = 30.0 + 12.0 * (rng.random((n2, 3)) - 0.5)
    pos = np.concatenate([clump, halo]).astype(np.float32)
    mass = np.full(n1 + n2, 1.0 / (n1 + n2), np.float32)
    p = refapi.make_params(n1 + n2, (32, 32, 32), (60.0, 60.0, 60.0))
    o = oracle("f64")
    pc, _, mcode = o.to_code_units(p, pos, None, mass)
    sr_ref = o.sr_forces(p, pc, mcode) / mcode[:, None]
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, None, mass)
        ctx.bin_sort()
        ctx.short_range()
        _, sr = ctx.acc_parts()
        checked, inside = ctx.pair_counts()
    assert rel_l2(sr, sr_ref) < 2e-5
    from scipy.spatial import cKDTree
    tree = cKDTree(pc)
    re = float(p.cutoffRadius) / float(p.H)
    expect = int(tree.count_neighbors(tree, re)) - (n1 + n2)  # ordered pairs i != j within the cutoff
    assert abs(inside - expect) <= 1e-3 * expect
    assert checked >= inside


# ------------------------------------------------------------------------------- whole force step
@pytest.mark.parametrize("case,p3m", [("plummer", False), ("plummer", True), ("disk", False), ("disk", True)])
def test_force_fp32_vs_oracle(case, p3m):
    mk = dict(plummer=plummer_case, disk=disk_case)[case]
    p, pos, vel, mass = mk(20000)
    o = oracle("f32")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    g = o.green(p)
    rho_ref, phi_ref, acc_ref = o.force(p, p3m, g, pc, mcode)
    with capi.Context(to_p3m(p, p3m=p3m, zero_degenerate=True)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.force()
        rho, phi = ctx.density(), ctx.potential()
        acc = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
    assert rel_l2(rho, rho_ref) < TOL32
    # potential: compare with the degenerate modes (reference = rounding noise there) projected out
    assert rel_l2(project_out_degenerate(phi), project_out_degenerate(phi_ref)) < TOL32
    assert rel_l2(acc, acc_ref) < TOL32


@pytest.mark.parametrize("p3m", [False, True])
def test_force_fp64_vs_oracle(p3m):
    p, pos, vel, mass = plummer_case(20000, zero_degenerate=True)
    o = oracle("f64")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    g = o.green(p)
    rho_ref, phi_ref, acc_ref = o.force(p, p3m, g, pc, mcode)
    with capi.Context(to_p3m(p, p3m=p3m, precision=capi.F64)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.force()
        rho, phi = ctx.density(f64=True), ctx.potential(f64=True)
        acc = ctx.get_particles(capi.UNITS_CODE, f64=True, want=("acc",))[2]
    assert rel_l2(rho, rho_ref) < TOL64
    assert rel_l2(phi, phi_ref) < TOL64
    assert rel_l2(acc, acc_ref) < TOL64


@pytest.mark.skipif(not refapi.have_ref(), reason="oracle/_ref not built")
def test_force_fp32_vs_compiled_reference():
    """Straight against the UNMODIFIED reference (oracle/_ref), P3M, reference demo parameters."""
    p, pos, vel, mass = plummer_case(20000)
    r = refapi.Ref().p3m_force(p, pos, vel, mass)
    with capi.Context(to_p3m(p, p3m=True, zero_degenerate=True)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.force()
        acc = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        mc, cc, order = ctx.cells()
    assert np.array_equal(cc, r["cell"])
    assert rel_l2(acc, r["acc"]) < TOL32


def test_force_is_deterministic_in_short_range_and_sort():
    p, pos, vel, mass = plummer_case(20000)
    out = []
    for _ in range(2):
        with capi.Context(to_p3m(p, p3m=True)) as ctx:
            ctx.set_particles(pos, vel, mass)
            ctx.force()
            out.append((ctx.cells()[2], ctx.acc_parts()[1]))
    assert np.array_equal(out[0][0], out[1][0])
    assert np.array_equal(out[0][1], out[1][1]), "short-range sums are order-fixed, hence bit-reproducible"


# ------------------------------------------------------------------------------------ A10, A11, N1
@pytest.mark.parametrize("case,p3m", [("plummer", True), ("disk", False)])
def test_run_100_steps_energy_momentum_drift(case, p3m):
    mk = dict(plummer=plummer_case, disk=disk_case)[case]
    n, steps = 4000, 100
    p, pos, vel, mass = mk(n, gfunc=refapi.DISCRETE_LAPLACIAN if not p3m else refapi.S1_OPTIMAL)
    diag_ref, pos_ref, vel_ref, _ = oracle("f32").run(p, p3m, pos, vel, mass, steps)
    assert diag_ref.shape[0] == steps + 1 and not diag_ref[:, 11].any(), "oracle run must not escape"
    rows = []
    with capi.Context(to_p3m(p, p3m=p3m, zero_degenerate=True)) as ctx:
        # head of run(): source/pmMethod.cpp:72-77 / source/p3mMethod.cpp:73-91
        ctx.set_particles(pos, vel, mass)
        ctx.green_init()
        ctx.force()
        ctx.kick(0.5)
        for t in range(steps + 1):
            ctx.drift()
            rows.append(ctx.diagnostics())
            assert not ctx.escaped()
            ctx.force()
            ctx.kick(1.0)
        gpos, gvel, _ = ctx.get_particles(capi.UNITS_CODE)
    d = np.array(rows)
    e_ref = diag_ref[:, 0] + diag_ref[:, 1]
    e = d[:, 0] + d[:, 1]
    scale = np.abs(diag_ref[:, 1]).max()
    # energy trajectory: same drift (difference small against the kinetic energy scale)
    assert np.abs(e - e_ref).max() < 2e-3 * scale, (np.abs(e - e_ref).max(), scale)
    assert np.abs((e[-1] - e[0]) - (e_ref[-1] - e_ref[0])) < 2e-3 * scale
    # momentum and angular momentum trajectories
    pscale = np.abs(mass.astype(np.float64)[:, None] * vel).sum()
    assert np.abs(d[:, 2:5] - diag_ref[:, 2:5]).max() < 2e-3 * pscale
    lscale = np.abs(diag_ref[:, 5:8]).max() + 1e-30
    assert np.abs(d[:, 5:8] - diag_ref[:, 5:8]).max() < 5e-3 * lscale
    # trajectories stay together (chaotic divergence is still small after 100 steps)
    assert rel_l2(gpos, pos_ref) < 1e-3


def test_step_matches_manual_sequence_and_stops_on_escape():
    p, pos, vel, mass = plummer_case(2000)
    with capi.Context(to_p3m(p, p3m=True)) as a, capi.Context(to_p3m(p, p3m=True)) as b:
        for ctx in (a, b):
            ctx.set_particles(pos, vel, mass)
            ctx.force()
            ctx.kick(0.5)
        assert a.step(5) == 5
        for _ in range(5):
            b.drift(); b.force(); b.kick(1.0)
        pa = a.get_particles(capi.UNITS_CODE)
        pb = b.get_particles(capi.UNITS_CODE)
        # same kernels, same order; only the REDG flush order of the density tiles differs run to run
        assert np.allclose(pa[0], pb[0], rtol=1e-5, atol=0) and np.allclose(pa[1], pb[1], rtol=1e-3, atol=1e-7)
    # a particle that leaves the box freezes the state (source/pmMethod.cpp:108-111)
    vel2 = vel.copy()
    vel2[0] = [40.0, 0.0, 0.0]
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, vel2, mass)
        ctx.force()
        ctx.kick(0.5)
        done = ctx.step(10)
        assert done < 10
        assert ctx.escaped()


# ------------------------------------------------------------------------------------------ errors
def test_error_behaviour():
    p = capi.default_params()
    p.nx = p.ny = p.nz = 16
    p.box[:] = [8.0, 8.0, 8.0]
    p.assignment = 7
    with pytest.raises(capi.P3MError) as e:
        capi.Context(p)
    assert e.value.code == -1 and "interpolation scheme" in str(e.value)
    p.assignment = capi.TSC
    with capi.Context(p) as ctx:
        with pytest.raises(capi.P3MError) as e:
            ctx.deposit()
        assert e.value.code == -4
        ctx.set_particles(np.zeros((0, 3), np.float32), None, np.zeros(0, np.float32))
        ctx.force()  # empty input is legal
        assert ctx.density().sum() == 0


# --------------------------------------------------------- full-size, size-independent properties
def test_config2_full_size_properties():
    """BASELINE config 2 (P3M Plummer, 2^20 particles, 128^3, TSC): properties that need no oracle."""
    n = 1 << 20
    p, pos, vel, mass = plummer_case(n, grid=(128, 128, 128), softening=0.5)
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.force()
        rho = ctx.density(f64=True)
        mc, cc, order = ctx.cells()
        _, sr = ctx.acc_parts()
        gpos, _, acc = ctx.get_particles(capi.UNITS_CODE)
        pc, _, mcode = code_units(p, pos, vel, mass)
    # mass conservation of the assignment
    assert abs(rho.sum() - mcode.sum(dtype=np.float64)) < 1e-5 * mcode.sum(dtype=np.float64)
    # order is a permutation, cell-sorted
    assert np.array_equal(np.sort(order), np.arange(n))
    dims, cell = oracle("f32").chaining_cells(p, pc)
    assert np.array_equal(cc, cell)
    key = morton3(cell % dims[0], (cell // dims[0]) % dims[1], cell // (dims[0] * dims[1]))
    assert np.all(np.diff(key[order].astype(np.int64)) >= 0)
    # Newton's third law for the short-range part; finite accelerations
    f = sr * mcode[:, None].astype(np.float64)
    assert np.abs(f.sum(axis=0)).max() < 1e-4 * np.abs(f).sum()
    assert np.isfinite(acc).all()
    # a sample of particles against the direct O(N) evaluation of the same short-range sum (fp64)
    o = oracle("f64")
    tab = o.sr_table(p)
    re2 = (float(p.cutoffRadius) / float(p.H)) ** 2
    d2t = re2 / 499.0
    rng = np.random.default_rng(0)
    pc64 = pc.astype(np.float64); m64 = mcode.astype(np.float64)
    for i in rng.choice(n, 8, replace=False):
        d = pc64[i][None, :] - pc64
        r2 = (d * d).sum(1)
        sel = (r2 < re2)
        xi = r2[sel] / d2t
        t = np.minimum(xi.astype(np.int64), 498)
        fr = xi - t
        F = tab[t] + fr * (tab[t + 1] - tab[t])
        a = (m64[sel] * F)[:, None] * d[sel]
        assert np.allclose(sr[i], a.sum(0), rtol=2e-3, atol=1e-3 * np.abs(a).sum(0).max() + 1e-30)


# ------------------------------------------------------- straight against the reference's own outputs
import glob as _glob
import os as _os

_GOLDEN = sorted(_glob.glob(_os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "golden", "*.npz")))


@pytest.mark.parametrize("path", _GOLDEN, ids=[_os.path.basename(g)[:-4] for g in _GOLDEN])
def test_gpu_vs_golden_reference_outputs(path):
    """tests/golden/*.npz were written by the UNMODIFIED reference (tests/golden/make_golden.py)."""
    from test_oracle import load
    z, p, p3m = load(path)
    with capi.Context(to_p3m(p, p3m=p3m, zero_degenerate=False)) as ctx:
        ctx.set_particles(z["pos"], z["vel"], z["mass"])
        ctx.set_green_table(z["green"])  # the reference's own table, incl. its noise modes
        ctx.force()
        rho, phi = ctx.density(), ctx.potential()
        acc = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        pm, sr = ctx.acc_parts()
        mc, cc, order = ctx.cells()
        gpos = ctx.get_particles(capi.UNITS_CODE, want=("pos",))[0]
    assert np.array_equal(gpos, z["pos_code"])
    assert rel_l2(rho, z["density"]) < TOL32
    assert rel_l2(phi, z["potential"]) < TOL32
    assert rel_l2(pm, z["acc_pm"]) < TOL32
    if p3m:
        assert np.array_equal(cc, z["cell"])
        assert rel_l2(sr * z["mass_code"][:, None].astype(np.float64), z["sr_force"]) < TOL32
        assert rel_l2(acc, z["acc"]) < TOL32


def test_p3m_edge_inputs():
    """Empty set, one particle, two coincident particles, a particle on the upper box face (reference:
    cell index -1 -> undefined; here clamped into the last cell): no crash, finite, symmetric."""
    p = refapi.make_params(4, (16, 16, 16), (8.0, 8.0, 8.0), H=1.0, zero_degenerate=True)
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(np.zeros((0, 3), np.float32), None, np.zeros(0, np.float32))
        ctx.force()
        assert ctx.density().sum() == 0
        ctx.set_particles(np.array([[4.0, 4.0, 4.0]], np.float32), None, np.ones(1, np.float32))
        ctx.force()
        a1 = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        assert np.isfinite(a1).all()
        pos = np.array([[3.0, 3.0, 3.0], [3.0, 3.0, 3.0], [5.0, 3.0, 3.0], [8.0, 8.0, 8.0]], np.float32)
        ctx.set_particles(pos, None, np.ones(4, np.float32))
        ctx.force()
        acc = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        _, sr = ctx.acc_parts()
    assert np.isfinite(acc).all()
    assert np.array_equal(sr[0], sr[1])  # coincident particles feel the same force, none from each other
    assert sr[2, 0] < 0 < sr[0, 0]       # attraction along x between the pair and the third particle


# ------------------------------------------------------------------------ round-1 additions (kernels)
@pytest.mark.parametrize("precision", [capi.F32, capi.F64])
@pytest.mark.parametrize("nz", [16, 32, 64, 128, 256, 512, 1024])
def test_fused_z_pass_matches_cufft_z_leg(nz, precision, monkeypatch):
    """k_poisson_z (forward z FFT + influence-function multiply + inverse z FFT in one pass, every radix
    plan 2^4..2^10) against the all-cuFFT 3-D transform + multiply kernel on the same density."""
    grid = (8, 4, nz)
    box = (60.0, 30.0, 60.0 * nz / 8)
    p = refapi.make_params(16, grid, box, gfunc=refapi.DISCRETE_LAPLACIAN)
    rng = np.random.default_rng(nz)
    rho = rng.standard_normal((nz, grid[1], grid[0])).astype(np.float32)
    out = {}
    for mode in ("fused", "cufft"):
        if mode == "cufft":
            monkeypatch.setenv("P3M_TUNE_CUFFT_Z", "1")
        else:
            monkeypatch.delenv("P3M_TUNE_CUFFT_Z", raising=False)
        with capi.Context(to_p3m(p, p3m=False, precision=precision)) as ctx:
            ctx.green_init()
            ctx.set_density(rho)
            ctx.poisson()
            out[mode] = ctx.potential(f64=True)
    monkeypatch.delenv("P3M_TUNE_CUFFT_Z", raising=False)
    tol = 2e-6 if precision == capi.F32 else 1e-13
    assert rel_l2(out["fused"], out["cufft"]) < tol


def test_short_range_with_unequal_masses_uses_the_general_kernel():
    """Every sampler of the reference hands out equal masses (folded into the force table); unequal masses
    take the per-pair mass multiply.  Both against the fp64 oracle, and against each other."""
    p, pos, vel, mass = plummer_case(6000)
    rng = np.random.default_rng(3)
    mass2 = (mass * rng.uniform(0.5, 1.5, mass.shape)).astype(np.float32)
    o = oracle("f64")
    res = {}
    for name, m in (("equal", mass), ("unequal", mass2)):
        pc, _, mcode = o.to_code_units(p, pos, vel, m)
        sr_ref = o.sr_forces(p, pc, mcode) / mcode[:, None]
        with capi.Context(to_p3m(p, p3m=True)) as ctx:
            ctx.set_particles(pos, vel, m)
            ctx.bin_sort()
            ctx.short_range()
            _, sr = ctx.acc_parts()
        assert rel_l2(sr, sr_ref) < 2e-5, name
        res[name] = sr
    # one particle 1 ulp heavier switches the kernel, not the physics
    mass3 = mass.copy()
    mass3[17] = np.nextafter(mass3[17], np.float32(1))
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, vel, mass3)
        ctx.bin_sort()
        ctx.short_range()
        _, sr3 = ctx.acc_parts()
    # (the equal-mass set runs the packed kernel, which sums even and odd sources separately)
    assert rel_l2(sr3, res["equal"]) < 5e-6


def test_pm_short_key_sort_stays_sorted_over_steps():
    """PM-only contexts sort on (tile, mesh cell) without the id: after any number of steps the particles
    are a permutation in non-decreasing key order, and two identical runs give the identical order."""
    p, pos, vel, mass = uniform_case(20000, gfunc=0)
    rng = np.random.default_rng(5)
    vel = (0.4 * rng.standard_normal(vel.shape)).astype(np.float32)
    orders = []
    for _ in range(2):
        with capi.Context(to_p3m(p, p3m=False)) as ctx:
            ctx.set_particles(pos, vel, mass)
            ctx.force()
            ctx.kick(0.5)
            ctx.step(4)
            ctx.bin_sort()
            mc, _, order = ctx.cells()
            gpos, _, _ = ctx.get_particles(capi.UNITS_CODE, want=("pos",))
        assert np.array_equal(np.sort(order), np.arange(len(mass)))
        t = gpos.astype(np.int32)
        key = (morton3(t[:, 0] >> 3, t[:, 1] >> 3, t[:, 2] >> 3) << np.uint64(9)) | \
              ((t[:, 2] & 7).astype(np.uint64) << np.uint64(6)) | ((t[:, 1] & 7).astype(np.uint64) << np.uint64(3)) | \
              (t[:, 0] & 7).astype(np.uint64)
        assert np.all(np.diff(key[order].astype(np.int64)) >= 0)
        orders.append(order)
    assert np.array_equal(orders[0], orders[1])


@pytest.mark.parametrize("p3m,n,sigma", [(False, 20000, 0.08), (False, 300000, 0.03), (True, 20000, 0.01), (True, 200000, 0.005),
                                         (False, 5000, 0.0)])
def test_incremental_sort_equals_full_sort(monkeypatch, p3m, n, sigma):
    """The per-step re-sort merges only the movers (incsort.cu: mover mask, hand-written radix sort of the movers,
    rank merge).  Its result must be, element for element, what the stable full radix sort of every particle gives
    (P3M_TUNE_FULL_SORT=1): same order after every step, same positions -- and it must really have been used."""
    p, pos, vel, mass = uniform_case(n, gfunc=0) if not p3m else uniform_case(n)
    rng = np.random.default_rng(7)
    vel = (sigma * rng.standard_normal(vel.shape)).astype(np.float32)
    runs = {}
    monkeypatch.setenv("P3M_TUNE_INC_SORT_DEN", "3")  # merge up to n / 3 movers (default: n / 12, the measured break-even)
    for name, env in (("incremental", None), ("full", "1")):
        if env:
            monkeypatch.setenv("P3M_TUNE_FULL_SORT", env)
        else:
            monkeypatch.delenv("P3M_TUNE_FULL_SORT", raising=False)
        orders = []
        with capi.Context(to_p3m(p, p3m=p3m)) as ctx:
            ctx.set_particles(pos, vel, mass)
            ctx.force()
            ctx.kick(0.5)
            for _ in range(6):
                ctx.step(1)
                _, _, order = ctx.cells()
                orders.append(order.copy())
            st = ctx.stats()
            gpos = ctx.get_particles(capi.UNITS_CODE, want=("pos",))[0]
        runs[name] = (orders, gpos, st)
    monkeypatch.delenv("P3M_TUNE_FULL_SORT", raising=False)
    inc, full = runs["incremental"], runs["full"]
    assert inc[2]["incremental_sorts"] >= 5 and full[2]["incremental_sorts"] == 0, (inc[2], full[2])
    assert 0 <= inc[2]["sort_movers"] <= n / 3 and (sigma == 0.0 or inc[2]["sort_movers"] > 0), inc[2]
    for a, b in zip(inc[0], full[0]):
        assert np.array_equal(np.sort(a), np.arange(n))
        assert np.array_equal(a, b)
    # same order every step => same arithmetic, up to the unordered floating-point REDs of the density flush
    assert np.abs(inc[1] - full[1]).max() < 1e-4


# ------------------------------------------------------------------------ round-2 additions (kernels)
def test_packed_fp32_pp_kernel_matches_the_scalar_kernel_and_the_oracle(monkeypatch):
    """k_pp_packed (FADD2 / FMUL2 / FFMA2 pair body, the default for fp32 + table + equal masses) against
    k_pp_tiled (P3M_TUNE_SCALAR_PP=1) and against the fp64 oracle, on a set with a dense core."""
    p, pos, vel, mass = plummer_case(30000)
    o = oracle("f64")
    pc, _, mcode = o.to_code_units(p, pos, vel, mass)
    sr_ref = o.sr_forces(p, pc, mcode) / mcode[:, None]
    res = {}
    for name, env in (("packed", None), ("scalar", "1")):
        if env:
            monkeypatch.setenv("P3M_TUNE_SCALAR_PP", env)
        else:
            monkeypatch.delenv("P3M_TUNE_SCALAR_PP", raising=False)
        with capi.Context(to_p3m(p, p3m=True)) as ctx:
            ctx.set_particles(pos, vel, mass)
            ctx.bin_sort()
            ctx.short_range()
            _, sr = ctx.acc_parts()
            assert bool(ctx.stats()["packed_pp"]) == (name == "packed")
            before = ctx.acc_parts()[1]
            checked, inside = ctx.pair_counts()  # read-only: no second accumulation
            assert np.array_equal(ctx.acc_parts()[1], before) and checked >= inside > 0
        assert rel_l2(sr, sr_ref) < 2e-5, name
        res[name] = sr
    monkeypatch.delenv("P3M_TUNE_SCALAR_PP", raising=False)
    assert rel_l2(res["packed"], res["scalar"]) < 1e-5  # summation order only


def test_stale_accelerations_are_refused_not_returned():
    """A re-sort permutes positions / velocities / ids but not the accelerations: kick, acceleration readbacks
    and diagnostics must fail with P3M_ESTATE until the next gather."""
    p, pos, vel, mass = plummer_case(3000)
    with capi.Context(to_p3m(p, p3m=True)) as ctx:
        ctx.set_particles(pos, vel, mass)
        ctx.force()
        a1 = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        ctx.drift()
        ctx.bin_sort()
        for call in (lambda: ctx.kick(1.0), lambda: ctx.get_particles(capi.UNITS_CODE, want=("acc",)),
                     ctx.acc_parts, ctx.diagnostics):
            with pytest.raises(capi.P3MError) as e:
                call()
            assert e.value.code == -4
        ctx.deposit(); ctx.poisson(); ctx.gather(); ctx.short_range()
        a2 = ctx.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        assert np.isfinite(a2).all() and rel_l2(a2, a1) < 0.5

"""Pins the CPU oracle (oracle/p3m_oracle.c) -- against the committed golden vectors that the
UNMODIFIED reference produced (tests/golden/make_golden.py), and, where oracle/_ref is present, against
the reference itself on fresh inputs."""
import glob
import os
import tempfile

import numpy as np
import pytest

import refapi
from common import plummer_case, disk_case
from refapi import Oracle, rel_l2

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))


def load(path):
    z = np.load(path)
    p = refapi.Params()
    for name, _ in p._fields_:
        v = z["param_" + name]
        if v.ndim:
            getattr(p, name)[:] = [float(x) for x in v]
        else:
            setattr(p, name, v.item())
    return z, p, bool(z["p3m"])


def test_golden_files_present():
    assert len(GOLDEN) >= 5


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_fp32_restatement_reproduces_the_reference(path):
    z, p, p3m = load(path)
    o = Oracle("f32")
    pc, vc, mc = o.to_code_units(p, z["pos"], z["vel"], z["mass"])
    assert np.array_equal(pc, z["pos_code"]) and np.array_equal(mc, z["mass_code"])
    g = o.green(p)
    assert np.array_equal(g, z["green"]), "influence function: bit-exact (same libm, same order)"
    rho = o.deposit(p, pc, mc)
    assert np.array_equal(rho, z["density"])
    # the FFT library differs (own mixed radix vs the reference's kissfft): fp32 round-off only
    phi = o.poisson(p, rho, g)
    assert rel_l2(phi, z["potential"]) < 2e-6
    assert np.array_equal(o.field(p, z["potential"]), z["field"])
    assert np.array_equal(o.gather(p, pc, z["field"]), z["acc_pm"])
    if p3m:
        dims, cell = o.chaining_cells(p, pc)
        assert np.array_equal(dims, z["chain_dims"]) and np.array_equal(cell, z["cell"])
        assert np.array_equal(o.chaining_order(p, pc), z["order"])
        if p.useTable:
            assert np.array_equal(o.sr_table(p), z["ftable"])
        assert rel_l2(o.sr_forces(p, pc, mc), z["sr_force"]) < 1e-6
    _, _, acc = o.force(p, p3m, g, pc, mc)
    assert rel_l2(acc, z["acc"] if p3m else z["acc_pm"]) < 5e-6


@pytest.mark.parametrize("path", GOLDEN, ids=[os.path.basename(g)[:-4] for g in GOLDEN])
def test_run_loop_diagnostics_follow_the_reference(path):
    z, p, p3m = load(path)
    diag, po, vo, ao = Oracle("f32").run(p, p3m, z["pos"], z["vel"], z["mass"], 20)
    ref = z["run_diag"]  # pe ke | px py pz | Lx Ly Lz | ex ey ez  (text files, 6 significant digits)
    assert diag.shape[0] == ref.shape[0] == 21
    scale = np.abs(ref[:, 1]).max()
    # energy.txt is written with std::to_string: 6 DECIMALS (source/stateRecorder.cpp:71-76), i.e. an
    # absolute resolution of 1e-6 on values of order 1e-4
    assert np.abs(diag[:, 0] - ref[:, 0]).max() < 5e-4 * max(np.abs(ref[:, 0]).max(), scale) + 1.1e-6
    assert np.abs(diag[:, 1] - ref[:, 1]).max() < 5e-5 * scale + 1.1e-6
    pscale = np.abs(z["mass"][:, None].astype(np.float64) * z["vel"]).sum()
    assert np.abs(diag[:, 2:5] - ref[:, 2:5]).max() < 1e-4 * pscale
    assert np.abs(diag[:, 5:8] - ref[:, 5:8]).max() < 1e-4 * (np.abs(ref[:, 5:8]).max() + pscale)
    assert np.abs(diag[:, 8:11] - ref[:, 8:11]).max() < 1e-4 * (np.abs(ref[:, 8:11]).max() + pscale)
    assert rel_l2(po, z["run_pos"]) < 1e-5


@pytest.mark.parametrize("path", GOLDEN[:3], ids=[os.path.basename(g)[:-4] for g in GOLDEN[:3]])
def test_fp64_restatement_is_the_same_algorithm(path):
    """Same inputs, double arithmetic: differs from the fp32 reference by fp32 round-off only."""
    z, p, p3m = load(path)
    o = Oracle("f64")
    pc, vc, mc = o.to_code_units(p, z["pos"], z["vel"], z["mass"])
    rho = o.deposit(p, pc, mc)
    assert rel_l2(z["density"], rho) < 2e-6
    g = o.green(p) if p.gfunc in (refapi.DISCRETE_LAPLACIAN, refapi.POOR_MAN) else z["green"].astype(np.float64)
    _, phi, acc = o.force(p, p3m, g, pc, mc)
    assert rel_l2(z["potential"], phi) < 2e-5
    assert rel_l2(z["acc"] if p3m else z["acc_pm"], acc) < 5e-5


def test_fft_contract_round_trip():
    """test/fftAdaptersTest.cpp:6-22: ifft(fft(x)) == x with the inverse normalised -- here via the
    Poisson operator with G == 1 (identity), incl. non-power-of-two lengths."""
    for grid in [(2, 2, 2), (8, 4, 2), (6, 10, 15), (7, 3, 5)]:
        p = refapi.make_params(1, grid, (6.0, 6.0, 6.0), H=1.0)
        rng = np.random.default_rng(1)
        rho = rng.standard_normal((grid[2], grid[1], grid[0]))
        for prec, tol in (("f32", 2e-6), ("f64", 1e-13)):
            o = Oracle(prec)
            ones = np.ones_like(rho)
            out = o.poisson(p, rho, ones)
            want = rho - rho.mean()  # the zero mode is cleared (source/pmMethod.cpp:341)
            assert rel_l2(out, want) < tol, (grid, prec)


@pytest.mark.skipif(not refapi.have_ref(), reason="oracle/_ref not built (needs /root/reference)")
@pytest.mark.parametrize("seed", [1, 2])
def test_against_live_reference(seed):
    p, pos, vel, mass = plummer_case(1500, grid=(16, 16, 16), seed=seed)
    ref = refapi.Ref()
    r = ref.p3m_force(p, pos, vel, mass)
    o = Oracle("f32")
    pc, _, mc = o.to_code_units(p, pos, vel, mass)
    dims, cell = o.chaining_cells(p, pc)
    assert np.array_equal(cell, r["cell"]) and np.array_equal(o.chaining_order(p, pc), r["order"])
    _, _, acc = o.force(p, True, o.green(p), pc, mc)
    assert rel_l2(acc, r["acc"]) < 5e-6
    with tempfile.TemporaryDirectory() as d:
        diag_ref, po, vo, _ = ref.run(p, pos, vel, mass, 5, True, d)
    diag, po2, vo2, _ = o.run(p, True, pos, vel, mass, 5)
    assert rel_l2(po2, po) < 1e-5 and rel_l2(vo2, vo) < 1e-4

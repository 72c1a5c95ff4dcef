"""N3: device-side initial conditions (csrc/ics.cu, p3m_generate_particles / p3m_sample_particles).

The reference's random streams are implementation-defined, so the DISTRIBUTIONS are what is
checked: two-sample Kolmogorov-Smirnov distance <= 1e-2 of radii / speeds / heights against the numpy
restatement (particlesimulation_b200/ics.py) and, where oracle/_ref is built, against the compiled reference's
own samplers (source/plummerSampler.cpp:11-83, source/diskSamplerLinear.cpp:10-74)."""
import numpy as np
import pytest

import refapi
from common import plummer_case, to_p3m
from particlesimulation_b200 import capi, ics
from refapi import rel_l2

pytestmark = pytest.mark.gpu
N = 400_000


def ks(a, b):
    a, b = np.sort(np.asarray(a, np.float64)), np.sort(np.asarray(b, np.float64))
    allv = np.concatenate([a, b])
    return float(np.abs(np.searchsorted(a, allv, side="right") / a.size -
                        np.searchsorted(b, allv, side="right") / b.size).max())


def radial(pos, c):
    return np.linalg.norm(pos.astype(np.float64) - np.asarray(c)[None, :], axis=1)


def test_plummer_sampler_distributions():
    c = (30.0, 30.0, 30.0)
    pos, vel, mass = capi.sample_particles(capi.ic_plummer(N, center=c, a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=7))
    rpos, rvel, rmass = ics.plummer(N, center=c, a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=11)
    assert ks(radial(pos, c), radial(rpos, c)) < 1e-2
    assert ks(np.linalg.norm(vel, axis=1), np.linalg.norm(rvel, axis=1)) < 1e-2
    assert np.array_equal(mass, rmass)
    r = radial(pos, c)
    # analytic CDF r^3 / (r^2 + a^2)^(3/2), clamped onto the r_max shell as the reference does (:50-52)
    assert r.max() <= 15.0 + 1e-4 and abs((r >= 15.0 - 1e-4).mean() - (1 - 15.0 ** 3 / (15.0 ** 2 + 4.0) ** 1.5)) < 2e-3
    u = (pos - np.asarray(c, np.float32)) / np.maximum(r, 1e-9)[:, None]
    assert np.abs(u.mean(0)).max() < 5e-3 and np.abs((u * u).mean(0) - 1 / 3).max() < 5e-3  # isotropic
    assert np.abs((vel / np.linalg.norm(vel, axis=1)[:, None]).mean(0)).max() < 5e-3
    # truncated variant: no shell
    pos_t, _, _ = capi.sample_particles(capi.ic_plummer(N, center=c, a=2.0, r_max=15.0, seed=7, truncate=True))
    rt = radial(pos_t, c)
    assert rt.max() < 15.0 + 1e-4 and (rt >= 15.0 - 1e-3).mean() < 1e-3
    if refapi.have_ref():
        qpos, qvel = refapi.Ref().sample_plummer(42, c, 2.0, 15.0, 1.0, 4.5e-3, 100_000)
        assert ks(radial(pos, c), radial(qpos, c)) < 1e-2
        assert ks(np.linalg.norm(vel, axis=1), np.linalg.norm(qvel, axis=1)) < 1.5e-2  # reference: Newton, tol 1e-3


def test_disk_linear_sampler_distributions():
    c = (30.0, 30.0, 15.0)
    pos, vel, mass = capi.sample_particles(capi.ic_disk_linear(N, center=c, seed=3))
    rpos, rvel, rmass = ics.disk_linear(N, center=c, seed=5)
    rho = lambda p: np.linalg.norm(p[:, :2].astype(np.float64) - np.asarray(c[:2])[None, :], axis=1)
    assert ks(rho(pos), rho(rpos)) < 1e-2
    assert ks(pos[:, 2], rpos[:, 2]) < 1e-2
    assert ks(np.linalg.norm(vel, axis=1), np.linalg.norm(rvel, axis=1)) < 1e-2
    assert np.array_equal(mass, rmass)
    # velocities are tangential
    d = pos[:, :2].astype(np.float64) - np.asarray(c[:2])
    assert np.abs((d * vel[:, :2]).sum(1)).max() < 1e-4 and np.all(vel[:, 2] == 0)
    if refapi.have_ref():
        qpos, qvel = refapi.Ref().sample_disk_linear(42, c, 3.0, 60.0, 15.0, 15.0, 0.3, 4.5e-3, 100_000)
        assert ks(rho(pos), rho(qpos)) < 1.5e-2  # reference: Newton with tolerance 0.01
        assert ks(np.linalg.norm(vel, axis=1), np.linalg.norm(qvel, axis=1)) < 1.5e-2


def test_uniform_and_disk_halo_samplers():
    lo, hi = (2.0, 3.0, 4.0), (50.0, 40.0, 30.0)
    pos, vel, mass = capi.sample_particles(capi.ic_uniform(N, lo, hi, total_mass=2.0, vel_sigma=0.25, seed=9))
    assert np.all(pos >= np.asarray(lo, np.float32)) and np.all(pos < np.asarray(hi, np.float32) + 1e-5)
    assert np.abs(pos.mean(0) - (np.asarray(lo) + np.asarray(hi)) / 2).max() < 0.1
    assert np.abs(vel.std(0) - 0.25).max() < 2e-3 and np.abs(vel.mean(0)).max() < 2e-3
    assert np.allclose(mass, 2.0 / N)
    dpos, dvel, dmass = capi.sample_particles(capi.ic_disk_halo(N, seed=4))
    rpos, rvel, _ = ics.clustered_disk_halo(N, seed=6)
    c = (30.0, 30.0, 30.0)
    h = N // 2
    assert ks(radial(dpos[h:], c), radial(rpos[h:], c)) < 1e-2       # halo radii (truncated Plummer, no shell)
    assert ks(radial(dpos[:h], c), radial(rpos[:h], c)) < 1e-2       # disk radii
    assert ks(np.linalg.norm(dvel[:h], axis=1), np.linalg.norm(rvel[:h], axis=1)) < 1e-2
    assert (radial(dpos[h:], c) >= 27.0 - 1e-3).mean() < 1e-3
    assert np.abs(dpos[:h, 1] - 30.0).max() <= 1.5 + 1e-4              # the disk lies in the x-z plane


def test_generated_set_is_the_sampled_set_and_counter_based():
    """p3m_generate_particles fills the context with exactly the particles p3m_sample_particles returns, any
    sub-range can be re-created on its own, and the force on the generated set equals the force on the uploaded
    one."""
    p, _, _, _ = plummer_case(20000)
    ic = capi.ic_plummer(20000, center=(30.0, 30.0, 30.0), a=2.0, r_max=15.0, seed=42)
    pos, vel, mass = capi.sample_particles(ic)
    part = capi.sample_particles(ic, first=5000, count=300)
    assert np.array_equal(part[0], pos[5000:5300]) and np.array_equal(part[1], vel[5000:5300])
    with capi.Context(to_p3m(p, p3m=True)) as a, capi.Context(to_p3m(p, p3m=True)) as b:
        a.generate_particles(ic)
        b.set_particles(pos, vel, mass)
        assert a.n == b.n == 20000
        ga, gb = a.get_particles(capi.UNITS_CODE, want=("pos", "vel")), b.get_particles(capi.UNITS_CODE, want=("pos", "vel"))
        assert np.array_equal(ga[0], gb[0]) and np.array_equal(ga[1], gb[1])
        a.force(); b.force()
        assert a.stats()["uniform_mass_table"] == 1.0
        (pma, sra), (pmb, srb) = a.acc_parts(), b.acc_parts()
        assert np.array_equal(sra, srb), "same particles, same order, same masses: the short-range sums are bit-identical"
        assert rel_l2(a.density(), b.density()) < 1e-6
        # the mesh part inherits the run-to-run REDG order of the P3M-context deposit (1e-7 on the density), amplified
        # by the differencing of a potential that is ~1e4 times larger than its cell-to-cell change
        assert rel_l2(pma, pmb) < 5e-3
        fa, fb = a.get_particles(capi.UNITS_CODE, want=("acc",))[2], b.get_particles(capi.UNITS_CODE, want=("acc",))[2]
        assert rel_l2(fa, fb) < 2e-4

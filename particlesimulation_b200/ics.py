"""Synthetic initial conditions of the bench harness.

numpy restatements of the DISTRIBUTIONS of the reference samplers -- not of their random streams,
which are implementation-defined (std::default_random_engine).  The same arrays are fed
to every implementation under comparison.

* plummer():      source/plummerSampler.cpp:11-83  (positions by inverse CDF, speeds by the
                  q^2 (1-q^2)^3.5 law, isotropic directions)
* disk_linear():  source/diskSamplerLinear.cpp:10-74 (linearly decreasing surface density, circular
                  velocities from the bulge + disk field)
* uniform_cube(): uniform positions, zero velocities (BASELINE.json config 3)
"""
from __future__ import annotations

import numpy as np


def _isotropic(rng, n):
    phi = 2 * np.pi * rng.random(n)
    cos_t = 1 - 2 * rng.random(n)
    sin_t = np.sqrt(np.maximum(0.0, 1 - cos_t * cos_t))
    return np.stack([sin_t * np.cos(phi), sin_t * np.sin(phi), cos_t], axis=1)


def plummer(n, center=(30.0, 30.0, 30.0), a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=42):
    rng = np.random.default_rng(seed)
    u = np.maximum(rng.random(n), 1e-12)
    r = a / np.sqrt(u ** (-2.0 / 3.0) - 1.0)
    r = np.minimum(r, r_max)  # source/plummerSampler.cpp:50-52
    pos = np.asarray(center)[None, :] + r[:, None] * _isotropic(rng, n)
    # speed: q in [0,1] with density q^2 (1-q^2)^(7/2), by rejection (max of g is ~0.092)
    q = np.empty(n)
    todo = np.arange(n)
    while todo.size:
        x = rng.random(todo.size)
        y = 0.1 * rng.random(todo.size)
        ok = y < x * x * (1 - x * x) ** 3.5
        q[todo[ok]] = x[ok]
        todo = todo[~ok]
    v_esc = np.sqrt(2 * G * M / np.sqrt(r * r + a * a))
    vel = (q * v_esc)[:, None] * _isotropic(rng, n)
    mass = np.full(n, M / n)
    return pos.astype(np.float32), vel.astype(np.float32), mass.astype(np.float32)


def _sph_rad_decr_field(pos, center, R, M, G):
    d = pos - np.asarray(center)[None, :]
    r = np.linalg.norm(d, axis=1)
    g = np.where(r > R, -G * M / np.maximum(r, 1e-30) ** 2, -(G * M / R ** 3) * r * (4 - 3 * r / R))
    return g[:, None] * d / np.maximum(r, 1e-30)[:, None]


def disk_linear(n, center=(30.0, 30.0, 15.0), rb=3.0, mb=60.0, rd=15.0, md=15.0, thickness=0.3,
                G=4.5e-3, seed=42, r0=0.0):
    rng = np.random.default_rng(seed)
    phi = 2 * np.pi * rng.random(n)
    # surface density ~ (rd - r) on [r0, rd]: CDF F(r) solves the cubic of
    # include/diskSamplerLinear.h:25-30; inverted here by bisection
    cdf = rng.random(n)
    lo = np.full(n, r0); hi = np.full(n, rd)
    norm = (rd - r0) ** 2 * (2 * r0 + rd)
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        F = (3 * rd * (mid ** 2 - r0 ** 2) - 2 * (mid ** 3 - r0 ** 3)) / norm
        hi = np.where(F > cdf, mid, hi)
        lo = np.where(F > cdf, lo, mid)
    r = 0.5 * (lo + hi)
    z = (thickness / 2) * (2 * rng.random(n) - 1)
    pos = np.asarray(center)[None, :] + np.stack([r * np.cos(phi), r * np.sin(phi), z], axis=1)
    # circular speed from bulge + disk field, source/diskSamplerLinear.cpp:37-65
    rvec = pos - np.asarray(center)[None, :]
    rr = np.linalg.norm(rvec, axis=1)
    rho = np.linalg.norm(rvec[:, :2], axis=1)
    ra = rr / rd
    sigma0 = 3 * md / (np.pi * rd * rd)
    k, h = 2.5, 0.66
    aa = -k / (h * h)
    gd_val = -G * sigma0 * (aa * (ra - h) ** 2 + k)
    gb = _sph_rad_decr_field(pos, center, rb, mb, G)
    gd = np.zeros_like(rvec)
    gd[:, :2] = gd_val[:, None] * rvec[:, :2] / np.maximum(rho, 1e-30)[:, None]
    g_val = np.linalg.norm(gb + gd, axis=1)
    v = np.sqrt(g_val * rho * rho / np.maximum(rr, 1e-30))
    vel = np.stack([-v * rvec[:, 1] / np.maximum(rho, 1e-30), v * rvec[:, 0] / np.maximum(rho, 1e-30),
                    np.zeros(n)], axis=1)
    mass = np.full(n, md / n)
    return pos.astype(np.float32), vel.astype(np.float32), mass.astype(np.float32)


def uniform_cube(n, lo, hi, total_mass=1.0, seed=42):
    rng = np.random.default_rng(seed)
    lo = np.asarray(lo, np.float64); hi = np.asarray(hi, np.float64)
    pos = lo[None, :] + (hi - lo)[None, :] * rng.random((n, 3))
    vel = np.zeros((n, 3))
    mass = np.full(n, total_mass / n)
    return pos.astype(np.float32), vel.astype(np.float32), mass.astype(np.float32)


def clustered_disk_halo(n, box=60.0, seed=42, halo_a=12.0, halo_rmax=27.0, disk_rd=24.0, thickness=3.0, M=1.0,
                        G=4.5e-3):
    """BASELINE.json config 4 ("clustered disk + halo"): half the particles in a disk with linearly
    decreasing surface density (triangular radial law), half in a Plummer halo of scale `halo_a`, both centred
    in the box; equal masses.  The disk lies in the x-z plane (normal along y), so the z-slab decomposition of
    the multi-GPU path cuts through it instead of handing the whole disk to one rank.  float32 throughout and O(10) vectorised passes, so that 2^26 particles are
    generated in seconds; velocities are circular (disk, from the enclosed mass) / isotropic at half the local
    escape speed (halo) -- synthetic data for throughput runs, not an equilibrium model."""
    rng = np.random.default_rng(seed)
    f32 = np.float32
    nd = n // 2
    nh = n - nd
    c = f32(box / 2)
    pos = np.empty((n, 3), f32)
    vel = np.zeros((n, 3), f32)
    # disk
    u = rng.random(nd, dtype=f32)
    r = f32(disk_rd) * (f32(1) - np.sqrt(f32(1) - u))
    phi = f32(2 * np.pi) * rng.random(nd, dtype=f32)
    cs, sn = np.cos(phi), np.sin(phi)
    pos[:nd, 0] = c + r * cs
    pos[:nd, 2] = c + r * sn
    pos[:nd, 1] = c + f32(thickness) * (rng.random(nd, dtype=f32) - f32(0.5))
    menc = f32(M) * (f32(0.5) * (f32(1) - (f32(1) - r / f32(disk_rd)) ** 2) +
                     f32(0.5) * (r ** 3) / (r * r + f32(halo_a) ** 2) ** f32(1.5))
    v = np.sqrt(f32(G) * menc / np.maximum(r, f32(1e-3)))
    vel[:nd, 0] = -v * sn
    vel[:nd, 2] = v * cs
    del u, r, phi, cs, sn, menc, v
    # halo
    # Plummer CDF truncated at halo_rmax (u scaled to [0, F(rmax)]): no particles beyond the cut and no shell of
    # clamped ones at it; evaluated in float64 (u^(-2/3) - 1 loses everything near u = 1 in float32)
    frac = halo_rmax ** 3 / (halo_rmax ** 2 + halo_a ** 2) ** 1.5
    u = np.maximum(rng.random(nh), 1e-12) * frac
    r = np.minimum(halo_a / np.sqrt(u ** (-2.0 / 3.0) - 1.0), halo_rmax).astype(f32)
    cos_t = f32(1) - f32(2) * rng.random(nh, dtype=f32)
    sin_t = np.sqrt(np.maximum(f32(0), f32(1) - cos_t * cos_t))
    phi = f32(2 * np.pi) * rng.random(nh, dtype=f32)
    pos[nd:, 0] = c + r * sin_t * np.cos(phi)
    pos[nd:, 1] = c + r * sin_t * np.sin(phi)
    pos[nd:, 2] = c + r * cos_t
    vesc = np.sqrt(f32(2 * G * M) / np.sqrt(r * r + f32(halo_a) ** 2))
    cos_t = f32(1) - f32(2) * rng.random(nh, dtype=f32)
    sin_t = np.sqrt(np.maximum(f32(0), f32(1) - cos_t * cos_t))
    phi = f32(2 * np.pi) * rng.random(nh, dtype=f32)
    vel[nd:, 0] = f32(0.5) * vesc * sin_t * np.cos(phi)
    vel[nd:, 1] = f32(0.5) * vesc * sin_t * np.sin(phi)
    vel[nd:, 2] = f32(0.5) * vesc * cos_t
    mass = np.full(n, M / n, f32)
    return pos, vel, mass

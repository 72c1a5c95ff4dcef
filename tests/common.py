"""Shared fixtures of the parity tests: one physical configuration expressed twice -- as the
oracle's OrcParams (tests/refapi.py) and as the product's p3m_params (include/p3m_b200.h)."""
from __future__ import annotations

import numpy as np

import refapi
from particlesimulation_b200 import capi, ics


def to_p3m(p: refapi.Params, *, p3m: bool, precision=capi.F32, timing=False, roundtrip=True,
           zero_degenerate=None) -> capi.P3MParams:
    q = capi.default_params()
    q.nx, q.ny, q.nz = p.nx, p.ny, p.nz
    q.box[:] = list(p.box)
    q.H, q.DT, q.G = p.H, p.DT, p.G
    q.assignment, q.fd_scheme, q.greens_function = p.is_, p.fds, p.gfunc
    q.particle_diameter = p.particleDiameter
    q.p3m = int(p3m)
    q.cutoff_radius, q.softening = p.cutoffRadius, p.softening
    q.cloud_shape, q.use_sr_table = p.cloudShape, p.useTable
    q.ext_kind = p.extKind
    q.ext_center[:] = list(p.extCenter)
    q.ext_R, q.ext_M = p.extR, p.extM
    q.precision = precision
    q.unit_roundtrip = int(roundtrip)
    q.green_zero_degenerate = int(p.greenZeroDegenerate if zero_degenerate is None else zero_degenerate)
    q.timing = int(timing)
    return q


def plummer_case(n, grid=(32, 32, 32), box=(60.0, 60.0, 60.0), seed=42, **kw):
    """The reference's Plummer set-up (source/demos.cpp:1416-1447) on a smaller mesh."""
    center = tuple(b / 2 for b in box)
    pos, vel, mass = ics.plummer(n, center=center, a=2.0, r_max=min(box) / 4, M=1.0, G=4.5e-3, seed=seed)
    p = refapi.make_params(n, grid, box, **kw)
    return p, pos, vel, mass


def disk_case(n, grid=(32, 32, 16), box=(60.0, 60.0, 30.0), seed=42, **kw):
    """The reference's galaxy set-up (source/demos.cpp:736-776): linear disk + bulge field."""
    center = (box[0] / 2, box[1] / 2, box[2] / 2)
    pos, vel, mass = ics.disk_linear(n, center=center, seed=seed)
    kw.setdefault("ext", dict(center=center, R=3.0, M=60.0))
    p = refapi.make_params(n, grid, box, **kw)
    return p, pos, vel, mass


def uniform_case(n, grid=(32, 32, 32), box=(60.0, 60.0, 60.0), seed=7, margin=0.08, **kw):
    # margin keeps every TSC stencil inside the mesh arrays (base cell >= 1), the reference's domain
    lo = [margin * b for b in box]
    hi = [(1 - margin) * b for b in box]
    pos, vel, mass = ics.uniform_cube(n, lo, hi, total_mass=1.0, seed=seed)
    p = refapi.make_params(n, grid, box, **kw)
    return p, pos, vel, mass


def morton3(x, y, z):
    def spread(v):
        v = v.astype(np.uint64) & 0x3ff
        v = (v | (v << 16)) & 0x030000FF
        v = (v | (v << 8)) & 0x0300F00F
        v = (v | (v << 4)) & 0x030C30C3
        v = (v | (v << 2)) & 0x09249249
        return v
    return spread(x) | (spread(y) << 1) | (spread(z) << 2)


def degenerate_mask(shape):
    """True where every k_i is 0 or N_i/2."""
    nz, ny, nx = shape
    kz, ky, kx = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    return ((2 * kx) % nx == 0) & ((2 * ky) % ny == 0) & ((2 * kz) % nz == 0)


def project_out_degenerate(phi):
    """Remove the <= 8 modes on which the reference's optimal influence function is rounding noise."""
    f = np.fft.fftn(np.asarray(phi, np.float64))
    f[degenerate_mask(phi.shape)] = 0
    return np.fft.ifftn(f).real

"""ctypes binding of the C ABI in include/p3m_b200.h (libp3m_b200.so).

This is plumbing for tests and bench.py; the product is the shared library.  There is no CPU
fallback: if the library is missing or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# P3M_B200_LIB: A/B-testing hook for an alternatively tuned build of the same library
LIB_PATH = os.environ.get("P3M_B200_LIB") or os.path.join(_HERE, "libp3m_b200.so")

NGP, CIC, TSC = 0, 1, 2
TWO_POINT, FOUR_POINT = 0, 1
DISCRETE_LAPLACIAN, S1_OPTIMAL, S2_OPTIMAL, POOR_MAN = 0, 1, 2, 3
S1, S2 = 0, 1
EXT_NONE, EXT_SPH_RAD_DECR = 0, 1
F32, F64 = 0, 1
UNITS_ORIGINAL, UNITS_CODE = 0, 1
NPHASE = 10


class P3MParams(C.Structure):
    """struct p3m_params (include/p3m_b200.h)."""

    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("nz", C.c_int32),
        ("box", C.c_float * 3),
        ("H", C.c_float), ("DT", C.c_float), ("G", C.c_float),
        ("assignment", C.c_int32), ("fd_scheme", C.c_int32), ("greens_function", C.c_int32),
        ("particle_diameter", C.c_float),
        ("p3m", C.c_int32),
        ("cutoff_radius", C.c_float), ("softening", C.c_float),
        ("cloud_shape", C.c_int32), ("use_sr_table", C.c_int32),
        ("ext_kind", C.c_int32), ("ext_center", C.c_float * 3),
        ("ext_R", C.c_float), ("ext_M", C.c_float),
        ("precision", C.c_int32), ("unit_roundtrip", C.c_int32),
        ("green_zero_degenerate", C.c_int32), ("device", C.c_int32), ("timing", C.c_int32),
        ("sr_particle_diameter", C.c_float),
    ]


IC_PLUMMER, IC_DISK_LINEAR, IC_UNIFORM, IC_DISK_HALO = 0, 1, 2, 3
SUM_SHORT_RANGE, SUM_NEWTON, SUM_CUTOFF_SHELL = 0, 1, 2


class P3MIc(C.Structure):
    """struct p3m_ic (include/p3m_b200.h): device-side initial conditions."""

    _fields_ = [
        ("kind", C.c_int32), ("truncate", C.c_int32), ("seed", C.c_uint64), ("n", C.c_int64),
        ("center", C.c_float * 3), ("total_mass", C.c_float), ("G", C.c_float),
        ("a", C.c_float), ("r_max", C.c_float),
        ("rb", C.c_float), ("mb", C.c_float), ("rd", C.c_float), ("md", C.c_float), ("thickness", C.c_float),
        ("r0", C.c_float),
        ("lo", C.c_float * 3), ("hi", C.c_float * 3), ("vel_sigma", C.c_float),
    ]


def ic_plummer(n, center=(30.0, 30.0, 30.0), a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=42, truncate=False):
    ic = P3MIc(kind=IC_PLUMMER, truncate=int(truncate), seed=seed, n=n, total_mass=M, G=G, a=a, r_max=r_max)
    ic.center[:] = center
    return ic


def ic_disk_linear(n, center=(30.0, 30.0, 15.0), rb=3.0, mb=60.0, rd=15.0, md=15.0, thickness=0.3, G=4.5e-3, seed=42,
                   r0=0.0):
    ic = P3MIc(kind=IC_DISK_LINEAR, seed=seed, n=n, total_mass=md, G=G, rb=rb, mb=mb, rd=rd, md=md,
               thickness=thickness, r0=r0)
    ic.center[:] = center
    return ic


def ic_uniform(n, lo, hi, total_mass=1.0, vel_sigma=0.0, seed=42):
    ic = P3MIc(kind=IC_UNIFORM, seed=seed, n=n, total_mass=total_mass, vel_sigma=vel_sigma)
    ic.lo[:] = lo
    ic.hi[:] = hi
    return ic


def ic_disk_halo(n, box=60.0, halo_a=12.0, halo_rmax=27.0, disk_rd=24.0, thickness=3.0, M=1.0, G=4.5e-3, seed=42):
    ic = P3MIc(kind=IC_DISK_HALO, truncate=1, seed=seed, n=n, total_mass=M, G=G, a=halo_a, r_max=halo_rmax,
               rd=disk_rd, thickness=thickness)
    ic.center[:] = (box / 2,) * 3
    return ic


def sample_particles(ic, first=0, count=None):
    """Particles [first, first + count) of a device-generated set as host arrays (needs a GPU, no context)."""
    count = int(ic.n - first if count is None else count)
    pos = np.empty((count, 3), np.float32); vel = np.empty((count, 3), np.float32); mass = np.empty(count, np.float32)
    _check(lib().p3m_sample_particles(C.byref(ic), int(first), count, _p(pos), _p(vel), _p(mass)))
    return pos, vel, mass


class P3MError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"p3m error {code}: {msg}")
        self.code = code


_lib = None

# every symbol include/p3m_b200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "p3m_last_error", "p3m_version", "p3m_default_params", "p3m_create", "p3m_destroy",
    "p3m_comm_unique_id", "p3m_create_dist", "p3m_get_local", "p3m_set_particles_ids", "p3m_num_global",
    "p3m_rank_info", "p3m_slab_cuts", "p3m_balanced_cuts",
    "p3m_generate_particles", "p3m_sample_particles", "p3m_set_particles", "p3m_get_particles", "p3m_get_particles_f64", "p3m_num_particles",
    "p3m_green_init", "p3m_set_green_table", "p3m_set_green_table_f64", "p3m_get_green_table",
    "p3m_bin_sort", "p3m_deposit", "p3m_poisson", "p3m_gradient", "p3m_gather", "p3m_short_range",
    "p3m_force", "p3m_kick", "p3m_drift", "p3m_step", "p3m_escaped", "p3m_diagnostics",
    "p3m_add_acceleration", "p3m_fft3d_c2c",
    "p3m_get_density", "p3m_get_potential", "p3m_get_field", "p3m_get_density_f64",
    "p3m_get_potential_f64", "p3m_set_density", "p3m_set_potential", "p3m_get_cells",
    "p3m_get_chaining_dims", "p3m_get_binning", "p3m_chaining_neighbors", "p3m_get_acc_parts", "p3m_get_sr_table", "p3m_get_sample",
    "p3m_get_phase_ms", "p3m_phase_name", "p3m_get_pair_counts", "p3m_get_stats", "p3m_direct_sum", "p3m_launch_count", "p3m_stream",
    "p3m_synchronize",
]


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise FileNotFoundError(
                f"{LIB_PATH} not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                "(there is no CPU fallback)")
        L = C.CDLL(LIB_PATH)
        L.p3m_last_error.restype = C.c_char_p
        L.p3m_phase_name.restype = C.c_char_p
        L.p3m_num_particles.restype = C.c_int64
        L.p3m_num_global.restype = C.c_int64
        L.p3m_num_global.argtypes = [C.c_void_p]
        L.p3m_get_local.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.p3m_set_particles_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        L.p3m_rank_info.argtypes = [C.c_void_p, C.c_void_p]
        L.p3m_comm_unique_id.argtypes = [C.c_void_p]
        L.p3m_create_dist.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.p3m_launch_count.restype = C.c_int64
        L.p3m_stream.restype = C.c_void_p
        for name in ("p3m_num_particles", "p3m_launch_count", "p3m_stream", "p3m_destroy",
                     "p3m_synchronize", "p3m_green_init", "p3m_bin_sort", "p3m_deposit", "p3m_poisson",
                     "p3m_gradient", "p3m_gather", "p3m_short_range", "p3m_force", "p3m_drift"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.p3m_kick.argtypes = [C.c_void_p, C.c_float]
        L.p3m_step.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.p3m_set_particles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int]
        for name in ("p3m_get_particles", "p3m_get_particles_f64"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        for name in ("p3m_set_green_table", "p3m_set_green_table_f64", "p3m_get_green_table",
                     "p3m_get_density", "p3m_get_potential", "p3m_get_field", "p3m_get_density_f64",
                     "p3m_get_potential_f64", "p3m_set_density", "p3m_set_potential", "p3m_diagnostics",
                     "p3m_get_sr_table", "p3m_get_chaining_dims", "p3m_get_binning", "p3m_escaped"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        L.p3m_get_cells.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3m_add_acceleration.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.p3m_fft3d_c2c.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
        L.p3m_get_acc_parts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3m_get_pair_counts.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3m_get_phase_ms.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.p3m_get_stats.argtypes = [C.c_void_p, C.c_void_p]
        L.p3m_get_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3m_direct_sum.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int64, C.c_double, C.c_void_p]
        L.p3m_generate_particles.argtypes = [C.c_void_p, C.c_void_p]
        L.p3m_sample_particles.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        L.p3m_chaining_neighbors.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise P3MError(rc, lib().p3m_last_error().decode(errors="replace"))


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params() -> P3MParams:
    p = P3MParams()
    lib().p3m_default_params(C.byref(p))
    return p


def slab_cuts(params, nranks):
    """Layer cuts of the z-slab decomposition (host only)."""
    cuts = np.zeros(9, np.int32)
    layers = C.c_int32(0)
    _check(lib().p3m_slab_cuts(C.byref(params), int(nranks), _p(cuts), C.byref(layers)))
    return cuts[: nranks + 1].copy(), int(layers.value)


def balanced_cuts(params, nranks, pos, units=UNITS_ORIGINAL):
    """Work-balanced layer cuts for a full particle set (host only; what p3m_set_particles uses on N ranks)."""
    pos = np.ascontiguousarray(pos, np.float32)
    cuts = np.zeros(9, np.int32)
    layers = C.c_int32(0)
    _check(lib().p3m_balanced_cuts(C.byref(params), int(nranks), _p(pos), C.c_int64(len(pos)), int(units), _p(cuts),
                                   C.byref(layers)))
    return cuts[: nranks + 1].copy(), int(layers.value)


def comm_unique_id() -> bytes:
    """NCCL unique id (128 bytes) for p3m_create_dist; call on rank 0 and broadcast."""
    buf = C.create_string_buffer(128)
    _check(lib().p3m_comm_unique_id(buf))
    return buf.raw


def fft3d_c2c(x, inverse=False):
    """FFTAdapter contract on a (nz, ny, nx) complex64 array."""
    x = np.ascontiguousarray(x, np.complex64)
    out = np.empty_like(x)
    nz, ny, nx = x.shape
    _check(lib().p3m_fft3d_c2c(nz, ny, nx, _p(x), _p(out), int(inverse)))
    return out


def chaining_neighbors(dims, cell):
    dims = np.ascontiguousarray(dims, np.int32)
    out = np.empty(14, np.int32)
    _check(lib().p3m_chaining_neighbors(_p(dims), int(cell), _p(out)))
    return out


class Context:
    """One p3m_ctx.  Thin: every method is one C-ABI call on host numpy buffers."""

    def __init__(self, params: P3MParams, unique_id: bytes | None = None, rank: int = 0, nranks: int = 1):
        self._h = C.c_void_p()
        self.params = params
        self.rank, self.nranks = rank, nranks
        if nranks > 1:
            assert unique_id is not None and len(unique_id) == 128
            uid = C.create_string_buffer(unique_id, 128)
            _check(lib().p3m_create_dist(C.byref(params), uid, rank, nranks, C.byref(self._h)))
        else:
            _check(lib().p3m_create(C.byref(params), C.byref(self._h)))
        self.M = params.nx * params.ny * params.nz
        self.shape = (params.nz, params.ny, params.nx)

    def close(self):
        if self._h:
            lib().p3m_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def n(self):
        return int(lib().p3m_num_particles(self._h))

    @property
    def n_global(self):
        return int(lib().p3m_num_global(self._h))

    def rank_info(self):
        d = np.zeros(8, np.int64)
        _check(lib().p3m_rank_info(self._h, _p(d)))
        return dict(zip(("rank", "nranks", "layer0", "layer1", "ghosts", "slab", "plane0", "planes"),
                        (int(v) for v in d)))

    def get_local(self, units=UNITS_ORIGINAL, want=("pos", "vel"), out=None):
        """This rank's particles with their global ids.  `out` = (ids, pos, vel) preallocated (e.g. pinned)
        arrays of at least self.n rows: filled in place, views of the first n rows returned."""
        n = self.n
        if out is not None:
            ids, pos, vel = out
            assert len(ids) >= n and len(pos) >= n and len(vel) >= n
            _check(lib().p3m_get_local(self._h, _p(ids), _p(pos), _p(vel), None, units))
            return ids[:n], pos[:n], vel[:n], None
        ids = np.empty(n, np.int32)
        out = {k: (np.empty((n, 3), np.float32) if k in want else None) for k in ("pos", "vel", "acc")}
        _check(lib().p3m_get_local(self._h, _p(ids), _p(out["pos"]), _p(out["vel"]), _p(out["acc"]), units))
        return ids, out["pos"], out["vel"], out["acc"]

    def set_particles_ids(self, pos, vel, mass, ids, units=UNITS_ORIGINAL):
        pos = np.ascontiguousarray(pos, np.float32); vel = np.ascontiguousarray(vel, np.float32)
        mass = np.ascontiguousarray(mass, np.float32); ids = np.ascontiguousarray(ids, np.int32)
        _check(lib().p3m_set_particles_ids(self._h, _p(pos), _p(vel), _p(mass), _p(ids), len(ids), units))

    def generate_particles(self, ic):
        """Device-side initial conditions (p3m_generate_particles): every rank keeps its own z-slab."""
        _check(lib().p3m_generate_particles(self._h, C.byref(ic)))

    def set_particles(self, pos, vel, mass, units=UNITS_ORIGINAL):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        n = mass.shape[0]
        assert pos.size == 3 * n and (vel is None or vel.size == 3 * n)
        _check(lib().p3m_set_particles(self._h, _p(pos), _p(vel), _p(mass), n, units))

    def get_particles(self, units=UNITS_CODE, f64=False, want=("pos", "vel", "acc")):
        n = self.n_global
        dt = np.float64 if f64 else np.float32
        out = {k: (np.empty((n, 3), dt) if k in want else None) for k in ("pos", "vel", "acc")}
        fn = lib().p3m_get_particles_f64 if f64 else lib().p3m_get_particles
        _check(fn(self._h, _p(out["pos"]), _p(out["vel"]), _p(out["acc"]), units))
        return out["pos"], out["vel"], out["acc"]

    def green_init(self): _check(lib().p3m_green_init(self._h))

    def set_green_table(self, table):
        table = np.ascontiguousarray(table)
        if table.dtype == np.float64:
            _check(lib().p3m_set_green_table_f64(self._h, _p(table)))
        else:
            _check(lib().p3m_set_green_table(self._h, _p(np.ascontiguousarray(table, np.float32))))

    def get_green_table(self):
        g = np.empty(self.M, np.float64)
        _check(lib().p3m_get_green_table(self._h, _p(g)))
        return g.reshape(self.shape)

    def bin_sort(self): _check(lib().p3m_bin_sort(self._h))
    def deposit(self): _check(lib().p3m_deposit(self._h))
    def poisson(self): _check(lib().p3m_poisson(self._h))
    def gradient(self): _check(lib().p3m_gradient(self._h))
    def gather(self): _check(lib().p3m_gather(self._h))
    def short_range(self): _check(lib().p3m_short_range(self._h))
    def force(self): _check(lib().p3m_force(self._h))
    def kick(self, f=1.0): _check(lib().p3m_kick(self._h, float(f)))
    def drift(self): _check(lib().p3m_drift(self._h))
    def synchronize(self): _check(lib().p3m_synchronize(self._h))

    def step(self, steps=1):
        done = C.c_int(0)
        _check(lib().p3m_step(self._h, int(steps), C.byref(done)))
        return done.value

    def escaped(self):
        e = C.c_int(0)
        _check(lib().p3m_escaped(self._h, C.byref(e)))
        return bool(e.value)

    def add_acceleration(self, acc, units=UNITS_ORIGINAL):
        _check(lib().p3m_add_acceleration(self._h, _p(np.ascontiguousarray(acc, np.float32)), units))

    def diagnostics(self):
        d = np.zeros(11, np.float64)
        _check(lib().p3m_diagnostics(self._h, _p(d)))
        return d

    def _mesh(self, fn, dt, comps=1):
        a = np.empty(self.M * comps, dt)
        _check(fn(self._h, _p(a)))
        return a.reshape(self.shape + ((comps,) if comps > 1 else ()))

    def density(self, f64=False):
        return self._mesh(lib().p3m_get_density_f64 if f64 else lib().p3m_get_density,
                          np.float64 if f64 else np.float32)

    def potential(self, f64=False):
        return self._mesh(lib().p3m_get_potential_f64 if f64 else lib().p3m_get_potential,
                          np.float64 if f64 else np.float32)

    def field(self):
        return self._mesh(lib().p3m_get_field, np.float32, 3)

    def set_density(self, rho):
        _check(lib().p3m_set_density(self._h, _p(np.ascontiguousarray(rho, np.float32))))

    def set_potential(self, phi):
        _check(lib().p3m_set_potential(self._h, _p(np.ascontiguousarray(phi, np.float32))))

    def cells(self):
        n = self.n_global
        mc = np.empty(n, np.int32); cc = np.empty(n, np.int32); order = np.empty(self.n, np.int32)
        _check(lib().p3m_get_cells(self._h, _p(mc), _p(cc), _p(order)))
        return mc, cc, order

    def chaining_dims(self):
        d = np.zeros(3, np.int32)
        _check(lib().p3m_get_chaining_dims(self._h, _p(d)))
        return d

    def binning(self):
        d = np.zeros(8, np.int32)
        _check(lib().p3m_get_binning(self._h, _p(d)))
        return dict(zip(("mx", "my", "mz", "mbits", "sbits", "idbits", "bshift", "p3m"), (int(v) for v in d)))

    def acc_parts(self):
        n = self.n_global
        pm = np.empty((n, 3), np.float64); sr = np.empty((n, 3), np.float64)
        _check(lib().p3m_get_acc_parts(self._h, _p(pm), _p(sr)))
        return pm, sr

    def sample(self, ids):
        """(pos, acc, acc_sr) rows, code units, of the particles with the given ascending global ids (zeros where
        this rank does not hold the particle)."""
        ids = np.ascontiguousarray(ids, np.int32)
        out = [np.zeros((len(ids), 3), np.float64) for _ in range(3)]
        _check(lib().p3m_get_sample(self._h, _p(ids), len(ids), _p(out[0]), _p(out[1]), _p(out[2])))
        return out

    def sr_table(self):
        t = np.empty(500, np.float64)
        _check(lib().p3m_get_sr_table(self._h, _p(t)))
        return t

    def phase_ms(self, reset=False):
        ms = np.zeros(NPHASE, np.float32)
        _check(lib().p3m_get_phase_ms(self._h, _p(ms), int(reset)))
        return {lib().p3m_phase_name(i).decode(): float(ms[i]) for i in range(NPHASE)}

    def pair_counts(self):
        a = C.c_uint64(0); b = C.c_uint64(0)
        _check(lib().p3m_get_pair_counts(self._h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def direct_sum(self, target_pos_code, mode=0, softening_code=0.0):
        """N4: fp64 brute-force sum over this rank's particles for the given target points (code units)."""
        t = np.ascontiguousarray(target_pos_code, np.float64).reshape(-1, 3)
        out = np.zeros_like(t)
        _check(lib().p3m_direct_sum(self._h, int(mode), _p(t), len(t), float(softening_code), _p(out)))
        return out

    STAT_NAMES = ("fused_z", "slab", "uniform_mass_table", "packed_pp", "incremental_sort", "migrated", "ghosts",
                  "a2a_bytes", "density_plane_bytes", "potential_plane_bytes", "migration_bytes", "ghost_bytes",
                  "sort_movers", "full_sorts", "incremental_sorts", "reserved")

    def stats(self):
        d = np.zeros(16, np.float64)
        _check(lib().p3m_get_stats(self._h, _p(d)))
        return dict(zip(self.STAT_NAMES, (float(v) for v in d)))

    @property
    def fused_z(self):
        return bool(self.stats()["fused_z"])

    @property
    def launches(self):
        return int(lib().p3m_launch_count(self._h))

    @property
    def stream(self):
        return lib().p3m_stream(self._h)

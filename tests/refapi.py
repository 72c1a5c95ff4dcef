"""ctypes bindings for the two CPU checkers (TEST INFRASTRUCTURE):

* ``oracle/libp3m_oracle.so`` -- the plain-C restatement (oracle/p3m_oracle.c), fp32 + fp64;
* ``oracle/_ref/libp3m_ref.so`` -- the UNMODIFIED reference compiled by oracle/Makefile
  (oracle/ref_driver.cpp); present wherever it was built in the container and shipped.

Nothing here is imported by the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libp3m_oracle.so")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libp3m_ref.so")

NGP, CIC, TSC = 0, 1, 2
TWO_POINT, FOUR_POINT = 0, 1
DISCRETE_LAPLACIAN, S1_OPTIMAL, S2_OPTIMAL, POOR_MAN = 0, 1, 2, 3
S1, S2 = 0, 1


class Params(C.Structure):
    """Layout of OrcParams (oracle/p3m_oracle.h) == RefParams (oracle/ref_driver.cpp)."""

    _fields_ = [
        ("n", C.c_int),
        ("nx", C.c_int), ("ny", C.c_int), ("nz", C.c_int),
        ("box", C.c_float * 3),
        ("H", C.c_float), ("DT", C.c_float), ("G", C.c_float),
        ("is_", C.c_int), ("fds", C.c_int), ("gfunc", C.c_int),
        ("particleDiameter", C.c_float),
        ("cutoffRadius", C.c_float),
        ("softening", C.c_float),
        ("cloudShape", C.c_int),
        ("useTable", C.c_int),
        ("ySort", C.c_int),
        ("extKind", C.c_int),
        ("extCenter", C.c_float * 3),
        ("extR", C.c_float), ("extM", C.c_float),
        ("greenZeroDegenerate", C.c_int),
    ]

    @property
    def M(self):
        return self.nx * self.ny * self.nz


def make_params(n, grid, box, *, H=None, DT=1.0, G=4.5e-3, is_=TSC, fds=TWO_POINT,
                gfunc=S1_OPTIMAL, diameter=None, cutoff=None, softening=0.5, cloud=S1,
                use_table=True, y_sort=True, ext=None, zero_degenerate=False) -> Params:
    """Defaults follow the reference demos: H = box_x / (Nx / 2) (source/demos.cpp:757-761),
    particle diameter 3H, cutoff 0.7 * diameter (source/demos.cpp:1416-1447)."""
    p = Params()
    p.n = int(n)
    p.nx, p.ny, p.nz = (int(g) for g in grid)
    p.box[:] = [np.float32(b) for b in box]
    f32 = np.float32
    p.H = f32(H) if H is not None else f32(f32(box[0]) / f32(p.nx // 2))
    p.DT, p.G = f32(DT), f32(G)
    p.is_, p.fds, p.gfunc = is_, fds, gfunc
    p.particleDiameter = f32(diameter) if diameter is not None else f32(f32(3) * f32(p.H))
    p.cutoffRadius = f32(cutoff) if cutoff is not None else f32(f32(0.7) * f32(p.particleDiameter))
    p.softening = f32(softening)
    p.cloudShape = cloud
    p.useTable, p.ySort = int(use_table), int(y_sort)
    p.greenZeroDegenerate = int(zero_degenerate)
    if ext is None:
        p.extKind = 0
    else:
        p.extKind = 1
        p.extCenter[:] = [f32(c) for c in ext["center"]]
        p.extR, p.extM = f32(ext["R"]), f32(ext["M"])
    return p


def _ptr(a, ct):
    if a is None:
        return None
    return a.ctypes.data_as(C.POINTER(ct))


def build_oracle(force=False):
    """Compile the C restatement (gcc only).  Building the checker is not using it."""
    src = [os.path.join(ORACLE_DIR, f) for f in ("p3m_oracle.c", "p3m_oracle_impl.h", "p3m_oracle.h")]
    if (not force and os.path.exists(ORACLE_SO)
            and os.path.getmtime(ORACLE_SO) >= max(os.path.getmtime(s) for s in src)):
        return ORACLE_SO
    subprocess.check_call(["make", "-C", ORACLE_DIR, "libp3m_oracle.so"], stdout=subprocess.DEVNULL)
    return ORACLE_SO


class Oracle:
    """The C restatement.  ``prec`` is 'f32' (reference precision) or 'f64'."""

    def __init__(self, prec="f32"):
        build_oracle()
        self.lib = C.CDLL(ORACLE_SO)
        self.prec = prec
        self.dt = np.float32 if prec == "f32" else np.float64
        self.ct = C.c_float if prec == "f32" else C.c_double

    def _f(self, name):
        return getattr(self.lib, f"{name}_{self.prec}")

    def _r(self, a):
        return _ptr(None if a is None else np.ascontiguousarray(a, self.dt), self.ct)

    def to_code_units(self, p, pos, vel, mass):
        pos = np.ascontiguousarray(pos, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        pc = np.empty((p.n, 3), self.dt); vc = np.empty((p.n, 3), self.dt); mc = np.empty(p.n, self.dt)
        self._f("orc_to_code_units")(C.byref(p), _ptr(pos, C.c_float), _ptr(vel, C.c_float),
                                     _ptr(mass, C.c_float), _ptr(pc, self.ct), _ptr(vc, self.ct),
                                     _ptr(mc, self.ct))
        return pc, vc, mc

    def green(self, p):
        g = np.empty(p.M, self.dt)
        self._f("orc_green")(C.byref(p), _ptr(g, self.ct))
        return g.reshape(p.nz, p.ny, p.nx)

    def deposit(self, p, pos_c, mass_c):
        pos_c = np.ascontiguousarray(pos_c, self.dt); mass_c = np.ascontiguousarray(mass_c, self.dt)
        d = np.empty(p.M, self.dt)
        self._f("orc_deposit")(C.byref(p), _ptr(pos_c, self.ct), _ptr(mass_c, self.ct), _ptr(d, self.ct))
        return d.reshape(p.nz, p.ny, p.nx)

    def poisson(self, p, density, green):
        density = np.ascontiguousarray(density, self.dt); green = np.ascontiguousarray(green, self.dt)
        phi = np.empty(p.M, self.dt)
        self._f("orc_poisson")(C.byref(p), _ptr(density, self.ct), _ptr(green, self.ct), _ptr(phi, self.ct))
        return phi.reshape(p.nz, p.ny, p.nx)

    def field(self, p, phi):
        phi = np.ascontiguousarray(phi, self.dt)
        f = np.empty((p.M, 3), self.dt)
        self._f("orc_field")(C.byref(p), _ptr(phi, self.ct), _ptr(f, self.ct))
        return f.reshape(p.nz, p.ny, p.nx, 3)

    def gather(self, p, pos_c, field):
        pos_c = np.ascontiguousarray(pos_c, self.dt); field = np.ascontiguousarray(field, self.dt)
        a = np.empty((p.n, 3), self.dt)
        self._f("orc_gather")(C.byref(p), _ptr(pos_c, self.ct), _ptr(field, self.ct), _ptr(a, self.ct))
        return a

    def sr_table(self, p):
        t = np.empty(500, self.dt)
        self._f("orc_sr_table")(C.byref(p), _ptr(t, self.ct))
        return t

    def chaining_cells(self, p, pos_c):
        pos_c = np.ascontiguousarray(pos_c, self.dt)
        dims = np.zeros(3, np.int32); cell = np.empty(p.n, np.int32)
        self._f("orc_chaining_cells")(C.byref(p), _ptr(pos_c, self.ct), _ptr(dims, C.c_int), _ptr(cell, C.c_int))
        return dims, cell

    def chaining_order(self, p, pos_c):
        pos_c = np.ascontiguousarray(pos_c, self.dt)
        order = np.full(p.n, -1, np.int32)
        self._f("orc_chaining_order")(C.byref(p), _ptr(pos_c, self.ct), _ptr(order, C.c_int))
        return order

    def chaining_neighbors(self, dims, cell):
        dims = np.ascontiguousarray(dims, np.int32); out = np.empty(14, np.int32)
        self.lib.orc_chaining_neighbors(_ptr(dims, C.c_int), C.c_int(int(cell)), _ptr(out, C.c_int))
        return out

    def sr_forces(self, p, pos_c, mass_c):
        pos_c = np.ascontiguousarray(pos_c, self.dt); mass_c = np.ascontiguousarray(mass_c, self.dt)
        sr = np.empty((p.n, 3), self.dt)
        self._f("orc_sr_forces")(C.byref(p), _ptr(pos_c, self.ct), _ptr(mass_c, self.ct), _ptr(sr, self.ct))
        return sr

    def force(self, p, p3m, green, pos_c, mass_c):
        pos_c = np.ascontiguousarray(pos_c, self.dt); mass_c = np.ascontiguousarray(mass_c, self.dt)
        green = np.ascontiguousarray(green, self.dt)
        rho = np.empty(p.M, self.dt); phi = np.empty(p.M, self.dt); acc = np.empty((p.n, 3), self.dt)
        self._f("orc_force")(C.byref(p), C.c_int(int(p3m)), _ptr(green, self.ct), _ptr(pos_c, self.ct),
                             _ptr(mass_c, self.ct), _ptr(rho, self.ct), _ptr(phi, self.ct), _ptr(acc, self.ct))
        return rho.reshape(p.nz, p.ny, p.nx), phi.reshape(p.nz, p.ny, p.nx), acc

    def run(self, p, p3m, pos, vel, mass, sim_length, diagnostics=True):
        pos = np.ascontiguousarray(pos, np.float32); vel = np.ascontiguousarray(vel, np.float32)
        mass = np.ascontiguousarray(mass, np.float32)
        diag = np.zeros((sim_length + 1, 12), self.dt) if diagnostics else None
        po = np.empty((p.n, 3), self.dt); vo = np.empty((p.n, 3), self.dt); ao = np.empty((p.n, 3), self.dt)
        fn = self._f("orc_run"); fn.restype = C.c_int
        rows = fn(C.byref(p), C.c_int(int(p3m)), _ptr(pos, C.c_float), _ptr(vel, C.c_float),
                  _ptr(mass, C.c_float), C.c_int(sim_length), _ptr(diag, self.ct), _ptr(po, self.ct),
                  _ptr(vo, self.ct), _ptr(ao, self.ct))
        return (diag[:rows] if diagnostics else None), po, vo, ao


def have_ref():
    return os.path.exists(REF_SO)


class Ref:
    """The unmodified reference (fp32 only), via oracle/ref_driver.cpp."""

    def __init__(self):
        if not have_ref():
            raise FileNotFoundError(REF_SO + " (build with: make -C oracle ref; needs /root/reference)")
        self.lib = C.CDLL(REF_SO)

    @staticmethod
    def _in(pos, vel, mass):
        pos = np.ascontiguousarray(pos, np.float32); mass = np.ascontiguousarray(mass, np.float32)
        vel = None if vel is None else np.ascontiguousarray(vel, np.float32)
        return pos, vel, mass

    def green(self, p):
        g = np.empty((p.M, 2), np.float32)
        self.lib.ref_green(C.byref(p), _ptr(g, C.c_float))
        return g.reshape(p.nz, p.ny, p.nx, 2)

    def pm_force(self, p, pos, vel, mass, want_green=False):
        pos, vel, mass = self._in(pos, vel, mass)
        out = dict(pos_code=np.empty((p.n, 3), np.float32), mass_code=np.empty(p.n, np.float32),
                   density=np.empty(p.M, np.float32), potential=np.empty(p.M, np.float32),
                   field=np.empty((p.M, 3), np.float32), acc=np.empty((p.n, 3), np.float32),
                   green=np.empty((p.M, 2), np.float32) if want_green else None)
        self.lib.ref_pm_force(C.byref(p), _ptr(pos, C.c_float), _ptr(vel, C.c_float), _ptr(mass, C.c_float),
                              *[_ptr(out[k], C.c_float) for k in
                                ("pos_code", "mass_code", "density", "potential", "field", "acc", "green")])
        for k in ("density", "potential"):
            out[k] = out[k].reshape(p.nz, p.ny, p.nx)
        out["field"] = out["field"].reshape(p.nz, p.ny, p.nx, 3)
        if want_green:
            out["green"] = out["green"].reshape(p.nz, p.ny, p.nx, 2)
        return out

    def p3m_force(self, p, pos, vel, mass):
        pos, vel, mass = self._in(pos, vel, mass)
        out = dict(acc_pm=np.empty((p.n, 3), np.float32), sr_force=np.empty((p.n, 3), np.float32),
                   acc=np.empty((p.n, 3), np.float32), cell=np.full(p.n, -1, np.int32),
                   order=np.full(p.n, -1, np.int32), dims=np.zeros(3, np.int32),
                   ftable=np.zeros(500, np.float32))
        self.lib.ref_p3m_force(C.byref(p), _ptr(pos, C.c_float), _ptr(vel, C.c_float), _ptr(mass, C.c_float),
                               _ptr(out["acc_pm"], C.c_float), _ptr(out["sr_force"], C.c_float),
                               _ptr(out["acc"], C.c_float), _ptr(out["cell"], C.c_int),
                               _ptr(out["order"], C.c_int), _ptr(out["dims"], C.c_int),
                               _ptr(out["ftable"], C.c_float))
        return out

    def chaining_cells(self, p, pos_code):
        pos_code = np.ascontiguousarray(pos_code, np.float32)
        cell = np.empty(p.n, np.int32); dims = np.zeros(3, np.int32)
        self.lib.ref_chaining_cells(C.byref(p), _ptr(pos_code, C.c_float), _ptr(cell, C.c_int), _ptr(dims, C.c_int))
        return dims, cell

    def chaining_neighbors(self, p, cell):
        out = np.empty(14, np.int32)
        self.lib.ref_chaining_neighbors(C.byref(p), C.c_int(int(cell)), _ptr(out, C.c_int))
        return out

    def run(self, p, pos, vel, mass, sim_length, p3m, out_dir, diagnostics=True):
        pos, vel, mass = self._in(pos, vel, mass)
        po = np.empty((p.n, 3), np.float32); vo = np.empty((p.n, 3), np.float32); ao = np.empty((p.n, 3), np.float32)
        self.lib.ref_run(C.byref(p), _ptr(pos, C.c_float), _ptr(vel, C.c_float), _ptr(mass, C.c_float),
                         C.c_int(sim_length), C.c_int(int(p3m)), C.c_int(int(diagnostics)),
                         C.c_char_p(str(out_dir).encode()), _ptr(po, C.c_float), _ptr(vo, C.c_float),
                         _ptr(ao, C.c_float))
        diag = None
        if diagnostics:
            e = np.loadtxt(os.path.join(out_dir, "energy.txt"), ndmin=2)
            m = np.loadtxt(os.path.join(out_dir, "momentum.txt"), ndmin=2)
            L = np.loadtxt(os.path.join(out_dir, "angular_momentum.txt"), ndmin=2)
            x = np.loadtxt(os.path.join(out_dir, "expected_momentum.txt"), ndmin=2)
            diag = np.concatenate([e, m, L, x], axis=1)
        return diag, po, vo, ao

    def time_steps(self, p, pos, vel, mass, steps, p3m):
        pos, vel, mass = self._in(pos, vel, mass)
        ms = np.zeros(11, np.float32); gi = C.c_float(0)
        self.lib.ref_time_steps(C.byref(p), _ptr(pos, C.c_float), _ptr(vel, C.c_float), _ptr(mass, C.c_float),
                                C.c_int(steps), C.c_int(int(p3m)), _ptr(ms, C.c_float), C.byref(gi))
        names = ["spreadMass", "forwardFFT", "fourierPotential", "inverseFFT", "fieldInCells",
                 "updateAccelerations", "chainingMeshSetup", "shortRangeForcesCalc",
                 "correctAccelerations", "integrate", "total"]
        return dict(zip(names, (float(v) for v in ms))), float(gi.value)

    def sample_plummer(self, seed, center, a, r_max, M, G, n):
        c = (C.c_float * 3)(*center)
        pos = np.empty((n, 3), np.float32); vel = np.empty((n, 3), np.float32)
        self.lib.ref_sample_plummer(C.c_uint(seed), c, C.c_float(a), C.c_float(r_max), C.c_float(M),
                                    C.c_float(G), C.c_int(n), _ptr(pos, C.c_float), _ptr(vel, C.c_float))
        return pos, vel

    def sample_disk_linear(self, seed, center, rb, mb, rd, md, thickness, G, n):
        c = (C.c_float * 3)(*center)
        pos = np.empty((n, 3), np.float32); vel = np.empty((n, 3), np.float32)
        self.lib.ref_sample_disk_linear(C.c_uint(seed), c, C.c_float(rb), C.c_float(mb), C.c_float(rd),
                                        C.c_float(md), C.c_float(thickness), C.c_float(G), C.c_int(n),
                                        _ptr(pos, C.c_float), _ptr(vel, C.c_float))
        return pos, vel

    def hardware_threads(self):
        return int(self.lib.ref_hardware_threads())


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))

"""Host-side classes of the reference API (particlesimulation_b200/host): ChainingMesh, the single-mode
influence functions, leapfrog free functions + LeapfrogStepper, unit conversions, SimInfo -- on the CPU --
and, with a GPU, Grid back-fill through `getGrid() const`, the PMMethodGPU extras and P3MMethod's argument
handling.  `host/host_units` only dumps what the classes return; everything is compared here with the
UNMODIFIED reference (oracle/_ref) or with a numpy restatement of the reference lines cited.

Also: the reference-tree build (host/Makefile target `reftree`): the body of the reference's own
galaxySimulationP3M compiled against the drop-in headers placed before the reference's include/."""
import os
import struct
import subprocess

import numpy as np
import pytest

import refapi
from common import degenerate_mask
from refapi import rel_l2

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HOST = os.path.join(ROOT, "particlesimulation_b200", "host")
UNITS = os.path.join(HOST, "host_units")
REFTREE = os.path.join(HOST, "reftree_build", "demo_reftree")
REFERENCE = "/root/reference"

needs_ref = pytest.mark.skipif(not refapi.have_ref(), reason="oracle/_ref not built")


def write_input(path, pos, vel, mass, box, cutoff, H, DT, G):
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(mass)))
        f.write(np.asarray(list(box) + [cutoff, H, DT, G], np.float32).tobytes())
        f.write(np.ascontiguousarray(pos, np.float32).tobytes())
        f.write(np.ascontiguousarray(vel, np.float32).tobytes())
        f.write(np.ascontiguousarray(mass, np.float32).tobytes())


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            head = f.read(32)
            if len(head) < 32:
                break
            name = head.split(b"\0", 1)[0].decode()
            dtype = f.read(1).decode()
            (count,) = struct.unpack("q", f.read(8))
            out[name] = np.frombuffer(f.read(4 * count), np.float32 if dtype == "f" else np.int32).copy()
    return out


def case(n=2000, seed=3, cubic=True):
    # (the reference's own short-range loop does not survive every non-cubic chaining mesh: the list-order
    # comparison, which runs it, uses a cubic box; the mesh-only GPU case uses the galaxy demo's flat box)
    rng = np.random.default_rng(seed)
    box = (60.0, 60.0, 60.0) if cubic else (60.0, 60.0, 30.0)
    pos = (np.array(box) * (0.1 + 0.8 * rng.random((n, 3)))).astype(np.float32)
    vel = (0.05 * rng.standard_normal((n, 3))).astype(np.float32)
    mass = rng.uniform(0.5, 1.5, n).astype(np.float32)
    p = refapi.make_params(n, (32, 32, 32 if cubic else 16), box, gfunc=refapi.DISCRETE_LAPLACIAN)
    return p, pos, vel, mass, box


def run_units(mode, tmp_path, p, pos, vel, mass, box):
    if not os.path.exists(UNITS):
        pytest.skip("host/host_units not built")
    inp, out = tmp_path / "in.bin", tmp_path / f"{mode}.bin"
    write_input(inp, pos, vel, mass, box, float(p.cutoffRadius), float(p.H), float(p.DT), float(p.G))
    r = subprocess.run([UNITS, mode, str(inp), str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    return read_dump(out)


# ------------------------------------------------------------------------------------------------ CPU
@needs_ref
def test_host_chaining_mesh_matches_reference(tmp_path):
    """ChainingMesh::fill / fillWithYSorting / getNeighborsAndSelf (source/chainingMesh.cpp:20-84): same cells,
    same LIST ORDER as the reference's insertion, same 14-entry neighbour arrays."""
    p, pos, vel, mass, box = case()
    d = run_units("cpu", tmp_path, p, pos, vel, mass, box)
    ref = refapi.Ref()
    f32 = np.float32
    pc = (pos / f32(p.H)).astype(np.float32)
    assert np.array_equal(d["pos_code"].reshape(-1, 3), pc)
    dims, cell = ref.chaining_cells(p, pc)
    assert np.array_equal(d["cm_dims"][:3], dims) and d["cm_dims"][3] == dims.prod()
    assert np.array_equal(d["cm_cell"], cell) and np.array_equal(d["cm_cell_ysort"], cell)
    for ysort, key in ((1, "cm_order_ysort"), (0, "cm_order")):
        p.ySort = ysort
        r = ref.p3m_force(p, pos, vel, mass)
        assert np.array_equal(d[key], r["order"]), key
    nb = d["cm_neighbors"].reshape(-1, 14)
    for c in list(range(0, nb.shape[0], 37)) + [nb.shape[0] - 1]:
        assert np.array_equal(nb[c], ref.chaining_neighbors(p, c))


@needs_ref
def test_host_single_mode_green_functions_match_reference(tmp_path):
    """GreenDiscreteLaplacian / GreenPoorMan / GreenOptimal (source/greensFunctions.cpp:122-220) on an
    8 x 6 x 4 mesh against the table the reference's initGreensFunction fills (fp32; ours evaluates in double)."""
    p, pos, vel, mass, box = case(50)
    d = run_units("cpu", tmp_path, p, pos, vel, mass, box)
    ref = refapi.Ref()
    grid = (8, 6, 4)

    def table(**kw):
        q = refapi.make_params(1, grid, (8.0, 6.0, 4.0), H=1.0, **kw)
        return ref.green(q)[..., 0].ravel()

    assert rel_l2(d["green_laplacian"], table(gfunc=refapi.DISCRETE_LAPLACIAN)) < 2e-6
    assert rel_l2(d["green_poorman"], table(gfunc=refapi.POOR_MAN)) < 2e-6
    m = ~degenerate_mask((4, 6, 8)).ravel()  # GreenOptimal is 0/0 rounding noise there
    g1 = table(gfunc=refapi.S1_OPTIMAL, is_=refapi.TSC, fds=refapi.TWO_POINT, diameter=3.0)
    g2 = table(gfunc=refapi.S2_OPTIMAL, is_=refapi.CIC, fds=refapi.FOUR_POINT, diameter=2.5)
    assert rel_l2(d["green_s1_tsc_2pt"][m], g1[m]) < 3e-5
    assert rel_l2(d["green_s2_cic_4pt"][m], g2[m]) < 3e-5


def test_host_leapfrog_units_and_siminfo(tmp_path):
    """source/leapfrog.cpp:5-24, include/unitConversions.h:8-50, source/simInfo.cpp:4-48,94-127 restated in
    numpy float32 with the same operation order."""
    p, pos, vel, mass, box = case(500)
    d = run_units("cpu", tmp_path, p, pos, vel, mass, box)
    f32 = np.float32
    n = len(mass)
    i = np.arange(n)
    acc = np.stack([f32(0.01) * (i % 7).astype(f32), f32(-0.02) * (i % 5).astype(f32),
                    f32(0.005) * (i % 3).astype(f32)], axis=1).astype(f32)
    v = vel + f32(0.5) * f32(1.0) * acc
    assert np.allclose(d["lf_half_vel"].reshape(-1, 3), v, rtol=1e-6, atol=1e-9)
    x = pos + f32(1.0) * v
    assert np.allclose(d["lf_pos"].reshape(-1, 3), x, rtol=1e-6)
    v2 = v + f32(0.5) * acc
    assert np.allclose(d["lf_vel"].reshape(-1, 3), v2, rtol=1e-6, atol=1e-9)
    assert np.allclose(d["lf_int_vel"].reshape(-1, 3), v2 + f32(0.5) * acc, rtol=1e-6, atol=1e-9)
    # LeapfrogStepper: drift, force, kick (twice, dt = 1 then 0.5) with a = -0.001 x
    xs, vs = pos.astype(np.float64), vel.astype(np.float64)
    for dt in (1.0, 0.5):
        xs = xs + dt * vs
        vs = vs + dt * (-0.001 * xs)
    assert np.allclose(d["stepper_pos"].reshape(-1, 3), xs, rtol=1e-5)
    assert np.allclose(d["stepper_vel"].reshape(-1, 3), vs, rtol=1e-4, atol=1e-7)
    # units
    H, DT, G = f32(p.H), f32(p.DT), f32(p.G)
    pi = f32(np.pi)
    mf = DT * DT * f32(4) * pi * G / (H * H * H)
    assert np.array_equal(d["mass_code"], (mf * mass).astype(f32))
    st = np.concatenate([pos, vel])
    assert np.allclose(d["state_roundtrip"].reshape(-1, 3), st, rtol=3e-7, atol=1e-12)
    two = f32(2)
    expect = [DT * DT * f32(4) * pi * G * two, two / (DT * DT * f32(4) * pi * G), two * H * H / (DT * DT), two / H,
              mf * two, (H * H * H) / (DT * DT * f32(4) * pi * G) * two]
    assert np.allclose(d["unit_scalars"], np.array(expect, f32), rtol=1e-6)
    # SimInfo
    a0 = np.array([0.01, 0.02, -0.01])
    vi = vel.astype(np.float64) + 0.5 * a0
    m64, x64 = mass.astype(np.float64), pos.astype(np.float64)
    ke = 0.5 * (m64 * (vel.astype(np.float64) ** 2).sum(1)).sum()
    mom = (m64[:, None] * vi).sum(0)
    L = (m64[:, None] * np.cross(x64, vi)).sum(0)
    k = 200
    sub = x64[:k]
    pe = 0.0
    for a in range(k):
        r = np.sqrt(((sub[a] - sub[a + 1:]) ** 2).sum(1) + 1e-4)
        pe -= float(p.G) * m64[a] * (m64[a + 1:k] / r).sum()
    ke2 = 0.5 * (m64[:k] * (vel[:k].astype(np.float64) ** 2).sum(1)).sum()
    mom2 = (m64[:k, None] * vel[:k]).sum(0)
    em = mom + 0.5 * np.array([1.0, 2.0, 3.0])
    expect = np.concatenate([[ke], mom, L, [pe, ke2], mom2, em])
    assert np.allclose(d["siminfo"], expect, rtol=2e-4, atol=1e-3 * np.abs(expect).max())


@pytest.mark.skipif(not os.path.isdir(os.path.join(REFERENCE, "source")), reason="needs the reference checkout")
def test_drop_in_headers_compile_against_the_reference_tree(tmp_path):
    """A TU that includes the reference's OWN sampler headers next to the drop-in pmMethod.h / p3mMethod.h
    (-I host/include -I reference/include) and holds the body of the reference's galaxySimulationP3M
    (source/demos.cpp, read from the checkout at build time; only the FFT-adapter line changed) compiles, and
    links with the reference's retained sources + host/src/*.cpp: `make -C host reftree`."""
    r = subprocess.run(["make", "-C", HOST, "-B", "reftree"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert os.path.exists(REFTREE)
    gen = open(os.path.join(HOST, "reftree_build", "galaxy_p3m_caller.cpp")).read()
    ref_lines = open(os.path.join(REFERENCE, "source", "demos.cpp")).read().splitlines()
    start = next(i for i, l in enumerate(ref_lines) if l.startswith("void galaxySimulationP3M(const char* outputDir)"))
    end = next(i for i, l in enumerate(ref_lines) if l.startswith("void galaxySimulationP3MTiming"))
    body = ref_lines[start:end]
    changed = [l for l in body if l not in gen.splitlines()]
    assert changed == ["  FFTWAdapter fftAdapter(dims);"], changed
    # a second, syntax-only TU: every header a reference demo pulls in, ours first
    tu = tmp_path / "tu.cpp"
    tu.write_text('#include "diskSamplerLinear.h"\n#include "plummerSampler.h"\n#include "diskSampler.h"\n'
                  '#include "barnesHut.h"\n#include "ppMethod.h"\n#include "RK4Stepper.h"\n#include "utils.h"\n'
                  '#include "pmMethod.h"\n#include "p3mMethod.h"\n#include "PMMethodGPU.h"\n#include "simInfo.h"\n'
                  '#include "leapfrog.h"\n#include "unitConversions.h"\n#include "chainingMesh.h"\n#include "grid.h"\n'
                  '#include "greensFunctions.h"\n#include "stateRecorder.h"\n#include "externalFields.h"\n'
                  'static_assert(P3M_B200_REFERENCE_TREE == 1, "the reference\'s PODs must be the ones in use");\n'
                  'int main() { Vec3 v = Vec3(1, 2, 3); Particle p(v, v, 1.0f); return (int)p.mass - 1; }\n')
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-w", "-include",
                        os.path.join(ROOT, "oracle", "shim", "msvc_compat.h"), "-I", os.path.join(HOST, "include"),
                        "-I", os.path.join(REFERENCE, "include"), str(tu)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


def test_standalone_headers_define_their_own_pods(tmp_path):
    tu = tmp_path / "tu.cpp"
    tu.write_text('#include "vec3.h"\n#include "stateRecorder.h"\n#include "pmMethod.h"\n#include "p3mMethod.h"\n'
                  'static_assert(P3M_B200_REFERENCE_TREE == 0, "standalone mode");\nint main() { return 0; }\n')
    r = subprocess.run(["g++", "-std=c++20", "-fsyntax-only", "-I", os.path.join(HOST, "include"), "-I",
                        os.path.join(HOST, "include", "standalone"), str(tu)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-3000:]


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_host_grid_backfill_and_pmmethodgpu_extras(tmp_path):
    """getGrid() const back-fill, Grid accessors (periodic getPotential), copyGrid*ToHost + getGridDensity /
    getGridPotential, copyParticles*, both SimInfo::potentialEnergy mesh overloads, P3MMethod argument checks."""
    p, pos, vel, mass, box = case(3000, cubic=False)
    p.extKind = 1
    p.extCenter[:] = [b / 2 for b in box]
    p.extR, p.extM = 3.0, 60.0
    d = run_units("gpu", tmp_path, p, pos, vel, mass, box)
    o = refapi.Oracle("f32")
    pc, _, mc = o.to_code_units(p, pos, vel, mass)
    rho, phi, acc = o.force(p, False, o.green(p), pc, mc)
    for key in ("grid_density", "gpu_density"):
        assert rel_l2(d[key], rho.ravel()) < 1e-4, key
    for key in ("grid_potential", "gpu_potential"):
        assert rel_l2(d[key], phi.ravel()) < 1e-4, key
    gphi = d["grid_potential"].reshape(phi.shape)
    assert d["grid_potential_wrap"][0] == gphi[1, 0, -1] and d["grid_potential_wrap"][1] == gphi[1, 0, -1]
    for key in ("pm_acc", "gpu_acc"):
        assert rel_l2(d[key].reshape(-1, 3), acc) < 1e-4, key
    assert np.array_equal(d["pm_pos"].reshape(-1, 3), pc)
    misc = d["misc"]
    H, DT, G = float(p.H), float(p.DT), float(p.G)
    internal = (rho.astype(np.float64) / (DT * DT * 4 * np.pi * G) * phi.astype(np.float64) * H * H / (DT * DT)).sum()
    c = np.array(box) / 2
    r = np.linalg.norm(pos.astype(np.float64) - c, axis=1)
    u = r / 3.0
    ext = np.where(r > 3.0, -G * 60.0 / r, G * 60.0 / 3.0 * (-2 + u * u * (2 - u)))
    pe = 0.5 * H ** 3 * internal + (mass.astype(np.float64) * ext).sum()
    assert abs(misc[0] - pe) < 2e-4 * abs(pe) and abs(misc[8] - pe) < 2e-4 * abs(pe)
    gg = np.where(r > 3.0, -G * 60.0 / r ** 2, -(G * 60.0 / 27.0) * r * (4 - 3 * r / 3.0))
    force = (mass.astype(np.float64)[:, None] * gg[:, None] * (pos - c) / r[:, None]).sum(0)
    assert np.allclose(misc[1:4], force, rtol=1e-3, atol=1e-4 * np.abs(force).max())
    assert misc[4] == 0.0 and np.allclose(misc[5:8], [H, DT, G])
    assert misc[9] == 1.0, "P3MMethod must refuse an H that differs from its PMMethod's"


@pytest.mark.gpu
@needs_ref
@pytest.mark.skipif(not os.path.exists(REFTREE), reason="host/reftree_build/demo_reftree not built")
def test_reference_tree_demo_runs_galaxy_p3m_like_the_reference(tmp_path):
    """The binary built from the reference's own galaxySimulationP3M body (50 000-particle linear disk,
    128 x 128 x 64, TSC, S1-optimal, P3M, bulge field through the std::function callback, 200 steps) runs on
    the GPU; its first diagnostic rows follow the UNMODIFIED reference's CPU run of the same demo."""
    out = tmp_path / "rt"
    out.mkdir()  # the reference's StateRecorder does not create its output directory
    r = subprocess.run([REFTREE, str(out)], capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    diag = np.concatenate([np.loadtxt(out / f, ndmin=2) for f in
                           ("energy.txt", "momentum.txt", "angular_momentum.txt", "expected_momentum.txt")], axis=1)
    assert diag.shape[0] == 201 and np.isfinite(diag).all()
    with open(out / "positions.dat", "rb") as f:
        n, frames = struct.unpack("ii", f.read(8))
        first = np.frombuffer(f.read(12 * n), np.float32).reshape(n, 3)
    assert (n, frames) == (50000, 201)
    # the reference's CPU run of the same demo, 3 steps (its disk sampler is deterministic under libstdc++)
    ref = refapi.Ref()
    n = 50000
    pos, vel = ref.sample_disk_linear(42, (30.0, 30.0, 15.0), 3.0, 60.0, 15.0, 15.0, 0.3, 4.5e-3, n)
    mass = np.full(n, np.float32(15.0) / np.float32(n), np.float32)
    p = refapi.make_params(n, (128, 128, 64), (60.0, 60.0, 30.0), gfunc=refapi.S1_OPTIMAL, softening=1.5,
                           ext=dict(center=(30.0, 30.0, 15.0), R=3.0, M=60.0))
    steps = 3
    dref, _, _, _ = ref.run(p, pos, vel, mass, steps, True, tmp_path / "ref")
    with open(tmp_path / "ref" / "positions.dat", "rb") as f:
        f.read(8)
        first_ref = np.frombuffer(f.read(12 * n), np.float32).reshape(n, 3)
    assert rel_l2(first, first_ref) < 1e-5
    scale = np.abs(dref[:, 1]).max()
    assert np.abs(diag[:steps + 1, :2].sum(1) - dref[:, :2].sum(1)).max() < 2e-3 * scale
    assert np.abs(diag[:steps + 1, 1] - dref[:, 1]).max() < 2e-3 * scale
    lscale = np.abs(dref[:, 5:8]).max()
    assert np.abs(diag[:steps + 1, 5:8] - dref[:, 5:8]).max() < 2e-3 * lscale

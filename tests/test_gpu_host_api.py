"""The reference-facing C++ API (particlesimulation_b200/host: PMMethod, P3MMethod, Grid, CuFFTAdapter,
StateRecorder) driven by a demo-style caller, checked against the oracle's run of the same initial
conditions.  The files it writes are the reference's formats (positions.dat, energy.txt, ...)."""
import os
import struct
import subprocess

import numpy as np
import pytest

import refapi
from common import disk_case, plummer_case
from refapi import Oracle, rel_l2

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEMO = os.path.join(ROOT, "particlesimulation_b200", "host", "demo_host")


def write_ic(path, pos, vel, mass):
    with open(path, "wb") as f:
        f.write(struct.pack("i", len(mass)))
        f.write(np.ascontiguousarray(pos, np.float32).tobytes())
        f.write(np.ascontiguousarray(vel, np.float32).tobytes())
        f.write(np.ascontiguousarray(mass, np.float32).tobytes())


def read_positions(path):
    """script/load_data.py:15-37 layout: int32 n, int32 frames, frames x n x float32[3]."""
    with open(path, "rb") as f:
        n, frames = struct.unpack("ii", f.read(8))
        data = np.frombuffer(f.read(), np.float32)
    return data.reshape(-1, n, 3), frames


def run_demo(mode, tmp_path, pos, vel, mass, steps, grid, env=None):
    ic = tmp_path / "ic.bin"
    out = tmp_path / ("out_" + mode + ("_cb" if env else ""))
    write_ic(ic, pos, vel, mass)
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run([DEMO, mode, str(ic), str(out), str(steps)] + [str(g) for g in grid],
                       capture_output=True, text=True, env=e, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    diag = np.concatenate([np.loadtxt(out / f, ndmin=2) for f in
                           ("energy.txt", "momentum.txt", "angular_momentum.txt", "expected_momentum.txt")], axis=1)
    return out, diag


def test_fft_adapter_contract():
    r = subprocess.run([DEMO, "fft-roundtrip"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_p3m_plummer_demo_matches_oracle_run(tmp_path):
    n, steps, grid = 3000, 20, (32, 32, 32)
    p, pos, vel, mass = plummer_case(n, grid=grid)
    out, diag = run_demo("p3m-plummer", tmp_path, pos, vel, mass, steps, grid)
    diag_ref, pos_ref, vel_ref, _ = Oracle("f32").run(p, True, pos, vel, mass, steps)
    assert diag.shape[0] == steps + 1
    ke = np.abs(diag_ref[:, 1]).max()
    assert np.abs(diag[:, 0] + diag[:, 1] - diag_ref[:, 0] - diag_ref[:, 1]).max() < 2e-3 * ke + 2.2e-6
    pscale = np.abs(mass.astype(np.float64)[:, None] * vel).sum()
    assert np.abs(diag[:, 2:5] - diag_ref[:, 2:5]).max() < 2e-3 * pscale
    frames, nframes = read_positions(out / "positions.dat")
    assert nframes == steps + 1 and frames.shape == (steps + 1, n, 3)
    final = np.fromfile(out / "final.bin", np.float32).reshape(2, n, 3)  # getParticles(): code units
    assert rel_l2(final[0], pos_ref) < 1e-4
    assert rel_l2(frames[-1], pos_ref * float(p.H)) < 1e-4


def test_pm_disk_demo_device_field_and_host_callback_agree(tmp_path):
    n, steps, grid = 3000, 10, (32, 32, 16)
    p, pos, vel, mass = disk_case(n, grid=grid, gfunc=refapi.DISCRETE_LAPLACIAN)
    _, d_dev = run_demo("pm-disk", tmp_path, pos, vel, mass, steps, grid)
    _, d_cb = run_demo("pm-disk", tmp_path, pos, vel, mass, steps, grid, env={"DEMO_HOST_CALLBACK": "1"})
    diag_ref, _, _, _ = Oracle("f32").run(p, False, pos, vel, mass, steps)
    scale = np.abs(diag_ref[:, 1]).max()
    for d in (d_dev, d_cb):
        assert np.abs(d[:, 0] + d[:, 1] - diag_ref[:, 0] - diag_ref[:, 1]).max() < 2e-3 * scale + 2.2e-6
        assert np.abs(d[:, 8:11] - diag_ref[:, 8:11]).max() < 2e-3 * (np.abs(diag_ref[:, 8:11]).max() + 1e-30) + 1e-6
    assert np.abs(d_dev[:, :2] - d_cb[:, :2]).max() < 1e-3 * scale + 2.2e-6

#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: P3M particle-steps/s.

  python bench.py --gpus N --steps K --warmup W            our arm   (C ABI -> sm_100a kernels)
  python bench.py --impl reference --steps K --warmup W     the reference's own CPU path, same host

One "step" = one pass of the hot path: drift -> bin/sort -> deposit -> Poisson -> gather -> short range
-> kick (the body of the reference's run loop, source/p3mMethod.cpp:101-153).  At N = 1 the workload is
BASELINE.json configs[1]: P3M Plummer sphere, 2^20 particles, 128^3 mesh, TSC, S1-optimal influence
function, a = 3H, re = 0.7a, eps = 0.5, tabulated short-range force (source/demos.cpp:1416-1447).
At N > 1 (torchrun) it is BASELINE.json configs[4], the COUPLED weak-scaling sweep: 2^23 particles per GPU of
one uniform distribution on meshes 256^3 / 256^2 x 512 / 512^2 x 256 / 512^3; at N = 8 the line also carries
extra.c3 (configs[2], strong) and extra.c4 (configs[3], 2^26 particles on 1024^3).  Every line has a `parity`
object measured in the same run.  Initial conditions are generated on the device (p3m_generate_particles).

`value`   particle-steps/s with the particles resident in HBM (CUDA events on the context's stream).
`e2e`     the same step through the C ABI with HOST buffers: p3m_set_particles (H2D from pinned
          memory) + p3m_step(1) + p3m_get_particles (D2H) inside the timed region -- what the
          reference's own CUDA build does every P3M step (source/p3mMethod.cpp:143-151).
`roofline` the dominant kernel (short-range PP, FP32 pipe) + per-kernel HBM figures in `roofline_kernels`.
`cpu_baseline` the UNMODIFIED reference compiled in oracle/_ref, timed on this host on a bounded sample.
Nothing under oracle/ is used by the GPU arm's timed path.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_C2 = 1 << 20
GRID_C2 = (128, 128, 128)
BOX_C2 = (60.0, 60.0, 60.0)


# ---------------------------------------------------------------------------------------- cudart
class Cuda:
    def __init__(self):
        self.rt = C.CDLL("libcudart.so.12")
        self.rt.cudaEventElapsedTime.argtypes = [C.POINTER(C.c_float), C.c_void_p, C.c_void_p]
        self.rt.cudaEventRecord.argtypes = [C.c_void_p, C.c_void_p]
        self.rt.cudaEventSynchronize.argtypes = [C.c_void_p]
        self.rt.cudaMemsetAsync.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_void_p]
        self.rt.cudaHostAlloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t, C.c_uint]
        self.rt.cudaMalloc.argtypes = [C.POINTER(C.c_void_p), C.c_size_t]

    def check(self, rc, what=""):
        if rc != 0:
            raise RuntimeError(f"CUDA error {rc} {what}")

    def event(self):
        e = C.c_void_p()
        self.check(self.rt.cudaEventCreate(C.byref(e)), "cudaEventCreate")
        return e

    def record(self, e, stream):
        self.check(self.rt.cudaEventRecord(e, stream), "cudaEventRecord")

    def elapsed_ms(self, a, b):
        self.check(self.rt.cudaEventSynchronize(b), "cudaEventSynchronize")
        ms = C.c_float(0)
        self.check(self.rt.cudaEventElapsedTime(C.byref(ms), a, b), "cudaEventElapsedTime")
        return float(ms.value)

    def pinned(self, shape, dtype=np.float32):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = C.c_void_p()
        self.check(self.rt.cudaHostAlloc(C.byref(p), max(n, 16), 0), "cudaHostAlloc")
        buf = (C.c_byte * n).from_address(p.value)
        return np.frombuffer(buf, dtype=dtype).reshape(shape)

    def malloc(self, nbytes):
        p = C.c_void_p()
        self.check(self.rt.cudaMalloc(C.byref(p), nbytes), "cudaMalloc")
        return p

    def set_device(self, d):
        self.check(self.rt.cudaSetDevice(d), "cudaSetDevice")

    def sync(self):
        self.check(self.rt.cudaDeviceSynchronize(), "cudaDeviceSynchronize")


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([time.time()] + [x.strip() for x in line.split(",")])

    def mark(self):
        """Start of the timed region: the sampler itself is started earlier (nvidia-smi takes ~0.5 s to
        deliver its first line), samples taken before the mark are dropped."""
        self.t0 = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        t0 = getattr(self, "t0", 0.0)
        rows = [r[1:] for r in self.rows if r[0] >= t0]
        if not rows and self.rows:  # region shorter than one sampling period: nearest sample
            rows = [self.rows[-1][1:]]
        for r in rows:
            if len(r) < 7:
                continue
            try:
                sm.append(float(r[0])); mx = float(r[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "reasons": sorted(reasons)}


# -------------------------------------------------------------------------------------- workloads
F32 = np.float32
H_C2 = float(F32(F32(BOX_C2[0]) / F32(GRID_C2[0] // 2)))  # 60 / 64


def p3m_params(capi, grid, box, timing=True, softening=0.5, p3m=1, gfunc=None, device=None):
    """The reference demo's P3M parameters (source/demos.cpp:1416-1447) on a given mesh: H = box_x / (Nx / 2),
    TSC, 2-point differences, S1-optimal influence function, a = 3H, re = 0.7a, tabulated short-range force."""
    p = capi.default_params()
    p.nx, p.ny, p.nz = grid
    p.box[:] = box
    p.H = F32(F32(box[0]) / F32(grid[0] // 2))
    p.DT, p.G = 1.0, 4.5e-3
    p.assignment, p.fd_scheme = capi.TSC, capi.TWO_POINT
    p.greens_function = capi.S1_OPTIMAL if gfunc is None else gfunc
    p.particle_diameter = F32(F32(3) * F32(p.H))
    p.p3m = p3m
    p.cutoff_radius = F32(F32(0.7) * F32(p.particle_diameter))
    p.softening = softening
    p.cloud_shape, p.use_sr_table = capi.S1, 1
    p.precision = capi.F32
    p.unit_roundtrip = 1
    p.green_zero_degenerate = 1
    p.timing = int(timing)
    if device is not None:
        p.device = device
    return p


def c2_params(capi, timing=True):
    """BASELINE.json configs[1]: P3M Plummer sphere, 2^20 particles, 128^3 mesh."""
    return p3m_params(capi, GRID_C2, BOX_C2, timing)


def c2_ic(capi, n):
    """PlummerSampler(42).sample(center=(30,30,30), a=2, rMax=15, M=1, G=4.5e-3, n) (source/demos.cpp:1350,
    1416-1447), from the device-side sampler (same distribution, counter-based stream)."""
    return capi.ic_plummer(n, center=(30.0, 30.0, 30.0), a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=42)


C5_GRIDS = {1: (256, 256, 256), 2: (256, 256, 512), 4: (512, 512, 256), 8: (512, 512, 512)}
N_C5_PER_GPU = 1 << 23
VEL_SIGMA_CELLS = float(os.environ.get("P3M_BENCH_VEL_SIGMA", 0.05))   # velocity dispersion of the uniform sets, mesh cells per step


def uniform_margin_cells(args):
    """Uniform sets drift freely (their self-gravity is negligible over a bench run): keep every particle
    inside the box for all the steps one context executes (warm-up + timed + end-to-end), 6 sigma."""
    total = args.warmup + 2 * max(args.steps, 10) + 4
    if "P3M_BENCH_MARGIN" in os.environ:
        return float(os.environ["P3M_BENCH_MARGIN"])
    return max(4.0, 6 * VEL_SIGMA_CELLS * total)


def c5_setup(capi, world, n_per_gpu=N_C5_PER_GPU, timing=True, device=None, margin=8.0):
    """BASELINE.json configs[4]: P3M weak-scaling sweep, 2^23 particles per GPU, meshes 256^3 / 256^2 x 512 /
    512^2 x 256 / 512^3 (0.5 particles per mesh cell throughout), ONE global distribution: a
    uniform cube filling the occupied half of the mesh per axis (~4 particles per occupied cell, ~160 partners
    inside the cutoff), with a small velocity dispersion, so that every slab boundary is crossed by migrating
    particles and lined with ghost layers.  `margin` (cells) keeps the drifting set inside the box."""
    grid = C5_GRIDS[world]
    box = tuple(H_C2 * (g // 2) for g in grid)
    prm = p3m_params(capi, grid, box, timing, device=device)
    n = n_per_gpu * world
    lo = [margin * H_C2] * 3
    hi = [b - margin * H_C2 for b in box]
    ic = capi.ic_uniform(n, lo, hi, total_mass=float(world), vel_sigma=VEL_SIGMA_CELLS * H_C2, seed=42)
    return prm, ic, grid


def c3_setup(capi, n=1 << 24, grid=512, timing=True, device=None, margin=8.0):
    """BASELINE.json configs[2]: PM uniform cube, 2^24 particles, 512^3 mesh, TSC, discrete Laplacian."""
    box = (60.0, 60.0, 60.0)
    prm = p3m_params(capi, (grid,) * 3, box, timing, p3m=0, gfunc=capi.DISCRETE_LAPLACIAN, device=device)
    H = float(prm.H)
    ic = capi.ic_uniform(n, [margin * H] * 3, [60.0 - margin * H] * 3, total_mass=1.0, vel_sigma=VEL_SIGMA_CELLS * H, seed=42)
    return prm, ic


def c4_setup(capi, n=1 << 26, grid=1024, timing=True, device=None, margin=None):
    """BASELINE.json configs[3]: P3M clustered disk + halo, 2^26 particles, 1024^3 mesh; the softening of C2 in
    mesh units."""
    box = BOX_C2
    Hc = float(F32(F32(box[0]) / F32(grid // 2)))
    prm = p3m_params(capi, (grid,) * 3, box, timing, softening=float(F32(0.5 * Hc / H_C2)), device=device)
    ic = capi.ic_disk_halo(n, box=60.0, seed=42)
    return prm, ic


# kept for the gloo host-logic test (tests/test_dist_cpu.py): one C2 cluster per z-slab
def multi_params(capi, world, timing=False):
    p = c2_params(capi, timing)
    p.nz = GRID_C2[2] * world
    p.box[2] = BOX_C2[2] * world
    return p


def multi_particles(world, n_per_gpu):
    from particlesimulation_b200 import ics
    pos, vel, mass = [], [], []
    for r in range(world):
        a, b, c = ics.plummer(n_per_gpu, center=(30.0, 30.0, 30.0 + 60.0 * r), a=2.0, r_max=15.0, M=1.0, G=4.5e-3,
                              seed=42 + r)
        pos.append(a); vel.append(b); mass.append(c)
    return np.concatenate(pos), np.concatenate(vel), np.concatenate(mass)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def profile_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the last (by tag) committed
    `ncu --set full` summary under profiles/ (tools/summarize_ncu.py); None if there is none."""
    import glob
    import re
    files = sorted(f for f in glob.glob(os.path.join(ROOT, "profiles", f"{kernel}_*.txt")) if "_sass_" not in f)  # by tag
    if not files:
        return None, None
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot = 0.0
    for line in open(files[-1]):
        m = re.match(r"dram__bytes_(read|write)\.sum\s+([0-9.,]+)\s+(\w+)", line)
        if m:
            tot += float(m.group(2).replace(",", "")) * scale.get(m.group(3), 1.0)
    return (tot if tot > 0 else None), os.path.basename(files[-1])


# ------------------------------------------------------------------------------------ reference arm
def reference_sample_params(n):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import refapi
    return refapi, refapi.make_params(n, GRID_C2, BOX_C2, softening=0.5)


def run_reference(args, as_baseline=False):
    """The reference's own CPU implementation of the path (oracle/_ref = its unmodified sources),
    on a bounded sample of the workload: same mesh, same physics, fewer particles."""
    n = int(os.environ.get("P3M_BENCH_CPU_N", 1 << 15))
    steps = max(1, args.steps if not as_baseline else 2)
    warm = 0
    refapi, p = reference_sample_params(n)
    from particlesimulation_b200 import ics  # numpy sampler: the reference arm needs no GPU
    pos, vel, mass = ics.plummer(n, center=(30.0, 30.0, 30.0), a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=42)
    if refapi.have_ref():
        ref = refapi.Ref()
        cores = ref.hardware_threads()
        t0 = time.time()
        ms, green_ms = ref.time_steps(p, pos, vel, mass, steps, True)
        wall = time.time() - t0
        kind = "reference"
        total_ms = ms["total"]
        breakdown = {k: round(v / steps, 3) for k, v in ms.items() if k != "total"}
        note = (f"unmodified reference sources (oracle/_ref, g++ -O3, kissfft; PM loops serial: libstdc++ PSTL "
                f"without TBB; short-range loop on {cores} std::threads)")
    else:
        o = refapi.Oracle("f32")
        cores = 1
        t0 = time.time()
        o.run(p, True, pos, vel, mass, steps - 1, diagnostics=False)
        total_ms = (time.time() - t0) * 1e3
        wall = total_ms / 1e3
        green_ms = float("nan")
        kind = "port"
        breakdown = {}
        note = "oracle/p3m_oracle.c (scalar C restatement, includes the influence-function set-up)"
    value = n * steps / (total_ms / 1e3)
    sample = (f"P3M Plummer N={n} (of 2^20) on the full 128^3 mesh, {steps} steps after one untimed force "
              f"evaluation; influence-function init {green_ms / 1e3:.1f} s excluded; {note}. Short-range cost grows "
              f"~N^2 in the Plummer core, so the full-size CPU step is far slower than this sample suggests.")
    base = {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": kind, "sample": sample,
            "ms_per_step": total_ms / steps, "ms_per_step_by_phase": breakdown, "wall_s": round(wall, 1)}
    if as_baseline:
        return base
    line = {"impl": "reference", "metric": "P3M particle-steps/s", "value": value, "unit": "particle-steps/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": total_ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C2 P3M Plummer 128^3 TSC (bounded CPU sample)", "particles": n,
                       "mesh": list(GRID_C2)},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))
    return line


# ------------------------------------------------------------------------------------------ parity
PARITY_SAMPLE = 4096
SHELL_NOTE = ("the short-range law is truncated at the cutoff (source/p3mMethod.cpp:258), i.e. discontinuous: a source whose "
              "r^2 lies within 2e-6 (relative) of cutoff^2 is in or out depending on the fp32 rounding of r^2, in the "
              "reference's own arithmetic as well; sr_rel_l2 is taken over the sampled targets without such a source "
              "(cutoff_shell_rows are the others), sr_rel_l2_all_rows over all of them")


def sr_direct_host(capi, ctx, pc, mcode, ids):
    """fp64 evaluation on the HOST of the short-range sum for the particles `ids` against ALL particles, from
    the reference's formulas (shortRangeForceFromTable, source/p3mMethod.cpp:240-245; table initSRForceTable
    :275-294): no cells, no sort, no device code.  pc = fp32 code-unit positions, mcode = code-unit masses."""
    tab = ctx.sr_table()
    prm = ctx.params
    re = np.float64(F32(F32(prm.cutoff_radius) / F32(prm.H)))
    re2 = re * re
    d2t = re2 / 499.0
    pc32 = np.ascontiguousarray(pc, np.float32)
    out = np.zeros((len(ids), 3))
    for k, i in enumerate(ids):
        d32 = pc32[i][None, :] - pc32                      # exact for close pairs
        r2 = np.einsum("ij,ij->i", d32, d32)
        sel = np.nonzero((r2 < np.float32(re2 * 1.001)))[0]
        d = pc32[i].astype(np.float64)[None, :] - pc32[sel].astype(np.float64)
        r2 = (d * d).sum(1)
        ok = (r2 < re2) & (r2 > 0)
        d, r2, m = d[ok], r2[ok], mcode[sel][ok].astype(np.float64)
        xi = r2 / d2t
        t = np.minimum(xi.astype(np.int64), 498)
        F = tab[t] + (xi - t) * (tab[t + 1] - tab[t])
        out[k] = ((m * F)[:, None] * d).sum(0)
    return out


def mass_code(prm, m):
    """massToCodeUnits (include/unitConversions.h:42-44), fp32, left to right."""
    DT, H, G, pi = F32(prm.DT), F32(prm.H), F32(prm.G), F32(np.pi)
    return float(F32(DT * DT * F32(4) * pi * G / (H * H * H)) * F32(m))


def rel_l2(a, b):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-300))


def parity_sample_ids(n, k=PARITY_SAMPLE):
    return np.sort(np.random.default_rng(12345).choice(n, min(k, n), replace=False))


def gather_rows(dist, torch, local_rows, n_rows, width):
    """id-indexed rows that are zero on the ranks that do not hold the particle -> sum over ranks."""
    t = torch.from_numpy(np.ascontiguousarray(local_rows, np.float64)).cuda()
    dist.all_reduce(t)
    return t.cpu().numpy()


# ------------------------------------------------------------------------------------------ our arm
def timed_window(cu, ctx, steps, flush, flush_bytes, clocks=None, barrier=None):
    """K timed steps with the L2 evicted before each one; returns per-step device ms."""
    ev = [(cu.event(), cu.event()) for _ in range(steps)]
    if barrier:
        barrier()
    cu.sync()
    if clocks:
        clocks.mark()
    for a, b in ev:
        cu.check(cu.rt.cudaMemsetAsync(flush, 0, flush_bytes, ctx.stream), "flush")  # evict L2 (126 MB)
        cu.record(a, ctx.stream)
        ctx.step(1)
        cu.record(b, ctx.stream)
    cu.sync()
    if barrier:
        barrier()
    return [cu.elapsed_ms(a, b) for a, b in ev]


def run_single(args, capi, cu, prm, ic, label, flush, flush_bytes, want_e2e=True, clocks=None):
    """One GPU: generate on the device, warm up, K timed steps (value + phase table from the SAME steps), then
    K end-to-end steps through host buffers on the same context."""
    ctx = capi.Context(prm)
    ctx.generate_particles(ic)
    n = ctx.n
    ctx.green_init()
    ctx.force()
    ctx.kick(0.5)  # setHalfStepVelocities
    for _ in range(args.warmup):
        ctx.step(1)
    ctx.phase_ms(reset=True)
    launches0 = ctx.launches
    t_wall0 = time.time()
    ms = timed_window(cu, ctx, args.steps, flush, flush_bytes, clocks)
    t_wall = time.time() - t_wall0
    if ctx.escaped():
        raise RuntimeError(f"{label}: a particle left the computational box during the timed steps (the run loop would "
                           "have stopped, source/pmMethod.cpp:108-111): the measurement is void")
    phases = {k: v / args.steps for k, v in ctx.phase_ms(reset=True).items()}
    launches = ctx.launches - launches0
    out = {"ctx": ctx, "n": n, "ms": ms, "phases": phases, "launches": launches, "wall": t_wall, "label": label}
    if want_e2e:
        hp, hv, hm = cu.pinned((n, 3)), cu.pinned((n, 3)), cu.pinned((n,))
        op, ov = cu.pinned((n, 3)), cu.pinned((n, 3))
        cpos, cvel, _ = ctx.get_particles(capi.UNITS_ORIGINAL, want=("pos", "vel"))
        hp[:], hv[:] = cpos, cvel
        hm[:] = np.float32(ic.total_mass) / np.float32(n)
        lib = capi.lib()
        t0 = None
        for it in range(args.steps + 1):
            if it == 1:
                cu.sync()
                t0 = time.time()
            cu.check(cu.rt.cudaMemsetAsync(flush, 0, flush_bytes, ctx.stream), "flush")
            ctx.set_particles(hp, hv, hm)          # H2D: pos, vel, mass (28 B / particle)
            ctx.step(1)                             # drift -> force -> kick (needs x, v, m only)
            rc = lib.p3m_get_particles(ctx._h, op.ctypes.data_as(C.c_void_p), ov.ctypes.data_as(C.c_void_p), None,
                                       capi.UNITS_ORIGINAL)  # D2H: pos, vel (24 B / particle)
            assert rc == 0
            hp, op = op, hp                         # the caller's next step starts from what came back
            hv, ov = ov, hv
        out["e2e_s"] = (time.time() - t0) / args.steps
    return out


def hbm_entry(bytes_, ms_, hbm_peak):
    gbs = bytes_ / (ms_ / 1e3) / 1e9 if ms_ > 0 else 0.0
    return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
            "algorithmic_bytes": bytes_, "ms": ms_}


def mesh_rooflines(n, M, ph, hbm_peak):
    """SURVEY section 8d algorithmic bytes per force evaluation."""
    return {
        "binSort": hbm_entry(76.0 * n, ph["binSort"], hbm_peak),
        "spreadMass": hbm_entry(16.0 * n + 8.0 * M, ph["spreadMass"], hbm_peak),
        "poisson(fwdFFT+multiply+invFFT)": hbm_entry(18.0 * M, ph["forwardFFT"] + ph["fourierPotential"] + ph["inverseFFT"], hbm_peak),
        "updateAccelerations(fused gradient+gather)": hbm_entry(4.0 * M + 28.0 * n, ph["updateAccelerations"], hbm_peak),
        "integrate": hbm_entry(2 * 48.0 * n, ph["integrate"], hbm_peak),
    }


def run_ours(args):
    from particlesimulation_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", 1))
    if world > 1 or args.gpus > 1:
        return run_ours_multi(args)
    cu = Cuda()
    cu.set_device(0)
    hbm_peak, peak_src, sm_max = load_peaks()
    flush_bytes = 256 << 20
    flush = cu.malloc(flush_bytes)
    if args.config == "mesh":
        return run_c3_single(args, capi, cu, flush, flush_bytes, hbm_peak, peak_src)
    if args.config == "c5":
        prm5, ic5, grid5 = c5_setup(capi, 1, int(os.environ.get("P3M_BENCH_N", N_C5_PER_GPU)), margin=uniform_margin_cells(args))
        r5 = run_single(args, capi, cu, prm5, ic5, "C5 at 1 GPU", flush, flush_bytes, want_e2e=False)
        ch5, in5 = r5["ctx"].pair_counts()
        st5 = r5["ctx"].stats()
        r5["ctx"].close()
        print(json.dumps({"metric": "P3M particle-steps/s (secondary: 1-GPU point of the weak-scaling sweep)",
                          "value": ic5.n * args.steps / (float(np.sum(r5["ms"])) / 1e3), "unit": "particle-steps/s", "n_gpus": 1,
                          "ms_per_step": float(np.mean(r5["ms"])), "config": {"workload": f"C5 at 1 GPU: 2^23 particles, mesh {grid5}"},
                          "ms_per_step_by_phase": {k: round(v, 4) for k, v in r5["phases"].items()},
                          "pairs_in_range_per_particle": in5 / ic5.n, "pairs_checked_per_particle": ch5 / ic5.n, "stats": st5,
                          "roofline_kernels": mesh_rooflines(ic5.n, grid5[0] * grid5[1] * grid5[2], r5["phases"], hbm_peak)}))
        return
    n = int(os.environ.get("P3M_BENCH_N", N_C2))
    clocks = ClockSampler(0)
    clocks.start()
    r = run_single(args, capi, cu, c2_params(capi), c2_ic(capi, n), "C2", flush, flush_bytes, clocks=clocks,
                   want_e2e=not args.quick)
    clk = clocks.stop()
    ctx, ms, phases = r["ctx"], r["ms"], r["phases"]
    total_ms = float(np.sum(ms))
    value = n * args.steps / (total_ms / 1e3)
    M = GRID_C2[0] * GRID_C2[1] * GRID_C2[2]
    if args.quick:
        print(json.dumps({"metric": "P3M particle-steps/s", "value": value, "ms_per_step": total_ms / args.steps, "quick": True,
                          "ms_per_step_by_phase": phases}))
        ctx.close()
        return

    # ---- roofline of the dominant kernel: exact pair statistics from the counting instantiation (read-only)
    checked, inside = ctx.pair_counts()
    stats = ctx.stats()
    pp_flops = 9.0 * checked + 14.0 * inside
    pp_ms = phases["shortRangeForcesCalc"]
    fp32_peak = 148 * 128 * 2 * sm_max * 1e6 / 1e12  # TFLOP/s at clocks.max.sm, FMA = 2 flop
    pp_tflops = pp_flops / (pp_ms / 1e3) / 1e12 if pp_ms > 0 else 0.0
    kname = "k_pp_packed" if stats["packed_pp"] else "k_pp_tiled"
    traffic, traffic_file = profile_traffic(kname)
    roofline = {"kernel": f"{kname} (short-range PP, dense chaining cells) + k_pp_sparse", "bound": "fp32",
                "achieved": pp_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": pp_tflops / fp32_peak if fp32_peak else None, "traffic": traffic,
                "traffic_source": f"dram bytes read + written per launch, ncu --set full capture profiles/{traffic_file}",
                "peak_source": f"148 SMs x 128 FP32 lanes x 2 flop x clocks.max.sm {sm_max:.0f} MHz (no tensor cores on "
                               f"this path; HBM peak for the other kernels: {peak_src})",
                "flops_per_launch": pp_flops, "pairs_checked": checked, "pairs_in_range": inside, "ms": pp_ms,
                "ms_source": "CUDA events around the short-range phase of the SAME timed steps as `value`"}

    # ---- parity of THIS run: short-range accelerations of a fixed sample against a host fp64 evaluation of the
    # reference's formulas on the same positions, and against the device brute-force sum (N4)
    gpos = ctx.get_particles(capi.UNITS_CODE, want=("pos",))[0]
    _, sr = ctx.acc_parts()
    ids = parity_sample_ids(n)
    mcode = np.full(n, mass_code(ctx.params, np.float32(1.0) / np.float32(n)), np.float64)
    sr_dev = ctx.direct_sum(gpos[ids].astype(np.float64), capi.SUM_SHORT_RANGE)
    shell = ctx.direct_sum(gpos[ids].astype(np.float64), capi.SUM_CUTOFF_SHELL)[:, 0] > 0
    hsub = np.arange(0, len(ids), 8)  # the host evaluation costs ~20 ms per particle: every 8th of the sample
    t0 = time.time()
    sr_host = sr_direct_host(capi, ctx, gpos, mcode, ids[hsub])
    host_s = time.time() - t0
    parity = {"sample": len(ids), "sr_rel_l2": rel_l2(sr[ids][~shell], sr_dev[~shell]), "tolerance": 1e-4,
              "sr_rel_l2_all_rows": rel_l2(sr[ids], sr_dev), "cutoff_shell_rows": int(shell.sum()), "cutoff_shell_note": SHELL_NOTE,
              "host_sample": len(hsub), "sr_rel_l2_vs_host_fp64": rel_l2(sr[ids[hsub]], sr_host),
              "device_direct_sum_vs_host_fp64_rel_l2": rel_l2(sr_dev[hsub], sr_host), "host_fp64_seconds": round(host_s, 1),
              "what": "short-range acceleration after the timed steps: 4096 sampled particles vs p3m_direct_sum (device fp64 "
                      "brute force over all particles, no cells / sort / culling); every 8th of them also vs a HOST fp64 "
                      "evaluation of source/p3mMethod.cpp:240-245 (numpy), which pins the brute-force kernel itself"}

    # ---- like-for-like sibling of the reference arm (which can only afford N = 2^15 on the CPU)
    n_small = int(os.environ.get("P3M_BENCH_CPU_N", 1 << 15))
    small = run_single(args, capi, cu, c2_params(capi), c2_ic(capi, n_small), "C2 at the reference arm's N", flush,
                       flush_bytes)
    small["ctx"].close()
    companion = {"particles": n_small, "value": n_small * args.steps / (float(np.sum(small["ms"])) / 1e3),
                 "ms_per_step": float(np.mean(small["ms"])), "e2e_value": n_small / small["e2e_s"],
                 "what": "the SAME workload `bench.py --impl reference` times (P3M Plummer, N = 2^15, 128^3 mesh), on "
                         "the GPU: divide by the reference arm's value for a same-config speed-up"}

    extra = {}
    if not args.no_extra:
        # the 1-GPU point of the coupled weak-scaling sweep that `--gpus N` (N > 1) reports (BASELINE configs[4])
        prm5, ic5, grid5 = c5_setup(capi, 1, margin=uniform_margin_cells(args))
        r5 = run_single(args, capi, cu, prm5, ic5, "C5 at 1 GPU", flush, flush_bytes, want_e2e=False)
        c5 = r5["ctx"]
        ch5, in5 = c5.pair_counts()
        c5.close()
        extra["c5_n1"] = {"workload": f"C5 at 1 GPU: P3M uniform cube, 2^23 particles, mesh {grid5}",
                          "value": ic5.n * args.steps / (float(np.sum(r5["ms"])) / 1e3),
                          "ms_per_step": float(np.mean(r5["ms"])), "ms_per_step_by_phase": r5["phases"],
                          "pairs_in_range_per_particle": in5 / ic5.n, "pairs_checked_per_particle": ch5 / ic5.n,
                          "roofline_kernels": mesh_rooflines(ic5.n, grid5[0] * grid5[1] * grid5[2], r5["phases"], hbm_peak)}

    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = run_reference(args, as_baseline=True)
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    line = {
        "metric": "P3M particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: P3M Plummer sphere, 2^20 particles, 128^3 mesh, TSC, S1-optimal Green, "
                               "chaining-mesh PP (re=0.7a, a=3H, eps=0.5, table)", "particles": n,
                   "mesh": list(GRID_C2), "l2": "256 MiB memset before every timed step (outside the events)",
                   "parallelism": "1 GPU", "initial_conditions": "device-side Plummer sampler (p3m_generate_particles)",
                   "note": "N = 1 is BASELINE configs[1] (C2); `--gpus N` with N > 1 runs the COUPLED weak-scaling sweep "
                           "configs[4] (C5) -- its 1-GPU point is extra.c5_n1 of this line"},
        "clocks": clk, "gpu_launches": r["launches"],
        "e2e": {"value": n / r["e2e_s"], "unit": "particle-steps/s", "h2d_bytes_per_step": 28 * n,
                "d2h_bytes_per_step": 24 * n, "ms_per_step": r["e2e_s"] * 1e3,
                "window": f"the {args.steps} steps right after the timed ones, same context, same L2 flush"},
        "roofline": roofline, "roofline_kernels": mesh_rooflines(n, M, phases, hbm_peak),
        "ms_per_step_by_phase": phases, "ms_per_step_each": ms, "wall_s_timed_region": r["wall"],
        "phase_source": "non-blocking CUDA events recorded during the timed steps themselves",
        "parity": parity, "same_n_companion": companion, "extra": extra, "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    ctx.close()
    return line


def run_c3_single(args, capi, cu, flush, flush_bytes, hbm_peak, peak_src):
    """Secondary measurement (not the headline line): BASELINE configs[2] on ONE GPU -- the mesh kernels
    (deposit, FFT, gather, sort, integrate) are the whole step.  Per-kernel HBM fractions."""
    n = int(os.environ.get("P3M_BENCH_N", 1 << 24))
    grid = int(os.environ.get("P3M_BENCH_GRID", 512))
    prm, ic = c3_setup(capi, n, grid, margin=uniform_margin_cells(args))
    r = run_single(args, capi, cu, prm, ic, "C3", flush, flush_bytes, want_e2e=False)
    st = r["ctx"].stats()
    r["ctx"].close()
    M = grid ** 3
    line = {"metric": "PM particle-steps/s (secondary, mesh-dominated)", "value": n * args.steps / (float(np.sum(r["ms"])) / 1e3),
            "unit": "particle-steps/s", "n_gpus": 1, "ms_per_step": float(np.mean(r["ms"])),
            "config": {"workload": f"C3: PM uniform cube, {n} particles, {grid}^3 mesh, TSC, 2-pt, discrete Laplacian"},
            "roofline_kernels": mesh_rooflines(n, M, r["phases"], hbm_peak),
            "ms_per_step_by_phase": {k: round(v, 4) for k, v in r["phases"].items()}, "stats": st,
            "hbm_peak": hbm_peak, "hbm_peak_source": peak_src}
    print(json.dumps(line))
    return line


def measure_multi(args, capi, pdist, dist, torch, cu, prm, ic, flush, flush_bytes, clocks=None, steps=None,
                  want_e2e=True, single_check=True, sr_check=True):
    """N GPUs (one process each): device-side generation of the local slabs, first force, parity of that first
    force (sample: N-GPU vs 1-GPU accelerations, short-range part vs fp64 brute force), warm-up, K timed steps
    with per-rank phase tables, exchange statistics, end-to-end loop."""
    rank, world = dist.get_rank(), dist.get_world_size()
    steps = steps or args.steps
    n_total = int(ic.n)
    ctx = pdist.create_context(prm, capi)
    ctx.generate_particles(ic)
    ctx.green_init()
    ctx.force()
    ids = parity_sample_ids(n_total)
    parity = {"sample": len(ids), "tolerance": 1e-4}
    # ---- parity on the first force evaluation (state = the initial conditions)
    pos_l, acc_l, sr_l = ctx.sample(ids)                                 # rows of the sample, zero where not held
    acc_s = gather_rows(dist, torch, acc_l, len(ids), 3)
    pos_s = gather_rows(dist, torch, pos_l, len(ids), 3)
    sr_s = gather_rows(dist, torch, sr_l, len(ids), 3)
    if sr_check and prm.p3m:
        part = ctx.direct_sum(pos_s, capi.SUM_SHORT_RANGE)   # this rank's particles against every sample point
        sr_direct = gather_rows(dist, torch, part, len(ids), 3)
        shell = gather_rows(dist, torch, ctx.direct_sum(pos_s, capi.SUM_CUTOFF_SHELL), len(ids), 3)[:, 0] > 0
        parity["sr_rel_l2"] = rel_l2(sr_s[~shell], sr_direct[~shell])
        parity["sr_rel_l2_all_rows"] = rel_l2(sr_s, sr_direct)
        parity["cutoff_shell_rows"] = int(shell.sum())
        parity["cutoff_shell_note"] = SHELL_NOTE
        parity["sr_check"] = "p3m_direct_sum: device fp64 brute force over ALL particles of all ranks (no cells / sort / culling)"
    if single_check:
        if rank == 0:
            sprm = type(prm).from_buffer_copy(prm)
            sprm.timing = 0
            one = capi.Context(sprm)
            one.generate_particles(ic)
            one.green_init()
            one.force()
            pos1, acc1, _ = one.sample(ids)
            one.close()
            parity["multi_vs_single_rel_l2"] = rel_l2(acc_s, acc1)
            parity["positions_identical"] = bool(np.array_equal(pos_s, pos1))
        dist.barrier()
    ctx.kick(0.5)
    for _ in range(args.warmup):
        ctx.step(1)
    ctx.phase_ms(reset=True)
    launches0 = ctx.launches
    ms = timed_window(cu, ctx, steps, flush, flush_bytes, clocks, barrier=dist.barrier)
    if ctx.escaped():
        raise RuntimeError("a particle left the computational box during the timed steps: the measurement is void")
    ph = ctx.phase_ms(reset=True)
    launches = ctx.launches - launches0
    names = sorted(ph)
    mine = torch.tensor([ph[k] / steps for k in names] + [float(np.sum(ms))], device="cuda", dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    table = torch.stack(allr).cpu().numpy()                 # [rank, phase]
    total_ms = float(table[:, -1].max())                    # device time, max over ranks
    phases = {k: [round(float(table[:, i].min()), 4), round(float(np.median(table[:, i])), 4), round(float(table[:, i].max()), 4)]
              for i, k in enumerate(names)}
    per_rank = {k: [round(float(x), 3) for x in table[:, i]] for i, k in enumerate(names)}
    per_rank["step_total"] = [round(float(x) / steps, 3) for x in table[:, -1]]
    st = ctx.stats()
    keys = ("migrated", "ghosts", "a2a_bytes", "density_plane_bytes", "potential_plane_bytes", "migration_bytes", "ghost_bytes")
    sv = torch.tensor([st[k] for k in keys] + [float(ctx.n)], device="cuda", dtype=torch.float64)
    alls = [torch.zeros_like(sv) for _ in range(world)]
    dist.all_gather(alls, sv)
    sm = torch.stack(alls).cpu().numpy()
    comm_ms = phases["comm"][2]
    sent = sm[:, 2:7].sum(axis=1)                           # bytes this rank sends per step over NVLink
    exchange = {"comm_ms_min_median_max": phases["comm"],
                "bytes_sent_per_rank_per_step_max": float(sent.max()), "bytes_sent_total_per_step": float(sent.sum()),
                "a2a_bytes_per_rank_per_step": float(sm[:, 2].max()),
                "nvlink_GBs_per_rank_achieved": float(sent.max() / (comm_ms / 1e3) / 1e9) if comm_ms > 0 else None,
                "nvlink_peak_GBs_per_direction": 770.0,
                "migrated_particles_per_step_total": float(sm[:, 0].sum()), "ghost_particles_total": float(sm[:, 1].sum()),
                "particles_per_rank_min_max": [int(sm[:, 7].min()), int(sm[:, 7].max())],
                "note": "comm = every NCCL exchange of a step (migration, ghost layers, density / potential planes, the two "
                        "all-to-all transposes, escape-flag all-reduce) INCLUDING the wait for the slowest rank; the other "
                        "phases contain no collective"}
    out = {"ctx": ctx, "ms_per_step": total_ms / steps, "value": n_total * steps / (total_ms / 1e3), "phases": phases,
           "exchange": exchange, "parity": parity, "launches": launches, "slab": bool(st["slab"]), "n_total": n_total,
           "per_rank": per_rank}
    if want_e2e:
        # every rank uploads the particles it holds (pinned host buffers, explicit ids), steps, and reads them back into
        # the other set of pinned buffers (ping-pong: no host-side copies in the loop)
        cap_h = int(1.3 * max(ctx.n, n_total // world)) + 4096
        bufs = [(cu.pinned((cap_h,), np.int32), cu.pinned((cap_h, 3)), cu.pinned((cap_h, 3))) for _ in range(2)]
        hm = cu.pinned((cap_h,))
        hm[:] = np.float32(ic.total_mass) / np.float32(n_total)
        idsl, lp, lv, _ = ctx.get_local(capi.UNITS_ORIGINAL, out=bufs[0])
        nl = len(idsl)
        h2d = d2h = 0
        t0 = None
        cur = 0
        for it in range(steps + 1):
            if it == 1:
                dist.barrier(); cu.sync(); t0 = time.time(); h2d = d2h = 0
            hid, hp, hv = bufs[cur]
            ctx.set_particles_ids(hp[:nl], hv[:nl], hm[:nl], hid[:nl])
            h2d += 32 * nl
            ctx.step(1)
            if ctx.n > cap_h:
                raise RuntimeError("end-to-end buffers too small for this rank's share")
            idsl, lp, lv, _ = ctx.get_local(capi.UNITS_ORIGINAL, out=bufs[cur ^ 1])
            nl = len(idsl)
            d2h += 28 * nl
            cur ^= 1
        cu.sync(); dist.barrier()
        te = torch.tensor([time.time() - t0], device="cuda")
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        tb = torch.tensor([float(h2d), float(d2h)], device="cuda")
        dist.all_reduce(tb)
        e2e_s = float(te.item()) / steps
        out["e2e"] = {"value": n_total / e2e_s, "unit": "particle-steps/s", "h2d_bytes_per_step": int(tb[0].item() / steps),
                      "d2h_bytes_per_step": int(tb[1].item() / steps), "ms_per_step": e2e_s * 1e3}
    return out


def run_ours_multi(args):
    """One process per GPU (torchrun), z-slab decomposition with NCCL inside the library."""
    import torch
    import torch.distributed as dist

    from particlesimulation_b200 import capi
    from particlesimulation_b200 import dist as pdist

    rank, world, local = pdist.init_process_group("nccl")
    assert world == args.gpus, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    cu = Cuda()
    cu.set_device(local)
    flush_bytes = 256 << 20
    flush = cu.malloc(flush_bytes)
    clocks = ClockSampler(local)   # every rank watches its own GPU
    clocks.start()
    if args.config == "mesh":
        prm, ic = c3_setup(capi, int(os.environ.get("P3M_BENCH_N", 1 << 24)), int(os.environ.get("P3M_BENCH_GRID", 512)), device=local,
                           margin=uniform_margin_cells(args))
        workload = f"C3: PM uniform cube, {ic.n} particles, {prm.nx}^3 mesh, strong scaling over {world} GPUs"
        scaling = "strong"
    elif args.config == "c4":
        prm, ic = c4_setup(capi, int(os.environ.get("P3M_BENCH_N", 1 << 26)), int(os.environ.get("P3M_BENCH_GRID", 1024)), device=local)
        workload = f"C4: P3M clustered disk + halo, {ic.n} particles, {prm.nx}^3 mesh, strong scaling over {world} GPUs"
        scaling = "strong"
    else:
        if world not in C5_GRIDS:
            raise SystemExit(f"--gpus {world}: the weak-scaling sweep is defined for 1/2/4/8 GPUs")
        prm, ic, grid = c5_setup(capi, world, int(os.environ.get("P3M_BENCH_N", N_C5_PER_GPU)), device=local,
                                 margin=uniform_margin_cells(args))
        workload = (f"C5 (BASELINE configs[4]): P3M weak-scaling sweep, 2^23 particles per GPU = {ic.n}, mesh "
                    f"{grid[0]}x{grid[1]}x{grid[2]}, ONE uniform distribution across all z-slabs (migration, ghost layers, "
                    f"plane exchanges and the slab FFT all active), TSC, S1-optimal Green, chaining-mesh PP (re=0.7a, a=3H)")
        scaling = "weak"
    r = measure_multi(args, capi, pdist, dist, torch, cu, prm, ic, flush, flush_bytes, clocks)
    clk_mine = clocks.stop()
    clk_all = [None] * world
    dist.all_gather_object(clk_all, clk_mine)
    clk = dict(clk_all[0])
    clk["per_rank"] = clk_all
    clk["reasons"] = sorted({x for c_ in clk_all for x in (c_.get("reasons") or [])})
    sm_all = [c_.get("sm_mhz") for c_ in clk_all if c_.get("sm_mhz")]
    if sm_all:
        clk["sm_mhz"] = float(min(sm_all))   # the slowest GPU of the job
    ctx = r.pop("ctx")
    pc = torch.tensor([0.0, 0.0], device="cuda", dtype=torch.float64)
    if prm.p3m:
        checked, inside = ctx.pair_counts()
        pc = torch.tensor([float(checked), float(inside)], device="cuda", dtype=torch.float64)
        dist.all_reduce(pc)
    ctx.close()
    extra = {}
    if args.config == "c2" and not args.no_extra:
        # the 1-GPU point of the same sweep, measured in this run on rank 0 (the others wait), so that the line
        # carries its own weak-scaling baseline
        if rank == 0:
            prm1, ic1, grid1 = c5_setup(capi, 1, int(os.environ.get("P3M_BENCH_N", N_C5_PER_GPU)), device=local,
                                        margin=uniform_margin_cells(args))
            r1 = run_single(args, capi, cu, prm1, ic1, "C5 at 1 GPU", flush, flush_bytes, want_e2e=False)
            r1["ctx"].close()
            v1 = ic1.n * args.steps / (float(np.sum(r1["ms"])) / 1e3)
            extra["c5_n1"] = {"workload": f"C5 at 1 GPU: 2^23 particles, mesh {grid1}", "value": v1,
                              "ms_per_step": float(np.mean(r1["ms"])), "ms_per_step_by_phase": r1["phases"],
                              "weak_scaling_efficiency_vs_this": r["value"] / (world * v1)}
        dist.barrier()
        if world == 8:
            for name, setup in (("c3", c3_setup), ("c4", c4_setup)):
                sprm, sic = setup(capi, device=local, margin=uniform_margin_cells(args))
                rr = measure_multi(args, capi, pdist, dist, torch, cu, sprm, sic, flush, flush_bytes,
                                   steps=max(args.steps, 10), want_e2e=False)
                c2x = rr.pop("ctx")
                spc = None
                if sprm.p3m:
                    ch, ins = c2x.pair_counts()
                    t = torch.tensor([float(ch), float(ins)], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t)
                    spc = {"pairs_checked_per_particle": float(t[0].item()) / sic.n, "pairs_in_range_per_particle": float(t[1].item()) / sic.n}
                c2x.close()
                extra[name] = {"workload": f"{name.upper()}: {'P3M clustered disk + halo' if sprm.p3m else 'PM uniform cube'}, "
                                           f"{sic.n} particles, {sprm.nx}^3 mesh, strong scaling over 8 GPUs",
                               "steps": max(args.steps, 10), "value": rr["value"], "ms_per_step": rr["ms_per_step"],
                               "ms_per_step_by_phase_min_median_max_over_ranks": rr["phases"], "exchange": rr["exchange"],
                               "parity": rr["parity"], "pairs": spc, "mesh_mode": "slab" if rr["slab"] else "replicated"}
    if rank == 0:
        line = {
            "metric": "P3M particle-steps/s", "value": r["value"], "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
            "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload, "particles": r["n_total"], "mesh": [prm.nx, prm.ny, prm.nz],
                       "l2": "256 MiB memset before every timed step (outside the events)",
                       "initial_conditions": "device-side sampler, every rank generates only its own z-slab",
                       "parallelism": f"z-slabs of particles over {world} GPUs (work-balanced cuts); NCCL over NVLink: migration, "
                                      f"ghost layers, density / potential plane exchange, slab-decomposed FFT with all-to-all "
                                      f"transpose (mesh mode: {'slab' if r['slab'] else 'replicated + all-reduce'})",
                       "note": "N = 1 of this bench is BASELINE configs[1] (C2, a different workload); the 1-GPU point of THIS "
                               "sweep is extra.c5_n1 (also in the N = 1 line)"},
            "clocks": clk, "gpu_launches": r["launches"], "e2e": r.get("e2e"),
            "ms_per_step_by_phase_min_median_max_over_ranks": r["phases"], "ms_per_step_by_phase_per_rank": r["per_rank"],
            "exchange": r["exchange"], "parity": r["parity"],
            "pairs_in_range_per_particle": float(pc[1].item()) / r["n_total"],
            "pairs_checked_per_particle": float(pc[0].item()) / r["n_total"],
            "extra": extra, "roofline": None, "cpu_baseline": None,
            "note": "roofline and cpu_baseline are reported by the N = 1 run",
        }
        print(json.dumps(line))
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="profiling runs: timed window only (no parity, companion, extras, CPU arm)")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra.* measurements (C5 at 1 GPU; C3 / C4 at 8 GPUs)")
    ap.add_argument("--config", default="c2", choices=["c2", "mesh", "c4", "c5"],
                    help="c2 = the driver's contract: BASELINE configs[1] at 1 GPU, the coupled weak-scaling sweep configs[4] at "
                         "N > 1; mesh = configs[2] (PM, 2^24 particles, 512^3; 1 GPU or strong scaling under torchrun); c4 = "
                         "configs[3] (P3M clustered disk + halo, 2^26 particles, 1024^3 mesh; torchrun)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        run_reference(args)
        return
    try:
        run_ours(args)
    except BaseException:
        if int(os.environ.get("WORLD_SIZE", 1)) > 1:
            # a failing rank must not linger in interpreter shutdown (destroying an NCCL communicator waits for
            # peers that are themselves waiting in a collective): report and leave, torchrun ends the job
            import traceback
            traceback.print_exc()
            sys.stderr.flush()
            sys.stdout.flush()
            os._exit(1)
        raise


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- the headline benchmark of BASELINE.json: P3M particle-steps/s.

  python bench.py --gpus N --steps K --warmup W            our arm   (C ABI -> sm_100a kernels)
  python bench.py --impl reference --steps K --warmup W     the reference's own CPU path, same host

One "step" = one pass of the hot path: drift -> bin/sort -> deposit -> Poisson -> gather -> short range
-> kick (the body of the reference's run loop, source/p3mMethod.cpp:101-153).  At N = 1 the workload is
BASELINE.json configs[1]: P3M Plummer sphere, 2^20 particles, 128^3 mesh, TSC, S1-optimal influence
function, a = 3H, re = 0.7a, eps = 0.5, tabulated short-range force (source/demos.cpp:1416-1447).

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
def c2_params(capi, timing=False):
    """BASELINE.json configs[1] with the reference demo's parameters (source/demos.cpp:1416-1447)."""
    f32 = np.float32
    p = capi.default_params()
    p.nx, p.ny, p.nz = GRID_C2
    p.box[:] = BOX_C2
    p.H = f32(f32(BOX_C2[0]) / f32(GRID_C2[0] // 2))
    p.DT, p.G = 1.0, 4.5e-3
    p.assignment, p.fd_scheme, p.greens_function = capi.TSC, capi.TWO_POINT, capi.S1_OPTIMAL
    p.particle_diameter = f32(f32(3) * f32(p.H))
    p.p3m = 1
    p.cutoff_radius = f32(f32(0.7) * f32(p.particle_diameter))
    p.softening = 0.5
    p.cloud_shape, p.use_sr_table = capi.S1, 1
    p.precision = capi.F32
    p.unit_roundtrip = 1
    p.green_zero_degenerate = 1
    p.timing = int(timing)
    return p


def c2_particles(n):
    from particlesimulation_b200 import ics
    return ics.plummer(n, center=(30.0, 30.0, 30.0), a=2.0, r_max=15.0, M=1.0, G=4.5e-3, seed=42)


def multi_params(capi, world, timing=False):
    """N > 1: one C2 cluster per GPU (weak scaling).  The box and the mesh grow along z with the rank
    count -- box 60 x 60 x 60 N, mesh 128 x 128 x 128 N -- every other parameter is C2's."""
    p = c2_params(capi, timing)
    p.nz = GRID_C2[2] * world
    p.box[2] = BOX_C2[2] * world
    return p


def c4_params(capi, grid, timing=False):
    """BASELINE.json configs[3]: P3M on a grid^3 mesh (1024 on 8 GPUs), TSC, S1-optimal influence function,
    a = 3H, re = 0.7a, the softening of C2 in mesh units."""
    f32 = np.float32
    p = capi.default_params()
    p.nx = p.ny = p.nz = grid
    p.box[:] = BOX_C2
    p.H = f32(f32(BOX_C2[0]) / f32(grid // 2))
    p.DT, p.G = 1.0, 4.5e-3
    p.assignment, p.fd_scheme, p.greens_function = capi.TSC, capi.TWO_POINT, capi.S1_OPTIMAL
    p.particle_diameter = f32(f32(3) * f32(p.H))
    p.p3m = 1
    p.cutoff_radius = f32(f32(0.7) * f32(p.particle_diameter))
    p.softening = f32(0.5 * float(p.H) / (60.0 / 64.0))
    p.cloud_shape, p.use_sr_table = capi.S1, 1
    p.precision = capi.F32
    p.unit_roundtrip = 1
    p.green_zero_degenerate = 1
    p.timing = int(timing)
    return p


def multi_particles(world, n_per_gpu):
    """One Plummer sphere per z-slab, centred in it; every rank builds the same global arrays."""
    from particlesimulation_b200 import ics
    pos, vel, mass = [], [], []
    for r in range(world):
        a, b, c = ics.plummer(n_per_gpu, center=(30.0, 30.0, 30.0 + 60.0 * r), a=2.0, r_max=15.0, M=1.0, G=4.5e-3,
                              seed=42 + r)
        pos.append(a); vel.append(b); mass.append(c)
    return np.concatenate(pos), np.concatenate(vel), np.concatenate(mass)


def run_mesh_multi(args):
    """Secondary measurement on N GPUs (torchrun): BASELINE configs[2] -- PM uniform cube, 2^24 particles,
    512^3 mesh, slab-decomposed FFT across the ranks.  Prints ms/step by phase (max over ranks)."""
    import torch
    import torch.distributed as dist
    from particlesimulation_b200 import capi, ics
    from particlesimulation_b200 import dist as pdist
    rank, world, local = pdist.init_process_group("nccl")
    cu = Cuda(); cu.set_device(local)
    n = int(os.environ.get("P3M_BENCH_N", 1 << 24))
    grid = int(os.environ.get("P3M_BENCH_GRID", 512))
    f32 = np.float32
    def params(timing):
        p = capi.default_params()
        p.nx = p.ny = p.nz = grid
        p.box[:] = (60.0, 60.0, 60.0)
        p.H = f32(f32(60.0) / f32(grid // 2))
        p.DT, p.G = 1.0, 4.5e-3
        p.assignment, p.fd_scheme, p.greens_function = capi.TSC, capi.TWO_POINT, capi.DISCRETE_LAPLACIAN
        p.p3m = 0
        p.timing = int(timing)
        p.device = local
        return p
    H = float(params(0).H)
    pos, vel, mass = ics.uniform_cube(n, [2 * H] * 3, [60.0 - 2 * H] * 3, total_mass=1.0, seed=42)
    out = {}
    for timing in (0, 1):
        ctx = pdist.create_context(params(timing), capi)
        ctx.set_particles(pos, vel, mass)
        ctx.green_init(); ctx.force(); ctx.kick(0.5)
        for _ in range(args.warmup):
            ctx.step(1)
        dist.barrier(); cu.sync()
        if not timing:
            a, b = cu.event(), cu.event()
            cu.record(a, ctx.stream); ctx.step(args.steps); cu.record(b, ctx.stream)
            t = torch.tensor([cu.elapsed_ms(a, b) / args.steps], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["ms_per_step"] = float(t.item())
            out["slab"] = ctx.rank_info()["slab"]
            out["n_local"] = ctx.n
        else:
            ctx.phase_ms(reset=True); ctx.step(args.steps)
            ph = ctx.phase_ms()
            names = sorted(ph)
            t = torch.tensor([ph[k] / args.steps for k in names], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            out["phases"] = dict(zip(names, [round(float(x), 4) for x in t.tolist()]))
        ctx.close()
    if rank == 0:
        M = grid ** 3
        nxh = grid // 2 + 1
        a2a = 2 * 2 * 8.0 * nxh * grid * grid / world * (world - 1) / world  # bytes sent per GPU per solve (fwd + inv)
        print(json.dumps({"metric": "PM particle-steps/s (secondary, mesh-dominated, multi-GPU)",
                          "value": n / (out["ms_per_step"] / 1e3), "unit": "particle-steps/s", "n_gpus": world,
                          "ms_per_step": out["ms_per_step"], "scaling": "strong",
                          "config": {"workload": f"C3: PM uniform cube, {n} particles, {grid}^3 mesh, TSC, 2-pt, discrete "
                                                 f"Laplacian, slab-decomposed FFT over {world} GPUs",
                                     "mesh_mode": "slab" if out["slab"] else "replicated"},
                          "ms_per_step_by_phase_max_over_ranks": out["phases"],
                          "all_to_all_bytes_sent_per_gpu_per_step": a2a, "mesh_cells": M}))
    dist.barrier()
    dist.destroy_process_group()


def run_mesh_config(args):
    """Secondary measurement (not the headline line): BASELINE configs[2]-style PM run on ONE GPU --
    uniform cube, 2^24 particles, 512^3 mesh, TSC, 2-point, discrete-Laplacian Green -- where the mesh
    kernels (deposit, FFT, gather, sort, integrate) are the whole step.  Prints per-kernel HBM fractions."""
    from particlesimulation_b200 import capi, ics
    cu = Cuda(); cu.set_device(0)
    n = int(os.environ.get("P3M_BENCH_N", 1 << 24))
    grid = int(os.environ.get("P3M_BENCH_GRID", 512))
    hbm_peak, peak_src, _ = load_peaks()
    f32 = np.float32
    def params(timing):
        p = capi.default_params()
        p.nx = p.ny = p.nz = grid
        p.box[:] = (60.0, 60.0, 60.0)
        p.H = f32(f32(60.0) / f32(grid // 2))
        p.DT, p.G = 1.0, 4.5e-3
        p.assignment, p.fd_scheme, p.greens_function = capi.TSC, capi.TWO_POINT, capi.DISCRETE_LAPLACIAN
        p.p3m = 0
        p.timing = int(timing)
        return p
    H = float(params(0).H)
    pos, vel, mass = ics.uniform_cube(n, [2 * H] * 3, [60.0 - 2 * H] * 3, total_mass=1.0, seed=42)
    out = {}
    for timing in (0, 1):
        ctx = capi.Context(params(timing))
        ctx.set_particles(pos, vel, mass)
        ctx.green_init(); ctx.force(); ctx.kick(0.5)
        for _ in range(args.warmup):
            ctx.step(1)
        if not timing:
            a, b = cu.event(), cu.event()
            cu.record(a, ctx.stream); ctx.step(args.steps); cu.record(b, ctx.stream)
            out["ms_per_step"] = cu.elapsed_ms(a, b) / args.steps
        else:
            ctx.phase_ms(reset=True); ctx.step(args.steps)
            out["phases"] = {k: v / args.steps for k, v in ctx.phase_ms().items()}
        ctx.close()
    M = grid ** 3
    ph = out["phases"]
    def hbm(b, ms):
        g = b / (ms / 1e3) / 1e9 if ms > 0 else 0.0
        return {"achieved_GBs": round(g, 1), "frac_of_measured_hbm": round(g / hbm_peak, 4), "ms": round(ms, 4), "algorithmic_bytes": b}
    line = {"metric": "PM particle-steps/s (secondary, mesh-dominated)", "value": n / (out["ms_per_step"] / 1e3),
            "unit": "particle-steps/s", "n_gpus": 1, "ms_per_step": out["ms_per_step"],
            "config": {"workload": f"C3-style PM: uniform cube, {n} particles, {grid}^3 mesh, TSC, 2-pt, discrete Laplacian"},
            "roofline_kernels": {
                "binSort": hbm(76.0 * n, ph["binSort"]), "spreadMass": hbm(16.0 * n + 8.0 * M, ph["spreadMass"]),
                "poisson": hbm(18.0 * M, ph["forwardFFT"] + ph["fourierPotential"] + ph["inverseFFT"]),
                "updateAccelerations": hbm(4.0 * M + 28.0 * n, ph["updateAccelerations"]),
                "integrate": hbm(96.0 * n, ph["integrate"])},
            "ms_per_step_by_phase": {k: round(v, 4) for k, v in ph.items()},
            "hbm_peak": hbm_peak, "hbm_peak_source": peak_src}
    print(json.dumps(line))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md)", 1965.0


def profile_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the newest committed
    `ncu --set full` summary under profiles/ (tools/summarize_ncu.py); None if there is none."""
    import glob
    import re
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", f"{kernel}_*.txt")), key=os.path.getmtime)
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
    pos, vel, mass = c2_particles(n)
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


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    from particlesimulation_b200 import capi

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    if world > 1 or args.gpus > 1:
        return run_ours_multi(args, rank, world, local)

    cu = Cuda()
    cu.set_device(0)
    n = int(os.environ.get("P3M_BENCH_N", N_C2))
    pos, vel, mass = c2_particles(n)
    hbm_peak, peak_src, sm_max = load_peaks()

    ctx = capi.Context(c2_params(capi))
    stream = ctx.stream
    ctx.set_particles(pos, vel, mass)
    ctx.green_init()
    ctx.force()
    ctx.kick(0.5)  # setHalfStepVelocities
    flush_bytes = 256 << 20
    flush = cu.malloc(flush_bytes)
    clocks = ClockSampler(0)
    clocks.start()
    for _ in range(args.warmup):
        ctx.step(1)
    ev = [(cu.event(), cu.event()) for _ in range(args.steps)]
    launches0 = ctx.launches
    cu.sync()
    clocks.mark()
    t_wall0 = time.time()
    for a, b in ev:
        cu.check(cu.rt.cudaMemsetAsync(flush, 0, flush_bytes, stream), "flush")  # evict L2 (126 MB)
        cu.record(a, stream)
        ctx.step(1)
        cu.record(b, stream)
    cu.sync()
    t_wall = time.time() - t_wall0
    clk = clocks.stop()
    launches = ctx.launches - launches0
    ms = [cu.elapsed_ms(a, b) for a, b in ev]
    total_ms = float(np.sum(ms))
    value = n * args.steps / (total_ms / 1e3)

    # ---- end to end through the C ABI with host buffers (pinned), one step per call triple
    hp, hv, hm = cu.pinned((n, 3)), cu.pinned((n, 3)), cu.pinned((n,))
    op, ov = cu.pinned((n, 3)), cu.pinned((n, 3))
    cpos, cvel, _ = ctx.get_particles(capi.UNITS_ORIGINAL)
    hp[:], hv[:], hm[:] = cpos, cvel, mass
    e2e_steps = max(2, min(args.steps, 5))
    lib = capi.lib()
    t0 = None
    for it in range(e2e_steps + 1):
        if it == 1:
            cu.sync()
            t0 = time.time()
        ctx.set_particles(hp, hv, hm)          # H2D: pos, vel, mass (28 B / particle)
        ctx.step(1)                             # drift -> force -> kick (self-contained: needs x, v, m only)
        rc = lib.p3m_get_particles(ctx._h, op.ctypes.data_as(C.c_void_p), ov.ctypes.data_as(C.c_void_p), None,
                                   capi.UNITS_ORIGINAL)  # D2H: pos, vel (24 B / particle)
        assert rc == 0
        hp, op = op, hp                         # the caller's next step starts from what came back
        hv, ov = ov, hv
    e2e_s = (time.time() - t0) / e2e_steps
    e2e_value = n / e2e_s

    # ---- per-phase breakdown and roofline figures from a timing context (CUDA events per phase)
    tctx = capi.Context(c2_params(capi, timing=True))
    tctx.set_particles(pos, vel, mass)
    tctx.green_init()
    tctx.force()
    tctx.kick(0.5)
    tctx.step(2)
    tctx.phase_ms(reset=True)
    psteps = 3
    tctx.step(psteps)
    phases = {k: v / psteps for k, v in tctx.phase_ms().items()}
    checked, inside = tctx.pair_counts()
    tctx.close()
    M = GRID_C2[0] * GRID_C2[1] * GRID_C2[2]
    # flops = 9 * P_checked + 14 * P_in
    pp_flops = 9.0 * checked + 14.0 * inside
    pp_ms = phases["shortRangeForcesCalc"]
    sm_count = 148
    fp32_peak = sm_count * 128 * 2 * sm_max * 1e6 / 1e12  # TFLOP/s at clocks.max.sm, FMA = 2 flop
    pp_tflops = pp_flops / (pp_ms / 1e3) / 1e12 if pp_ms > 0 else 0.0

    def hbm(bytes_, ms_):
        gbs = bytes_ / (ms_ / 1e3) / 1e9 if ms_ > 0 else 0.0
        return {"bound": "hbm", "achieved": gbs, "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak,
                "algorithmic_bytes": bytes_, "ms": ms_}

    kernels = {
        "binSort": hbm(76.0 * n, phases["binSort"]),
        "spreadMass": hbm(16.0 * n + 8.0 * M, phases["spreadMass"]),
        "poisson(fwdFFT+multiply+invFFT)": hbm(18.0 * M, phases["forwardFFT"] + phases["fourierPotential"] + phases["inverseFFT"]),
        "updateAccelerations(fused gradient+gather)": hbm(4.0 * M + 28.0 * n, phases["updateAccelerations"]),
        "integrate": hbm(2 * 48.0 * n, phases["integrate"]),
    }
    roofline = {"kernel": "k_pp_tiled (short-range PP, dense chaining cells)", "bound": "fp32",
                "achieved": pp_tflops, "peak": fp32_peak, "unit": "TFLOP/s",
                "frac": pp_tflops / fp32_peak if fp32_peak else None, "traffic": profile_traffic("k_pp_tiled")[0],
                "traffic_source": f"dram bytes read + written per launch, ncu --set full capture profiles/{profile_traffic('k_pp_tiled')[1]}",
                "peak_source": f"148 SMs x 128 FP32 lanes x 2 flop x clocks.max.sm {sm_max:.0f} MHz (no tensor cores on "
                               f"this path; HBM peak for the other kernels: {peak_src})",
                "flops_per_launch": pp_flops, "pairs_checked": checked, "pairs_in_range": inside,
                "ms": pp_ms}

    cpu = None
    if not args.no_cpu_baseline:
        try:
            cpu = run_reference(args, as_baseline=True)
        except Exception as e:  # the GPU numbers stand on their own
            cpu = {"value": None, "unit": "particle-steps/s", "cores": 0, "kind": "reference",
                   "sample": f"failed: {e}"}

    line = {
        "metric": "P3M particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C2: P3M Plummer sphere, 2^20 particles, 128^3 mesh, TSC, S1-optimal Green, "
                               "chaining-mesh PP (re=0.7a, a=3H, eps=0.5, table)", "particles": n,
                   "mesh": list(GRID_C2), "l2": "256 MiB memset between timed steps (outside the events)",
                   "parallelism": "1 GPU"},
        "clocks": clk, "gpu_launches": launches,
        "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": 28 * n,
                "d2h_bytes_per_step": 24 * n, "ms_per_step": e2e_s * 1e3},
        "roofline": roofline, "roofline_kernels": kernels,
        "ms_per_step_by_phase": phases, "ms_per_step_each": ms, "wall_s_timed_region": t_wall,
        "cpu_baseline": cpu,
    }
    print(json.dumps(line))
    ctx.close()
    return line


def run_ours_multi(args, rank, world, local):
    """One process per GPU (torchrun), z-slab decomposition with NCCL inside the library."""
    import torch
    import torch.distributed as dist

    from particlesimulation_b200 import capi
    from particlesimulation_b200 import dist as pdist

    rank, world, local = pdist.init_process_group("nccl")
    assert world == args.gpus, f"WORLD_SIZE {world} != --gpus {args.gpus}"
    cu = Cuda()
    cu.set_device(local)
    c4 = args.config == "c4"
    if c4:
        from particlesimulation_b200 import ics
        n_total = int(os.environ.get("P3M_BENCH_N", 1 << 26))
        grid = int(os.environ.get("P3M_BENCH_GRID", 1024))
        pos, vel, mass = ics.clustered_disk_halo(n_total, seed=42)
        prm = c4_params(capi, grid)
        workload = (f"C4: P3M clustered disk + halo, {n_total} particles, {grid}^3 mesh, TSC, S1-optimal Green, "
                    f"chaining-mesh PP (re=0.7a, a=3H), strong scaling over {world} GPUs")
        mesh = [grid, grid, grid]
    else:
        n_per = int(os.environ.get("P3M_BENCH_N", N_C2))
        pos, vel, mass = multi_particles(world, n_per)
        n_total = len(mass)
        prm = multi_params(capi, world)
        workload = (f"C5-style weak scaling of C2: one P3M Plummer sphere of 2^20 particles per GPU, "
                    f"mesh 128x128x{128 * world}, TSC, S1-optimal Green, chaining-mesh PP")
        mesh = [GRID_C2[0], GRID_C2[1], GRID_C2[2] * world]
    prm.device = local
    ctx = pdist.create_context(prm, capi)
    stream = ctx.stream
    ctx.set_particles(pos, vel, mass)
    ctx.green_init()
    ctx.force()
    ctx.kick(0.5)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    for _ in range(args.warmup):
        ctx.step(1)
    flush_bytes = 256 << 20
    flush = cu.malloc(flush_bytes)
    ev = [(cu.event(), cu.event()) for _ in range(args.steps)]
    launches0 = ctx.launches
    dist.barrier()
    cu.sync()
    clocks.mark()
    for a, b in ev:
        cu.check(cu.rt.cudaMemsetAsync(flush, 0, flush_bytes, stream), "flush")
        cu.record(a, stream)
        ctx.step(1)
        cu.record(b, stream)
    cu.sync()
    dist.barrier()
    clk = clocks.stop() if rank == 0 else None
    ms = [cu.elapsed_ms(a, b) for a, b in ev]
    t = torch.tensor([float(np.sum(ms))], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)  # device time, max over ranks
    total_ms = float(t.item())
    launches = ctx.launches - launches0
    nloc = torch.tensor([ctx.n], device="cuda")
    nmin, nmax = nloc.clone(), nloc.clone()
    dist.all_reduce(nmin, op=dist.ReduceOp.MIN)
    dist.all_reduce(nmax, op=dist.ReduceOp.MAX)

    # end to end: every rank uploads the particles it holds (pinned host buffers, explicit ids), steps,
    # and reads them back into the other set of pinned buffers (ping-pong: no host-side copies in the loop)
    cap_h = int(1.25 * max(ctx.n, n_total // world)) + 4096
    bufs = [(cu.pinned((cap_h,), np.int32), cu.pinned((cap_h, 3)), cu.pinned((cap_h, 3))) for _ in range(2)]
    hm = cu.pinned((cap_h,))
    equal_mass = bool(mass.min() == mass.max())
    hm[:] = mass[0]  # equal masses (the samplers' M / n): no per-step gather; else refreshed from the ids below
    ids, lp, lv, _ = ctx.get_local(capi.UNITS_ORIGINAL, out=bufs[0])
    nl = len(ids)
    e2e_steps = max(2, min(args.steps, 5))
    h2d = d2h = 0
    t0 = None
    cur = 0
    for it in range(e2e_steps + 1):
        if it == 1:
            dist.barrier(); cu.sync(); t0 = time.time(); h2d = d2h = 0
        hid, hp, hv = bufs[cur]
        if not equal_mass:
            hm[:nl] = mass[hid[:nl]]
        ctx.set_particles_ids(hp[:nl], hv[:nl], hm[:nl], hid[:nl])
        h2d += 32 * nl
        ctx.step(1)
        if ctx.n > cap_h:
            raise RuntimeError("end-to-end buffers too small for this rank's share")
        ids, lp, lv, _ = ctx.get_local(capi.UNITS_ORIGINAL, out=bufs[cur ^ 1])
        nl = len(ids)
        d2h += 28 * nl
        cur ^= 1
    cu.sync(); dist.barrier()
    te = torch.tensor([time.time() - t0], device="cuda")
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item()) / e2e_steps
    tb = torch.tensor([float(h2d), float(d2h)], device="cuda")
    dist.all_reduce(tb)
    extra = {}
    if c4:
        # phase breakdown (max over ranks) and exact pair statistics from a timing context
        ctx.close()
        tprm = c4_params(capi, grid, timing=True)
        tprm.device = local
        ctx = pdist.create_context(tprm, capi)
        ctx.set_particles(pos, vel, mass)
        ctx.green_init(); ctx.force(); ctx.kick(0.5); ctx.step(1)
        ctx.phase_ms(reset=True)
        ctx.step(2)
        ph = ctx.phase_ms()
        names = sorted(ph)
        tp = torch.tensor([ph[k] / 2 for k in names], device="cuda")
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
        checked, inside = ctx.pair_counts()
        pc = torch.tensor([float(checked), float(inside)], device="cuda", dtype=torch.float64)
        dist.all_reduce(pc)
        extra = {"ms_per_step_by_phase_max_over_ranks": dict(zip(names, [round(float(x), 3) for x in tp.tolist()])),
                 "pairs_in_range_per_particle": float(pc[1].item()) / n_total,
                 "pairs_checked_per_particle": float(pc[0].item()) / n_total}
    if rank == 0:
        value = n_total * args.steps / (total_ms / 1e3)
        line = {
            "metric": "P3M particle-steps/s", "value": value, "unit": "particle-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
            "higher_is_better": True, "scaling": "strong" if c4 else "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload, "particles": n_total, "mesh": mesh, **extra,
                       "l2": "256 MiB memset between timed steps (outside the events)",
                       "parallelism": f"z-slabs of particles over {world} GPUs; NCCL over NVLink: migration, ghost layers, density/"
                                      f"potential plane exchange, slab-decomposed FFT with all-to-all transpose "
                                      f"(mesh mode: {'slab' if ctx.rank_info()['slab'] else 'replicated + all-reduce'})",
                       "particles_per_rank_min_max": [int(nmin.item()), int(nmax.item())]},
            "clocks": clk, "gpu_launches": launches,
            "e2e": {"value": n_total / e2e_s, "unit": "particle-steps/s",
                    "h2d_bytes_per_step": int(tb[0].item() / e2e_steps), "d2h_bytes_per_step": int(tb[1].item() / e2e_steps),
                    "ms_per_step": e2e_s * 1e3},
            "roofline": None, "cpu_baseline": None,
            "note": "roofline and cpu_baseline are reported by the N = 1 run",
        }
        print(json.dumps(line))
    ctx.close()
    dist.barrier()
    dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "mesh", "c4"],
                    help="c2 = the headline workload; mesh = secondary mesh-dominated PM run; c4 = BASELINE configs[3] "
                         "(P3M clustered disk+halo, 2^26 particles, 1024^3 mesh; torchrun, 8 GPUs)")
    args = ap.parse_args()
    if args.impl == "reference":
        if int(os.environ.get("RANK", 0)) != 0:
            return
        run_reference(args)
        return
    if args.config == "mesh":
        if int(os.environ.get("WORLD_SIZE", 1)) > 1:
            run_mesh_multi(args)
        else:
            run_mesh_config(args)
        return
    run_ours(args)


if __name__ == "__main__":
    main()
